"""Cross-GPU parity of the live product path (tools/parity_check.py) under one process per rank.

On a box with >= 2 GPUs every rank owns its own GPU, so the compared bytes crossed
NVLink/NVSwitch through CUDA-IPC mapped slabs (`test_parity_across_physical_gpus`, skipped
below 2 GPUs).  The same check also runs with two ranks sharing cuda:0 so that the single-GPU
round-end box still exercises it.

Checked, all bit-for-bit (integer path) unless a tolerance is stated:
  * what each peer wrote into my slab (packed bytes, bf16 params) == C oracle restatement of
    op_util.py:189-209 on the SENDER's host, every byte of every (peer, bit) segment;
  * dequantised halo rows == oracle unpack (op_util.py:211-236); fp32 exchange rows;
  * forward0 halo of the product == halo of the reference flow around the REFERENCE's own
    quant_cuda kernels (oracle/_ref), same generator seed;
  * one training step, product vs the reference flow (sequential form; its helper-thread overlap
    reads send_messages across streams without an event and is not deterministic): final-layer
    activations and first-layer weight gradient within parity_check.ACT_TOL (max, relative to the
    largest magnitude), ACT_MEAN_TOL (mean) and LOSS_TOL;
  * the flattened NCCL all-reduce of the gradients == per-parameter gloo
    all-reduce (runtime_util.py:71-77) -- bit-exact at W=2 (one fp32 add per element,
    commutative), <= 1e-6 relative otherwise (reduction order).
"""
import os
import socket
import sys
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmp, mode, model_name, scheme, dataset, ngpu, out):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank % ngpu),
                       "ADAQP_SYNTH_SCALE": "0.004" if dataset == "ogbn-products" else "0.01", "ADAQP_SEED": "11",
                       "ADAQP_SYNTHETIC": "1"})
    sys.path.insert(0, ROOT)
    os.chdir(tmp)
    from argparse import Namespace
    from adaqp_b200 import Trainer
    from adaqp_b200.communicator import Communicator as comm
    from adaqp_b200.trainer import runtime_util as ru
    from tools import parity_check
    args = Namespace(dataset=dataset, num_parts=world, backend="gloo", init_method="env://",
                     model_name=model_name, mode=mode, assign_scheme=scheme, logger_level="WARNING",
                     num_epoches=2, exp_path=f"{tmp}/exp")
    tr = Trainer(args)
    ru.sync_seed()
    tr.model.reset_parameters()
    ru.sync_model(tr.model)
    res = parity_check.exchange_parity(window_groups=0)
    res["activations"] = parity_check.activation_parity(tr)
    if os.environ.get("ADAQP_TEST_OVERLAPPED_REF") == "1":
        res["activations_overlapped_ref"] = parity_check.activation_parity(tr, overlapped=True)
    # 8f-1: bucketed NCCL reduction vs the reference's per-parameter gloo reduction
    res["nccl"] = None
    grp = ru._reduce_group()
    if grp is not None:
        grads = [p.grad.detach().clone() for p in tr.model.parameters() if p.grad is not None]
        want = [g.clone() for g in grads]
        for g in want:
            comm.all_reduce_sum(g)                       # gloo, one call per parameter
        ru.average_gradients(tr.model)                   # flattened bucket over NCCL
        got = [p.grad for p in tr.model.parameters() if p.grad is not None]
        worst = max(float((a - b).abs().max() / (b.abs().max() + 1e-30)) for a, b in zip(got, want))
        exact = all(torch.equal(a, b) for a, b in zip(got, want))
        res["nccl"] = {"max_rel": worst, "bit_exact": exact}
    torch.cuda.synchronize()
    comm.ctx.delete_buffer()
    out.put((rank, res))


def _run(world, ngpu, mode, model_name, scheme, dataset="ogbn-products"):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    with tempfile.TemporaryDirectory() as tmp:
        procs = [ctx.Process(target=_worker, args=(r, world, port, tmp, mode, model_name, scheme, dataset, ngpu, out))
                 for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=900)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        res = dict(out.get(timeout=5) for _ in procs)
    return res[0]


def _check(res, world, quant=True):
    from tools.parity_check import ACT_MEAN_TOL, ACT_TOL, LOSS_TOL
    assert res["mismatches"] == 0, res
    assert res["fp32_values_compared"] > 0
    if quant:
        assert res["bytes_compared"] > 0 and res["halo_values_compared"] > 0 and res["params_compared"] > 0
    act = res["activations"]
    print("PARITY", {k: v for k, v in res.items() if not k.startswith("activations")}, act, res.get("activations_overlapped_ref"))
    assert "error" not in act and "unavailable" not in act, act
    if quant:
        assert act["forward0_halo_vs_reference_kernels"]["mismatches"] == 0, act
    assert act["act_max_rel_err"] <= ACT_TOL and act["act_mean_rel_err"] <= ACT_MEAN_TOL, act
    assert act["grad0_max_rel_err"] <= ACT_TOL and act["loss_rel_diff"] <= LOSS_TOL, act


@pytest.mark.parametrize("mode,model_name,scheme,dataset", [
    ("AdaQP", "gcn", "random", "ogbn-products"), ("AdaQP-q", "sage", "random", "ogbn-products"),
    ("AdaQP", "gcn", "uniform", "reddit"), ("Vanilla", "gcn", "uniform", "ogbn-products"),
    ("AdaQP", "gcn", "uniform", "ogbn-products"), ("AdaQP-q", "gcn", "random", "ogbn-products"),
    ("AdaQP-p", "gcn", "uniform", "ogbn-products")])
def test_parity_two_ranks(mode, model_name, scheme, dataset):
    ngpu = torch.cuda.device_count()
    res = _run(2, ngpu, mode, model_name, scheme, dataset)
    _check(res, 2, quant=mode in ("AdaQP", "AdaQP-q"))


@pytest.mark.parametrize("mode,model_name,scheme", [("AdaQP", "gcn", "random"), ("AdaQP", "sage", "random")])
def test_parity_across_physical_gpus(mode, model_name, scheme):
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (one per rank)")
    world = 4 if ngpu >= 4 else 2
    res = _run(world, ngpu, mode, model_name, scheme)
    assert res["distinct_gpus"] == world
    _check(res, world)
    assert res["nccl"] is not None, "NCCL bucket path not taken with one GPU per rank"
    if world == 2:
        assert res["nccl"]["bit_exact"], res["nccl"]
    assert res["nccl"]["max_rel"] <= 1e-6, res["nccl"]


def _evalcache_worker(rank, world, port, tmp, ngpu, out):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world),
                       "LOCAL_RANK": str(rank % ngpu), "ADAQP_SYNTH_SCALE": "0.004", "ADAQP_SEED": "3", "ADAQP_SYNTHETIC": "1"})
    sys.path.insert(0, ROOT)
    os.chdir(tmp)
    from argparse import Namespace
    from adaqp_b200 import Trainer
    from adaqp_b200.communicator import Communicator as comm
    from adaqp_b200.manager import GraphEngine as engine
    from adaqp_b200.trainer import runtime_util as ru
    tr = Trainer(Namespace(dataset="ogbn-products", num_parts=world, backend="gloo", init_method="env://", model_name="gcn",
                           mode="AdaQP", assign_scheme="uniform", logger_level="WARNING", num_epoches=2, exp_path=f"{tmp}/exp"))
    ru.sync_seed()
    tr.model.reset_parameters()
    ru.sync_model(tr.model)
    eng, ex = engine.ctx, comm.ctx.comm_buffer.p2p
    tr.model.eval()

    def fwd():
        with torch.no_grad():
            y = tr.model(eng.graph, eng.feats).clone()
        eng.timer.clear(is_train=False)
        torch.cuda.synchronize()
        return y

    res = {}
    s0 = ex.seq["test0"]
    a = fwd()                                   # computes and caches the layer-0 exchange + aggregation
    s1 = ex.seq["test0"]
    b = fwd()                                   # hit: no layer-0 exchange
    s2 = ex.seq["test0"]
    res["first_pass_exchanged"] = s1 - s0 == 1
    res["hit_skips_exchange"] = s2 == s1
    res["hit_bit_exact"] = bool(torch.equal(a, b))
    os.environ["ADAQP_EVAL_CACHE"] = "0"
    c = fwd()                                   # cache off: recomputed
    del os.environ["ADAQP_EVAL_CACHE"]
    res["uncached_bit_exact"] = bool(torch.equal(a, c)) and ex.seq["test0"] == s2 + 1
    eng.feats[: eng.feats.shape[0] // 2] += 0.25       # in-place mutation bumps the version counter
    s3 = ex.seq["test0"]
    d = fwd()
    res["mutation_invalidates"] = ex.seq["test0"] == s3 + 1 and not torch.equal(a, d)
    os.environ["ADAQP_EVAL_CACHE"] = "0"
    e = fwd()
    del os.environ["ADAQP_EVAL_CACHE"]
    res["after_mutation_bit_exact"] = bool(torch.equal(d, e))
    other = eng.feats.clone()                           # a DIFFERENT tensor never hits (identity, not address)
    with torch.no_grad():
        f = tr.model(eng.graph, other)
    eng.timer.clear(is_train=False)
    res["other_tensor_recomputed"] = bool(torch.equal(f, d))
    ex.check_status()
    torch.cuda.synchronize()
    comm.ctx.delete_buffer()
    out.put((rank, res))


def test_eval_layer0_cache_on_gpu():
    """the evaluation forward's constant layer-0 exchange + aggregation is cached; the cached
    output equals the recomputed one bit for bit, the cached pass does not exchange, and an in-place change of
    the feature matrix invalidates it (trainer.py:181 evaluates every epoch)."""
    ngpu = torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    with tempfile.TemporaryDirectory() as tmp:
        procs = [ctx.Process(target=_evalcache_worker, args=(r, 2, port, tmp, ngpu, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=600)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        res = dict(out.get(timeout=5) for _ in procs)
    for r in (0, 1):
        assert all(res[r].values()), res[r]
