"""world_size-2 gloo runs on CPU (BASELINE.json configs[0] plumbing): control plane, index
contracts, Trainer loop and the gloo data path, without a GPU."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, tmp, mode, model_name, out, dataset="reddit"):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank), "ADAQP_DEVICE": "cpu",
                       "ADAQP_SYNTH_SCALE": "0.004" if dataset == "reddit" else "0.003", "ADAQP_SEED": "7", "OMP_NUM_THREADS": "1", "ADAQP_SYNTHETIC": "1"})
    sys.path.insert(0, ROOT)
    os.chdir(tmp)
    from argparse import Namespace
    from adaqp_b200 import Trainer
    from adaqp_b200.communicator import Communicator as comm
    from adaqp_b200.manager import GraphEngine as engine
    from adaqp_b200.model.op_util import msg_all2all_GLOO
    args = Namespace(dataset=dataset, num_parts=world, backend="gloo", init_method="env://", model_name=model_name,
                     mode=mode, assign_scheme="uniform", logger_level="WARNING", num_epoches=3, exp_path=f"{tmp}/exp")
    tr = Trainer(args)
    eng = engine.ctx
    # exchange check: what arrives at halo position j is the owner's feature row of that node
    send = eng.feats[eng.total_send_idx]
    remote = msg_all2all_GLOO(send, "forward0", is_train=False)
    eng.timer.clear(is_train=False)
    gathered = comm.gather_all({"feat_sum": eng.feats.sum(1), "send_rows": send.sum(1), "send_idx": eng.send_idx})
    for p, idx in eng.recv_idx.items():
        lo, hi = gathered[p]["send_idx"][rank]
        assert torch.allclose(remote[idx].sum(1), gathered[p]["send_rows"][lo:hi])
    # evaluation forward: layer 0 (constant features, fp32 exchange) is computed once and reused
    tr.model.eval()
    with torch.no_grad():
        a = tr.model(eng.graph, eng.feats)
        eng.timer.clear(is_train=False)
        b = tr.model(eng.graph, eng.feats)
    assert torch.equal(a, b)
    if getattr(eng, "marginal_pool", None) is not None:      # overlap modes: helper-thread exchange == synchronous exchange
        pool, eng.marginal_pool = eng.marginal_pool, None
        with torch.no_grad():
            eng.timer.clear(is_train=False)
            c = tr.model(eng.graph, eng.feats)
        eng.marginal_pool = pool
        assert torch.equal(a, c)
    assert not any(k.startswith("forward0") for k in list(eng.timer._record) + list(eng.timer._events))
    assert any(k.startswith("forward1") for k in list(eng.timer._record) + list(eng.timer._events))
    eng.timer.clear(is_train=False)
    rec = tr.train()
    tr.save(rec)
    losses_ok = bool(torch.isfinite(rec).all())
    out.put((rank, losses_ok, float(eng.recorder.epoches_metrics[:3, 0].max())))


@pytest.mark.parametrize("mode,model_name,dataset", [("Vanilla", "gcn", "reddit"), ("AdaQP-p", "sage", "reddit"),
                                                     ("Vanilla", "sage", "yelp")])
def test_two_rank_cpu_training(mode, model_name, dataset):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    with tempfile.TemporaryDirectory() as tmp:
        procs = [ctx.Process(target=_worker, args=(r, 2, port, tmp, mode, model_name, out, dataset)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=600)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        res = sorted(out.get(timeout=5) for _ in procs)
        assert all(ok for _, ok, _ in res)
        assert os.path.exists(f"{tmp}/exp/{dataset}/2part/{model_name}/time/{mode}.csv")
