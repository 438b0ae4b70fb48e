"""Cross-GPU parity of the LIVE product exchange against the CPU oracle and the reference flow.

TEST INFRASTRUCTURE (imports oracle/): used by `bench.py --verify` (outside the timed region) and
by tests/test_gpu_multigpu.py.  Runs inside an initialised product `Trainer` (one process per
rank; one GPU per rank on a multi-GPU box, so every byte compared below crossed NVLink/NVSwitch
through the P2P stores of csrc/exchange.cu).

exchange_parity()
    For each layer key the ranks run one real exchange on seeded inputs.  Every SENDER restates,
    with the C oracle on its host, windows of byte-rows of every (peer, bit-width) segment it
    sent -- row min/max, scale, bf16 params, stochastic quantise + pack with the segment's Philox
    (seed, offset), exactly AdaQP/model/op_util.py:189-209 -- and hands the expected bytes to the
    receivers over the gloo control plane.  Every RECEIVER compares them bit-for-bit with what
    landed in its slab (`recv_region`, the reference's wire format) and compares the dequantised
    halo rows with the oracle's unpack (op_util.py:211-236).  Windows = first, last (ragged tail)
    and evenly spaced interior byte-rows; `window_groups <= 0` compares every byte (tests).

activation_parity()
    One training step (forward + backward) of the product model and of oracle/ref_path.py -- the
    reference's flow around the REFERENCE's own quant_cuda kernels built into oracle/_ref -- from
    the same weights, the same bit assignment and the same CUDA generator seed: first the
    quantised halo of `forward0` (identical inputs => bit-identical, reference kernels vs ours),
    then final-layer activations and the first layer's weight gradient (tolerance stated there).
"""
from __future__ import annotations

import os
import sys
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BITS_SET = (2, 4, 8)
# Final-layer activations / first-layer weight gradient, product vs reference flow, same weights,
# bit assignment and generator seed.  The aggregation sums in a different fp32 order (our CSR kernel
# vs cuSPARSE), which moves layer-1 activations by ~1e-6 relative; now and then that flips ONE
# stochastic-rounding decision of a later exchange by one level ((max-min)/(2^b-1) on one element of
# one halo row, i.e. up to a third of the row's range at 2 bits), so the MAX error is flip-dominated
# and grows with the number of quantised elements, while the MEAN error and the loss stay at fp32
# rounding level.  Measured on B200 (profiles/r02_parity.md): max 8.5e-6 .. 6.8e-4, mean <= 8.5e-6.
ACT_TOL = 5e-3          # max|a - b| / max|b|   (asserted by the small-graph tests)
ACT_MEAN_TOL = 1e-4     # sum|a - b| / sum|b|   (the bench line's `within_tolerance`, any size)
LOSS_TOL = 1e-4         # |loss_a - loss_b| / |loss_b|


def _windows(G: int, wg: int, n_windows: int) -> List[int]:
    """Start groups of the compared windows inside a segment of G byte-rows."""
    if wg <= 0 or G <= wg:
        return [0]
    starts = {0, G - wg}
    for i in range(1, max(n_windows - 1, 1)):
        starts.add(min(G - wg, (G * i) // max(n_windows - 1, 1)))
    return sorted(starts)


def _seeded_rows(n: int, F: int, seed: int, device) -> torch.Tensor:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x = torch.relu(torch.randn(n, F, generator=g, device=device))
    if n:
        x[::17] = 0.0                      # constant rows: scale = inf (all-zero gradient rows)
        x[5::29] *= 1.0e3                  # wide-range rows
        x[3::31] -= 0.75                   # negative minima
    return x.contiguous()


def exchange_parity(keys: Optional[Sequence[str]] = None, window_groups: int = 64, n_windows: int = 4,
                    base_seed: int = 4242) -> Dict:
    from oracle import oracle as O
    from adaqp_b200.communicator import Communicator as comm
    from adaqp_b200.helper import BitType
    from adaqp_b200.manager import GraphEngine as engine

    eng = engine.ctx
    rank, W = comm.get_rank(), comm.get_world_size()
    out = {"n_gpus": W, "keys": [], "bytes_compared": 0, "params_compared": 0, "halo_values_compared": 0,
           "fp32_values_compared": 0, "mismatches": 0,
           "distinct_gpus": len(set(comm.gather_all(str(torch.cuda.get_device_properties(comm.ctx.device).uuid))))}
    if W == 1 or comm.ctx.transport != "p2p":
        return out
    buf = comm.ctx.comm_buffer
    ex = buf.p2p
    dev = comm.ctx.device
    quant = eng.bit_type == BitType.QUANT
    if keys is None:
        L = ex.num_layers
        keys = (["forward0", "forward1", f"backward{L - 1}"] if quant else []) + ["test0"]
    mism = 0
    for ki, key in enumerate(keys):
        F = ex.dims[key]
        x = _seeded_rows(eng.num_inner, F, base_seed + 131 * rank + ki, dev)
        expect: Dict[int, Dict] = {p: {} for p in ex.send_peers}
        if key.startswith("test"):
            # fp32 exchange (comm.py:166-191 + op_util.py:168-170): first / last rows of every slice
            ex.post_send_fp(key, x)
            halo = ex.complete_recv_fp(key)
            torch.cuda.synchronize()
            comm.barrier()
            K = window_groups if window_groups > 0 else 1 << 60
            for p in ex.send_peers:
                lo, hi = ex.send_idx[p]
                sel = np.unique(np.concatenate([np.arange(lo, min(hi, lo + K)), np.arange(max(lo, hi - K), hi)]))
                rows = x[torch.from_numpy(ex.total_send_idx[sel]).to(dev)].cpu().numpy()
                expect[p] = {"sel": sel - lo, "rows": rows}
            gathered = comm.gather_all(expect)
            for src in ex.recv_peers:
                e = gathered[src][rank]
                dst = torch.from_numpy(ex.recv_idx[src][e["sel"]]).to(dev)
                got = halo[dst].cpu().numpy()
                mism += int((got.view(np.uint32) != e["rows"].view(np.uint32)).sum())
                out["fp32_values_compared"] += int(got.size)
            ex.release_fp(key)
            torch.cuda.synchronize()
            comm.barrier()
            out["keys"].append(key)
            continue
        seed, offset = base_seed + rank, 4096 * (ki + 1)
        ex.post_send_quant(key, x, seed, offset)
        ex.wait_flags_quant(key)
        halo = ex.complete_recv_quant(key)
        torch.cuda.synchronize()
        comm.barrier()
        # ---- sender side: oracle restatement of windows of what I sent
        rel = 0
        for p in ex.send_peers:
            lo, hi = ex.send_idx[p]
            for b in BITS_SET:
                ids = buf.send_original_idx_buffers[key][p].get(b)
                if ids is None:
                    continue
                ids = ids.cpu().numpy()
                wpt = 8 // b
                G = (ids.size + wpt - 1) // wpt
                wg = window_groups if window_groups > 0 else G
                for g0 in _windows(G, wg, n_windows):
                    g1 = min(G, g0 + wg)
                    rows = ids[g0 * wpt:min(ids.size, g1 * wpt)]
                    src = torch.from_numpy(ex.total_send_idx[lo + rows]).to(dev)
                    sub = x[src].cpu().numpy()
                    rmin, _rmax, scale = O.minmax_scale(sub, b)
                    packed = O.pack_at(sub, rmin, scale, b, seed, offset + rel, g0)
                    expect[p][(b, g0)] = (packed, O.to_bf16(scale), O.to_bf16(rmin))
                rel += ((F * wpt + 3) // 4) * 4        # one philox_engine_inputs per pack call
        assert rel == ex.quant_plans[key].philox_increment
        gathered = comm.gather_all(expect)
        # ---- receiver side: my slab and my halo rows against the senders' expectations
        for src in ex.recv_peers:
            q, prm = ex.recv_region(key, src)
            prm16 = prm.view(torch.int16)
            sizes = buf.recv_original_size_buffers[key][src]
            origs = buf.recv_original_idx_buffers[key][src]
            seg_off = prm_off = 0
            for b in BITS_SET:
                if b not in sizes:
                    continue
                qs, n = sizes[b]
                wpt = 8 // b
                orig = origs[b].cpu().numpy()
                for (bb, g0), (packed, sc16, mn16) in gathered[src][rank].items():
                    if bb != b:
                        continue
                    nrows = sc16.size
                    got_q = q[seg_off + g0 * F:seg_off + g0 * F + packed.size].cpu().numpy().view(np.uint8)
                    mism += int((got_q != packed).sum())
                    out["bytes_compared"] += int(packed.size)
                    p0 = prm_off + g0 * wpt
                    got_sc = prm16[0, p0:p0 + nrows].cpu().numpy().view(np.uint16)
                    got_mn = prm16[1, p0:p0 + nrows].cpu().numpy().view(np.uint16)
                    mism += int((got_sc != sc16).sum()) + int((got_mn != mn16).sum())
                    out["params_compared"] += 2 * nrows
                    want = O.unpack(packed, b, O.from_bf16(sc16), O.from_bf16(mn16), nrows, F)
                    dst = torch.from_numpy(ex.recv_idx[src][orig[g0 * wpt:g0 * wpt + nrows]]).to(dev)
                    got_h = halo[dst].cpu().numpy()
                    mism += int((got_h.view(np.uint32) != want.view(np.uint32)).sum())
                    out["halo_values_compared"] += int(want.size)
                seg_off += qs
                prm_off += n
        out["keys"].append(key)
        comm.barrier()
    ex.check_status()
    tot = torch.tensor([mism, out["bytes_compared"], out["params_compared"], out["halo_values_compared"],
                        out["fp32_values_compared"]], dtype=torch.int64)
    comm.all_reduce_sum(tot)
    out["mismatches"], out["bytes_compared"], out["params_compared"] = int(tot[0]), int(tot[1]), int(tot[2])
    out["halo_values_compared"], out["fp32_values_compared"] = int(tot[3]), int(tot[4])
    out["window_groups"] = window_groups
    return out


def current_assignment() -> Dict[str, Dict[int, torch.Tensor]]:
    """The bit assignment the product's CommBuffer currently holds, as the Assigner returned it."""
    from adaqp_b200.communicator import Communicator as comm
    from adaqp_b200.manager import GraphEngine as engine
    buf = comm.ctx.comm_buffer
    out = {}
    for key, per in buf.send_original_idx_buffers.items():
        out[key] = {}
        for p, groups in per.items():
            lo, hi = engine.ctx.send_idx[p]
            bits = torch.zeros(hi - lo, dtype=torch.int32)
            for b, ids in groups.items():
                bits[ids.cpu()] = b
            out[key][p] = bits
    return out


def activation_parity(trainer, seed: int = 1234, overlapped: bool = False) -> Dict:
    from oracle import build as obuild
    from oracle import ref_path
    from adaqp_b200.communicator import Communicator as comm
    from adaqp_b200.helper import BitType
    from adaqp_b200.manager import GraphEngine as engine
    from adaqp_b200.model.op_util import halo_exchange

    eng = engine.ctx
    dev = comm.ctx.device
    rank, W = comm.get_rank(), comm.get_world_size()
    cfg = trainer.config
    kind = cfg["runtime"]["model_name"]
    quant = eng.bit_type == BitType.QUANT and W > 1
    if kind == "sage" and cfg["model"]["aggregator_type"] != "mean":
        return {"unavailable": "reference-flow restatement covers the 'mean' aggregator only"}
    qc = None
    if quant:
        if not obuild.ref_available():
            return {"unavailable": "oracle/_ref/quant_cuda.so (reference kernels) not built"}
        qc = obuild.load_ref()
    dims = [cfg["data"]["num_feats"]] + [cfg["model"]["hidden_dim"]] * (cfg["model"]["num_layers"] - 1)
    # The reference's overlap (helper thread + side stream, ops.py:156-193) reorders execution, not
    # arithmetic -- and reads send_messages on the side stream without an event (ops.py:164 ->
    # op_util.py:101-110), which makes an overlapped run non-deterministic at small sizes.  The check
    # therefore runs the reference flow in its sequential form (full_graph_propagation, ops.py:132-154);
    # `overlapped=True` restates the helper-thread form for comparison.
    st = ref_path.RefState(eng.layout, dev, dims, quant, bool(overlapped) and eng.use_parallel and W > 1, qc, kind)
    if quant:
        st.update_quant(current_assignment())
    ref = ref_path.make_model(kind, dims, cfg["data"]["num_classes"], cfg["model"]["dropout_rate"]).to(dev)
    ref_path.load_product_state(ref, trainer.model.state_dict())
    multilabel = cfg["data"]["is_multilabel"]
    crit = torch.nn.BCEWithLogitsLoss(reduction="sum") if multilabel else torch.nn.CrossEntropyLoss(reduction="sum")
    feats, labels, mask = eng.feats, eng.labels, eng.train_mask
    n_train = torch.LongTensor([mask.numel()])
    comm.all_reduce_sum(n_train)
    res: Dict = {"seed": seed, "model": kind, "reference_flow": "overlapped" if st.parallel else "sequential"}

    # (1) quantised halo of forward0: identical inputs on both flows => bit-identical halos
    if quant:
        torch.cuda.manual_seed(seed)
        pend = halo_exchange(feats, "forward0", True)
        ours = pend.halo.clone()
        pend.release()
        torch.cuda.synchronize()
        comm.barrier()
        torch.cuda.manual_seed(seed)
        was = st.parallel
        st.parallel = False                       # all2all on the calling thread / current stream
        theirs = st.all2all(feats[st.total_send_idx], "forward0", True)
        st.parallel = was
        torch.cuda.synchronize()
        bad = torch.tensor([int((ours.view(torch.int32) != theirs.view(torch.int32)).sum())], dtype=torch.int64)
        comm.all_reduce_sum(bad)
        res["forward0_halo_vs_reference_kernels"] = {"values": int(ours.numel()), "mismatches": int(bad.item())}

    # (2) one training step from the same generator state
    def step(model, call):
        model.train()
        model.zero_grad(set_to_none=True)
        torch.cuda.manual_seed(seed)
        logits = call()
        loss = crit(logits[mask], labels[mask]) / n_train.item()
        loss.backward()
        torch.cuda.synchronize()
        comm.barrier()
        return logits.detach(), float(loss.item())

    lp, loss_p = step(trainer.model, lambda: trainer.model(eng.graph, feats))
    eng.timer.clear()
    lr, loss_r = step(ref, lambda: ref(st, feats))
    gp = dict(trainer.model.named_parameters())["convs.0.weight" if kind == "gcn" else "sages.0.fc_neigh.weight"].grad
    gr = (ref.w[0].grad if kind == "gcn" else ref.fc_neigh[0].weight.grad)
    v = torch.tensor([float((lp - lr).abs().max()), float(lr.abs().max()),
                      float((gp - gr).abs().max()), float(gr.abs().max())], dtype=torch.float64)
    comm.all_reduce_max(v)
    d = (lp - lr).abs()
    sums = torch.tensor([float(d.sum()), float(lr.abs().sum()), float((d > 1e-3 * float(v[1])).sum()), float(d.numel())],
                        dtype=torch.float64)
    comm.all_reduce_sum(sums)
    act_max = float(v[0] / max(float(v[1]), 1e-30))
    act_mean = float(sums[0] / max(float(sums[1]), 1e-30))
    res.update({"act_max_rel_err": act_max, "act_mean_rel_err": act_mean,
                "act_frac_elems_over_1e-3": float(sums[2] / max(float(sums[3]), 1.0)),
                "grad0_max_rel_err": float(v[2] / max(float(v[3]), 1e-30)),
                "loss_product": loss_p, "loss_reference_flow": loss_r,
                "loss_rel_diff": abs(loss_p - loss_r) / max(abs(loss_r), 1e-30)})
    res["tolerance"] = {"act_max": ACT_TOL, "act_mean": ACT_MEAN_TOL, "loss": LOSS_TOL}
    res["within_tolerance"] = bool(act_mean <= ACT_MEAN_TOL and res["loss_rel_diff"] <= LOSS_TOL)
    st.pool.close()
    if comm.ctx.comm_buffer.p2p is not None:
        comm.ctx.comm_buffer.p2p.check_status()
    return res


def parity_check(trainer, window_groups: int = 64) -> Dict:
    out = exchange_parity(window_groups=window_groups)
    try:
        out["activations"] = activation_parity(trainer)
        if "act_max_rel_err" in out["activations"]:
            out["act_max_rel_err"] = out["activations"]["act_max_rel_err"]
            h = out["activations"].get("forward0_halo_vs_reference_kernels")
            if h:
                out["mismatches"] += h["mismatches"]
    except Exception as e:                # the check must never take the bench line down with it
        out["activations"] = {"error": f"{type(e).__name__}: {e}"}
    return out
