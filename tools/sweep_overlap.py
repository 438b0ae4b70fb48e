"""Sweep of the exchange kernels' grid caps against the overlapped aggregation, in ONE process per
rank (the partitions are generated once):  torchrun --nproc-per-node N tools/sweep_overlap.py

For every (exch_send_ctas, exch_recv_ctas) pair: warm-up epochs, then timed training epochs of the
bench workload; reports epoch ms (CUDA events, max over ranks), exposed comm ms, aggregation ms and
the exchange kernels' own ms per epoch.  0 = one resident wave over all SMs (round-1 behaviour)."""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="ogbn-products")
    ap.add_argument("--model_name", default="gcn")
    ap.add_argument("--mode", default="AdaQP")
    ap.add_argument("--assign_scheme", default="random")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--caps", default="0:0,96:96,64:64,48:48,32:32,64:32,32:64,16:16")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29531")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("LOCAL_RANK", "0")
    os.environ["ADAQP_SYNTH_SCALE"] = str(a.scale)
    os.environ.setdefault("ADAQP_SYNTHETIC", "1")
    os.environ.setdefault("ADAQP_SEED", "2024")
    import numpy as np
    import torch
    import torch.distributed as dist
    from argparse import Namespace
    from adaqp_b200 import build
    build.build()
    from adaqp_b200 import Trainer, _lib
    from adaqp_b200.communicator import Communicator as comm
    from adaqp_b200.manager import GraphEngine as engine
    from adaqp_b200.trainer.runtime_util import sync_model, sync_seed, train_for_one_epoch
    world = int(os.environ["WORLD_SIZE"])
    tr = Trainer(Namespace(dataset=a.dataset, num_parts=world, backend="gloo", init_method="env://", model_name=a.model_name,
                           mode=a.mode, assign_scheme=a.assign_scheme, logger_level="WARNING", exp_path="/tmp/adaqp_sweep_exp"))
    eng = engine.ctx
    cfg = tr.config
    sync_seed()
    tr.model.reset_parameters()
    sync_model(tr.model)
    opt = torch.optim.Adam(tr.model.parameters(), lr=cfg["runtime"]["learning_rate"])
    crit = torch.nn.BCEWithLogitsLoss(reduction="sum") if cfg["data"]["is_multilabel"] else torch.nn.CrossEntropyLoss(reduction="sum")
    n_train = torch.LongTensor([eng.train_mask.numel()])
    comm.all_reduce_sum(n_train)
    n_train = int(n_train.item())
    ep = [1]

    def step():
        ep[0] += 1
        return train_for_one_epoch(ep[0] + 1, eng.graph, tr.model, eng.feats, eng.labels, opt, crit, n_train, eng.train_mask)

    out = []
    for pair in a.caps.split(","):
        sc, rc = (int(x) for x in pair.split(":"))
        _lib.set_option("exch_send_ctas", sc)
        _lib.set_option("exch_recv_ctas", rc)
        for _ in range(3):
            step()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        exposed, agg, exch = [], [], []
        e0.record()
        for _ in range(a.steps):
            _, _, traced, _ = step()
            exposed.append(eng.last_exposed_comm_ms)
            agg.append(1e3 * (traced[3] + traced[4] + traced[5]))
            exch.append(1e3 * (traced[1] + traced[2]))
        e1.record()
        torch.cuda.synchronize()
        v = torch.tensor([e0.elapsed_time(e1) / a.steps, float(np.mean(exposed)), float(np.mean(agg)), float(np.mean(exch))],
                         dtype=torch.float64)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        rec = {"n_gpus": world, "exch_send_ctas": sc, "exch_recv_ctas": rc, "ms_per_epoch": float(v[0]),
               "exposed_comm_ms": float(v[1]), "aggregation_ms": float(v[2]), "exchange_kernels_ms": float(v[3])}
        out.append(rec)
        if comm.get_rank() == 0:
            print(json.dumps(rec), flush=True)
    if comm.ctx.comm_buffer.p2p is not None:
        comm.ctx.comm_buffer.p2p.check_status()
    if comm.get_rank() == 0 and a.json:
        with open(a.json, "w") as f:
            json.dump({"dataset": a.dataset, "model": a.model_name, "mode": a.mode, "scale": a.scale, "results": out}, f, indent=1)
    comm.ctx.delete_buffer()


if __name__ == "__main__":
    main()
