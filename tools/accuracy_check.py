"""Validation accuracy of the quantised exchange vs fp32 (north star: within 0.3 % of Vanilla).

Trains the same synthetic, learnable ogbn-products-shaped task (labels = community class,
features = class centroid + noise, homophilous edges) with --mode Vanilla and --mode AdaQP
for the same number of epochs and seed, one process per GPU, and prints best / final
validation accuracy of each.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/accuracy_check.py --epochs 100 --scale 0.05
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


KEEP = []   # keep Trainer/Communicator objects alive: Communicator.__del__ tears the process group down


def run(mode, scheme, args, rank, world):
    from argparse import Namespace
    import torch
    from adaqp_b200 import Trainer
    from adaqp_b200.manager import GraphEngine as engine
    t = Trainer(Namespace(dataset=args.dataset, num_parts=world, backend="gloo", init_method="env://",
                          model_name=args.model_name, mode=mode, assign_scheme=scheme, logger_level="WARNING",
                          num_epoches=args.epochs, exp_path="/tmp/adaqp_acc_exp"))
    KEEP.append(t)
    if scheme == "adaptive":
        t.assigner.assign_cycle = args.assign_cycle
    t.train()
    m = engine.ctx.recorder.epoches_metrics[:args.epochs]
    best = int(m[:, 1].argmax())
    return {"mode": mode, "scheme": scheme, "best_val": float(m[best, 1]), "test_at_best": float(m[best, 2]),
            "final_val": float(m[-1, 1]), "final_train": float(m[-1, 0])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=100)
    ap.add_argument("--scale", type=float, default=0.05)
    ap.add_argument("--dataset", type=str, default="ogbn-products")
    ap.add_argument("--model_name", type=str, default="gcn")
    ap.add_argument("--assign_cycle", type=int, default=25)
    ap.add_argument("--json", type=str, default=None)
    ap.add_argument("--feature-signal", dest="signal", type=float, default=0.12, help="class-centroid strength of the synthetic features")
    args = ap.parse_args()
    os.environ["ADAQP_SYNTH_SCALE"] = str(args.scale)
    os.environ.setdefault("ADAQP_SYNTHETIC", "1")
    os.environ["ADAQP_SYNTH_SIGNAL"] = str(args.signal)
    os.environ.setdefault("ADAQP_SEED", "123")
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    import __graft_entry__ as entry
    if rank == 0:
        entry.build()
    out = []
    for mode, scheme in (("Vanilla", "uniform"), ("AdaQP", "uniform"), ("AdaQP", "random"), ("AdaQP", "adaptive")):
        out.append(run(mode, scheme, args, rank, world))
        if rank == 0:
            print(json.dumps(out[-1]), flush=True)
    if rank == 0:
        base = out[0]["best_val"]
        summary = {"world": world, "epochs": args.epochs, "scale": args.scale, "feature_signal": args.signal, "model": args.model_name, "runs": out,
                   "delta_best_val_vs_vanilla_pct": {f"{o['mode']}/{o['scheme']}": 100 * (o["best_val"] - base) for o in out[1:]}}
        print(json.dumps(summary), flush=True)
        if args.json:
            with open(args.json, "w") as f:
                json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
