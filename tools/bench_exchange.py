"""Loopback micro-benchmark of the fused exchange kernels on ONE GPU.

W ranks of an ogbn-products-shaped partitioning are simulated in one process (their slabs
address each other directly -- the same stores a peer GPU receives over NVLink, minus the
link), so the codec kernels can be timed and profiled (ncu) without a multi-GPU box.
Reports per kernel: event-timed duration, achieved GB/s against the algorithmic bytes of
SURVEY.md 8d (send: 4F+4 read + F*b/8+4 written per row; recv: F*b/8+4 read + 4F written).

    python tools/bench_exchange.py [--world 8] [--scale 0.25] [--reps 20] [--json out.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--scale", type=float, default=0.25)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--dataset", type=str, default="ogbn-products")
    ap.add_argument("--json", type=str, default=None)
    ap.add_argument("--only", type=str, default=None, help="key:bits, e.g. forward1:mixed")
    args = ap.parse_args()
    from adaqp_b200 import build
    build.build()
    from adaqp_b200.communicator.p2p import PeerExchange, update_quant_in_process, wire_in_process
    from adaqp_b200.manager.layout import prepare_all_in_process
    from adaqp_b200.manager.partition_synth import spec_from_config

    cfg = yaml.safe_load(open(os.path.join(ROOT, "adaqp_b200", "config", f"{args.dataset}.yaml")))
    spec = spec_from_config(cfg, args.world, args.scale)
    lays = prepare_all_in_process(spec)
    dev = torch.device("cuda:0")
    dims = [cfg["data"]["num_feats"], cfg["model"]["hidden_dim"], cfg["model"]["hidden_dim"]]
    exs = [PeerExchange(L.rank, args.world, dev, dims, L.send_idx, {p: torch.from_numpy(v) for p, v in L.recv_idx.items()},
                        torch.from_numpy(L.total_send_idx), L.n_halo, timeout_ns=10_000_000_000) for L in lays]
    wire_in_process(exs)
    peak = 6480.5
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    rng = np.random.RandomState(0)
    results = []
    L0 = lays[0]
    S0, R0 = int(L0.total_send_idx.size), int(L0.n_halo)
    for key, F in (("forward0", dims[0]), ("forward1", dims[1])):
        xs = [torch.relu(torch.randn(L.n_inner, F, device=dev)) for L in lays]
        for label, pick in (("2bit", [2]), ("4bit", [4]), ("8bit", [8]), ("mixed", [2, 4, 8])):
            if args.only and args.only != f"{key}:{label}":
                continue
            assign = [{key: {p: torch.from_numpy(np.array(pick, np.int32)[rng.randint(0, len(pick), hi - lo)])
                             for p, (lo, hi) in L.send_idx.items()}} for L in lays]
            update_quant_in_process(exs, assign)
            bits0 = np.concatenate([assign[0][key][p].numpy() for p in L0.send_idx])
            rbits0 = np.concatenate([assign[p][key][0].numpy() for p in L0.recv_idx])
            send_bytes = S0 * (4 * F + 4) + int((bits0.astype(np.int64) * F // 8 + 4).sum())
            recv_bytes = int((rbits0.astype(np.int64) * F // 8 + 4).sum()) + R0 * 4 * F
            t_send, t_recv = [], []
            for rep in range(args.reps + 3):
                evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                for e, x in zip(exs[1:], xs[1:]):
                    e.post_send_quant(key, x, 7, 0)
                evs[0].record()
                exs[0].post_send_quant(key, xs[0], 7, 0)
                evs[1].record()
                for e in exs[1:]:
                    e.complete_recv_quant(key)
                evs[2].record()
                exs[0].complete_recv_quant(key)
                evs[3].record()
                torch.cuda.synchronize()
                if rep >= 3:
                    t_send.append(evs[0].elapsed_time(evs[1]))
                    t_recv.append(evs[2].elapsed_time(evs[3]))
            ts, tr = float(np.median(t_send)), float(np.median(t_recv))
            results.append({"key": key, "F": F, "bits": label, "send_rows": S0, "recv_rows": R0,
                            "send_ms": ts, "recv_ms": tr, "send_GBps": send_bytes / ts / 1e6, "recv_GBps": recv_bytes / tr / 1e6,
                            "send_frac_hbm": send_bytes / ts / 1e6 / peak, "recv_frac_hbm": recv_bytes / tr / 1e6 / peak})
            print(json.dumps(results[-1]), flush=True)
        # fp32 exchange
        if args.only and args.only != f"{key}:fp32":
            continue
        t_fp = []
        for rep in range(args.reps + 3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for e, x in zip(exs[1:], xs[1:]):
                e.post_send_fp(key, x)
            a.record()
            exs[0].post_send_fp(key, xs[0])
            b.record()
            for e in exs:
                e.complete_recv_fp(key)
                e.release_fp(key)
            torch.cuda.synchronize()
            if rep >= 3:
                t_fp.append(a.elapsed_time(b))
        tf = float(np.median(t_fp))
        results.append({"key": key, "F": F, "bits": "fp32", "send_rows": S0, "send_ms": tf,
                        "send_GBps": S0 * 8 * F / tf / 1e6, "send_frac_hbm": S0 * 8 * F / tf / 1e6 / peak})
        print(json.dumps(results[-1]), flush=True)
    for e in exs:
        e.check_status()
        e.close()
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"world": args.world, "scale": args.scale, "peak_hbm_GBps": peak, "results": results}, f, indent=1)


if __name__ == "__main__":
    main()
