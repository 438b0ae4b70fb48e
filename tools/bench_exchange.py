"""Micro-benchmark of the fused exchange kernels driven from ONE process.

W ranks of an ogbn-products-shaped partitioning live in one process.  `--devices 1` (default):
all ranks on cuda:0, their slabs address each other directly -- the stores a peer GPU would
receive, minus the link (codec-only loopback).  `--devices D` (D >= 2): rank r lives on
cuda:(r % D) with peer access enabled between the devices, so rank 0's send kernel stores
into slabs on the OTHER GPUs over NVLink/NVSwitch and its receive kernel reads what peers
stored: a single process that ncu can profile with the link counters
(nvltx__bytes / nvlrx__bytes / syslts__t_sectors_aperture_peer*; profiles/README.md).
Reports per kernel: event-timed duration of the kernel ALONE on its stream, achieved GB/s
against the algorithmic bytes of SURVEY.md 8d (send: 4F+4 read from HBM, F*b/8+4 stored
to the peer per row; recv: F*b/8+4 read + 4F written per row).

    python tools/bench_exchange.py [--world 8] [--devices 1] [--scale 0.25] [--reps 20] [--json out.json]
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
    ap.add_argument("--devices", type=int, default=1, help="GPUs the ranks are spread over (rank r on cuda:(r %% D))")
    args = ap.parse_args()
    from adaqp_b200 import build
    build.build()
    from adaqp_b200 import _lib
    from adaqp_b200.communicator.p2p import PeerExchange, update_quant_in_process, wire_in_process
    from adaqp_b200.manager.layout import prepare_all_in_process
    from adaqp_b200.manager.partition_synth import spec_from_config

    cfg = yaml.safe_load(open(os.path.join(ROOT, "adaqp_b200", "config", f"{args.dataset}.yaml")))
    spec = spec_from_config(cfg, args.world, args.scale)
    lays = prepare_all_in_process(spec)
    D = max(1, min(args.devices, torch.cuda.device_count()))
    devs = [torch.device(f"cuda:{r % D}") for r in range(args.world)]
    dev = devs[0]
    for a in range(D):                      # one process drives D GPUs: map every peer's memory
        for b in range(D):
            if a != b:
                with torch.cuda.device(a):
                    _lib.check(_lib.load().adaqp_enable_peer_access(b), "adaqp_enable_peer_access")
    dims = [cfg["data"]["num_feats"], cfg["model"]["hidden_dim"], cfg["model"]["hidden_dim"]]
    exs = [PeerExchange(L.rank, args.world, devs[L.rank], dims, L.send_idx, {p: torch.from_numpy(v) for p, v in L.recv_idx.items()},
                        torch.from_numpy(L.total_send_idx), L.n_halo, timeout_ns=10_000_000_000) for L in lays]
    wire_in_process(exs)

    class on:                               # launches go to the current stream of the rank's device
        def __init__(self, e):
            self.ctx = torch.cuda.device(e.device)

        def __enter__(self):
            return self.ctx.__enter__()

        def __exit__(self, *a):
            return self.ctx.__exit__(*a)

    def sync_all():
        for d in range(D):
            torch.cuda.synchronize(d)
    nvlink_peak = 770.0                     # GB/s per direction, the figure the bench line quotes (B200_PROFILING.md)
    peak = 6480.5
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    rng = np.random.RandomState(0)
    results = []
    L0 = lays[0]
    S0, R0 = int(L0.total_send_idx.size), int(L0.n_halo)
    for key, F in (("forward0", dims[0]), ("forward1", dims[1])):
        xs = [torch.relu(torch.randn(L.n_inner, F, device=devs[L.rank])) for L in lays]
        for label, pick in (("2bit", [2]), ("4bit", [4]), ("8bit", [8]), ("mixed", [2, 4, 8])):
            if args.only and args.only != f"{key}:{label}":
                continue
            assign = [{key: {p: torch.from_numpy(np.array(pick, np.int32)[rng.randint(0, len(pick), hi - lo)])
                             for p, (lo, hi) in L.send_idx.items()}} for L in lays]
            update_quant_in_process(exs, assign)
            bits0 = np.concatenate([assign[0][key][p].numpy() for p in L0.send_idx])
            rbits0 = np.concatenate([assign[p][key][0].numpy() for p in L0.recv_idx])
            wire0 = int((bits0.astype(np.int64) * F // 8 + 4).sum())          # bytes rank 0 stores into its peers
            send_bytes = S0 * (4 * F + 4) + wire0
            recv_bytes = int((rbits0.astype(np.int64) * F // 8 + 4).sum()) + R0 * 4 * F
            # rows of rank 0 whose destination lives on another GPU (all of them when D == world)
            off_dev = sum(hi - lo for p, (lo, hi) in L0.send_idx.items() if devs[p] != devs[0])
            link_bytes = int(wire0 * off_dev / max(S0, 1))
            t_send, t_recv = [], []
            for rep in range(args.reps + 3):
                evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                for e, x in zip(exs[1:], xs[1:]):
                    with on(e):
                        e.post_send_quant(key, x, 7, 0)
                sync_all()                  # the peers' payloads have landed: time rank 0's kernels alone
                with on(exs[0]):
                    evs[0].record()
                    exs[0].post_send_quant(key, xs[0], 7, 0)
                    evs[1].record()
                for e in exs[1:]:
                    with on(e):
                        e.complete_recv_quant(key)
                with on(exs[0]):
                    evs[2].record()
                    exs[0].complete_recv_quant(key)
                    evs[3].record()
                sync_all()
                if rep >= 3:
                    t_send.append(evs[0].elapsed_time(evs[1]))
                    t_recv.append(evs[2].elapsed_time(evs[3]))
            ts, tr = float(np.median(t_send)), float(np.median(t_recv))
            results.append({"key": key, "F": F, "bits": label, "send_rows": S0, "recv_rows": R0, "devices": D,
                            "send_ms": ts, "recv_ms": tr, "send_GBps": send_bytes / ts / 1e6, "recv_GBps": recv_bytes / tr / 1e6,
                            "send_frac_hbm": send_bytes / ts / 1e6 / peak, "recv_frac_hbm": recv_bytes / tr / 1e6 / peak,
                            "send_link_bytes": link_bytes, "send_link_GBps": link_bytes / ts / 1e6,
                            "send_frac_nvlink": link_bytes / ts / 1e6 / nvlink_peak,
                            "exch_send_ctas": _lib.get_option("exch_send_ctas"), "exch_recv_ctas": _lib.get_option("exch_recv_ctas")})
            print(json.dumps(results[-1]), flush=True)
        # fp32 exchange
        if args.only and args.only != f"{key}:fp32":
            continue
        t_fp = []
        for rep in range(args.reps + 3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for e, x in zip(exs[1:], xs[1:]):
                with on(e):
                    e.post_send_fp(key, x)
            sync_all()
            with on(exs[0]):
                a.record()
                exs[0].post_send_fp(key, xs[0])
                b.record()
            for e in exs:
                with on(e):
                    e.complete_recv_fp(key)
                    e.release_fp(key)
            sync_all()
            if rep >= 3:
                t_fp.append(a.elapsed_time(b))
        tf = float(np.median(t_fp))
        off_dev = sum(hi - lo for p, (lo, hi) in L0.send_idx.items() if devs[p] != devs[0])
        results.append({"key": key, "F": F, "bits": "fp32", "send_rows": S0, "send_ms": tf, "devices": D,
                        "send_GBps": S0 * 8 * F / tf / 1e6, "send_frac_hbm": S0 * 8 * F / tf / 1e6 / peak,
                        "send_link_bytes": off_dev * 4 * F, "send_link_GBps": off_dev * 4 * F / tf / 1e6,
                        "send_frac_nvlink": off_dev * 4 * F / tf / 1e6 / nvlink_peak})
        print(json.dumps(results[-1]), flush=True)
    sync_all()
    for e in exs:
        with on(e):
            e.check_status()
            e.close()
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"world": args.world, "scale": args.scale, "peak_hbm_GBps": peak, "results": results}, f, indent=1)


if __name__ == "__main__":
    main()
