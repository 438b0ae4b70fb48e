"""bench.py -- epochs/sec and exposed comm ms of full-graph GCN training on ogbn-products-shaped
synthetic partitions (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench.py --gpus N ...

A step = one training epoch = the timed region of train_for_one_epoch
(AdaQP/trainer/runtime_util.py:98-111: forward, backward, gradient all-reduce, optimizer
step; the per-epoch evaluation is outside, as in the reference).  The graph is fixed, so
scaling over N is STRONG.  Prints ONE JSON line on rank 0.

`--impl reference` runs the reference's own flow restated in oracle/ref_path.py (gloo ring
through pinned host buffers, per-(peer, bit) Python loops around the reference-built
quant_cuda of oracle/_ref, torch.sparse aggregation standing in for DGL) on the same
partitions; it executes oracle/ code by design (test infrastructure), never the product.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    p.add_argument("--dataset", type=str, default="ogbn-products")
    p.add_argument("--model_name", type=str, default="gcn")
    p.add_argument("--mode", type=str, default="AdaQP")
    p.add_argument("--assign_scheme", type=str, default="random")
    p.add_argument("--assign_bits", type=int, default=None, help="uniform bit-width (overrides the yaml's assign_bits)")
    p.add_argument("--scale", type=float, default=float(os.environ.get("ADAQP_SYNTH_SCALE", "1.0")))
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-verify", action="store_true",
                   help="skip the parity_check leg (cross-GPU exchange vs the CPU oracle, one training step vs the "
                        "reference flow); it runs OUTSIDE the timed region")
    p.add_argument("--verify-window", type=int, default=64, help="byte-rows per compared window (<=0: everything)")
    return p.parse_args()


def workload_config(args, cfg, world, dims):
    """`config` of the JSON line -- built by ONE function for both arms so the dicts are identical."""
    return {"workload": f"{args.dataset}-shape {args.model_name} 3x256 full-graph training epoch, {world} partition(s), "
                        f"mode {args.mode}, bits {args.assign_scheme}{{2,4,8}}",
            "nodes": int(cfg["synthetic"]["num_nodes"] * args.scale), "edges": int(cfg["synthetic"]["num_edges"] * args.scale),
            "layer_dims": dims, "parallelism": f"graph-partition x{world}", "l2": "inputs >> L2 (no flush needed)"}


def setup_env(args):
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29517")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("LOCAL_RANK", "0")
    os.environ["ADAQP_SYNTH_SCALE"] = str(args.scale)
    os.environ.setdefault("ADAQP_SYNTHETIC", "1")      # datasets / DGL partitions are not in the image: synthetic shape
    os.environ.setdefault("ADAQP_SEED", "2024")
    world = int(os.environ["WORLD_SIZE"])
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    return int(os.environ["RANK"]), world


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        queries = [self.Q, self.Q.replace("clocks_event_reasons", "clocks_throttle_reasons")]
        while not self.stop_flag.is_set():
            for q in list(queries):
                try:
                    r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                        "-i", str(self.index)], capture_output=True, text=True, timeout=5)
                    out = r.stdout.strip()
                    if r.returncode == 0 and out and out.split(",")[0].strip().isdigit():
                        self.samples.append([x.strip() for x in out.split(",")])
                        queries = [q]
                        break
                except Exception:
                    pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def spmm_algorithmic_bytes(eng, dims, use_parallel):
    """Compulsory bytes of the aggregation launches of one training epoch:
    4*nnz (int32 indices) + 8*(rows+1) (int64 indptr) + 4*F*(n_src + rows) + 4*(n_src + rows)."""
    import numpy as np
    L = eng.layout
    passes = [dims[0], dims[1], dims[2], dims[2], dims[1]]     # fwd0, fwd1, fwd2, bwd2, bwd1
    ranges = [(0, L.n_central, L.n_inner), (L.n_central, L.n_inner, L.n_inner + L.n_halo)] if use_parallel \
        else [(0, L.n_inner, L.n_inner + L.n_halo)]
    total, launches = 0, 0
    for F in passes:
        for lo, hi, n_src in ranges:
            if hi == lo:
                continue
            nnz = int(L.indptr[hi] - L.indptr[lo])
            rows = hi - lo
            total += 4 * nnz + 8 * (rows + 1) + 4 * F * (n_src + rows) + 4 * (n_src + rows)
            launches += 1
    return total, launches


NVLINK_PEAK_GBPS = 770.0      # measured peer copy per direction, B200_PROFILING.md (nominal 900)


def exchange_stats(ex, eng, traced_all):
    """Rank 0's boundary traffic per training epoch and the event-timed duration of the
    exchange region (send + flag wait + receive, side stream)."""
    import numpy as np
    recv_bytes = rows = 0
    for key, plan in ex.quant_plans.items():
        for p, (qbytes, n) in plan.wire.items():
            recv_bytes += qbytes + 4 * n
            rows += n
    fp_bytes = sum(4 * ex.dims[k] * ex.num_remote for k in ex.keys if not k.startswith("test"))
    exch_ms = 1e3 * float(np.mean([t[1] + t[2] for t in traced_all]))      # comm + quant buckets
    b = recv_bytes if recv_bytes else fp_bytes
    return {"halo_rows_per_exchange": ex.num_remote, "send_rows_per_exchange": int(eng.total_send_idx.numel()),
            "bytes_in_per_epoch": int(b), "fp32_equivalent_bytes_per_epoch": int(fp_bytes),
            "exchange_region_ms_per_epoch": exch_ms}


def dram_rate(traffic_per_launch, launches, agg_ms, hbm_peak):
    """Measured DRAM rate of the aggregation launches: ncu dram bytes per launch (profiles/spmm_traffic.json) x the
    launches of an epoch / their event-timed duration, next to the copy peak.  None where no capture exists."""
    if not traffic_per_launch or not agg_ms or not launches:
        return None
    gbps = float(traffic_per_launch) * int(launches) / (float(agg_ms) * 1e-3) / 1e9
    return {"GBps": gbps, "frac_of_peak": gbps / float(hbm_peak),
            "note": "ncu DRAM bytes (read + write) of the epoch's aggregation launches / their event-timed duration: what the kernel "
                    "actually moves, mostly L2 misses this graph makes unavoidable (profiles/r02_spmm.md)"}


def exchange_roofline(ex, eng, times, epochs, hbm_peak):
    """Second roofline block: the send and receive kernels of the exchange, each timed ALONE with CUDA
    events on the stream it runs on (rank 0, sums over the timed epochs), against both bounds.
    Algorithmic bytes: send reads 4F+4 (+8 index) per boundary row from HBM and stores
    F*b/8+4 to the peer over NVLink; receive reads F*b/8+4 per halo row and writes 4F."""
    S = int(eng.total_send_idx.numel())
    send_hbm = send_link = recv_hbm = 0
    if ex.quant_plans:
        for key, plan in ex.quant_plans.items():
            F = ex.dims[key]
            wire_in = sum(q + 4 * n for q, n in plan.wire.values())
            send_hbm += S * (4 * F + 12)
            recv_hbm += wire_in + 4 * F * ex.num_remote
            send_link += plan.n_send * F + 4 * S     # F packed bytes per byte-row + bf16 (scale, min) per row
        send_hbm += send_link
    else:
        for key in ex.keys:
            if key.startswith("test"):
                continue
            F = ex.dims[key]
            send_hbm += S * (8 * F + 16)
            send_link += S * 4 * F
    out = {"rank": 0, "epochs": epochs, "launches": times["launches"]}
    if times["send"] > 0:
        ms = times["send"] / epochs
        out["send"] = {"kernel": "send_quant_kernel" if ex.quant_plans else "send_fp32_kernel", "ms_per_epoch": ms,
                       "hbm_bytes_per_epoch": int(send_hbm), "hbm_GBps": send_hbm / ms / 1e6, "hbm_frac": send_hbm / ms / 1e6 / hbm_peak,
                       "nvlink_bytes_per_epoch": int(send_link), "nvlink_GBps": send_link / ms / 1e6,
                       "nvlink_frac": send_link / ms / 1e6 / NVLINK_PEAK_GBPS}
    if times["recv"] > 0:
        ms = times["recv"] / epochs
        out["recv"] = {"kernel": "recv_quant_kernel", "ms_per_epoch": ms, "hbm_bytes_per_epoch": int(recv_hbm),
                       "hbm_GBps": recv_hbm / ms / 1e6, "hbm_frac": recv_hbm / ms / 1e6 / hbm_peak}
    out["peaks"] = {"hbm_GBps": hbm_peak, "nvlink_GBps_per_direction": NVLINK_PEAK_GBPS}
    out["note"] = ("parity-mode Philox (one Philox4x32-10 block per packed byte, mandated by curand_init(seed, k, offset)) makes the "
                   "send kernel integer-issue bound, not HBM/NVLink bound; both kernels overlap the aggregation")
    return out


def cpu_baseline_port(eng, dims, seconds_budget=15.0):
    """Oracle (C port) of the hot path on the host cores: the five aggregations of one epoch over a bounded,
    evenly spaced sample of destination rows (sized from a probe so that the leg takes ~`seconds_budget`
    seconds), extrapolated to the whole epoch by rows."""
    import numpy as np
    from oracle import oracle as O
    L = eng.layout
    n = L.n_inner
    passes = [dims[0], dims[1], dims[2], dims[2], dims[1]]
    rng = np.random.default_rng(0)
    n_src = L.n_inner + L.n_halo
    xs = {F: rng.standard_normal((n_src, F), dtype=np.float32) for F in set(passes)}
    pre = np.ones(n_src, np.float32)

    def run(sample):
        rows = np.linspace(0, n - 1, sample).astype(np.int64)
        ip = np.concatenate([[0], np.cumsum(np.diff(L.indptr)[rows])]).astype(np.int64)
        ix = np.concatenate([L.indices[L.indptr[r]:L.indptr[r + 1]] for r in rows]).astype(np.int64)
        t0 = time.time()
        for F in passes:
            O.aggregate(ip, ix, xs[F], pre=pre, post=pre[:sample])
        return time.time() - t0

    probe = max(1, min(n, 4096))
    t_probe = run(probe)
    sample = int(max(probe, min(n, probe * seconds_budget / max(t_probe, 1e-6))))
    dt = run(sample) if sample > probe else t_probe
    t_epoch = dt * (n / sample)
    return {"value": 1.0 / t_epoch, "unit": "epochs/s", "cores": 1, "kind": "port", "seconds": dt,
            "sample": f"oracle_aggregate (C, 1 thread) over {sample} of {n} destination rows x 5 passes ({dt:.1f} s), "
                      f"extrapolated by rows; aggregation only (no exchange at N=1, no dense GEMM)"}


def run_ours(args, rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from argparse import Namespace
    import __graft_entry__ as entry
    entry.build()
    from adaqp_b200 import Trainer
    from adaqp_b200.communicator import Communicator as comm
    from adaqp_b200.manager import GraphEngine as engine
    from adaqp_b200.trainer.runtime_util import sync_model, sync_seed, train_for_one_epoch

    targs = Namespace(dataset=args.dataset, num_parts=world, backend="gloo", init_method="env://",
                      model_name=args.model_name, mode=args.mode, assign_scheme=args.assign_scheme,
                      logger_level="WARNING", exp_path="/tmp/adaqp_bench_exp", assign_bits=getattr(args, "assign_bits", None))
    tr = Trainer(targs)
    eng = engine.ctx
    dev = comm.ctx.device
    cfg = tr.config
    dims = [cfg["data"]["num_feats"]] + [cfg["model"]["hidden_dim"]] * (cfg["model"]["num_layers"] - 1)
    sync_seed()
    tr.model.reset_parameters()
    sync_model(tr.model)
    opt = torch.optim.Adam(tr.model.parameters(), lr=cfg["runtime"]["learning_rate"])
    crit = torch.nn.BCEWithLogitsLoss(reduction="sum") if cfg["data"]["is_multilabel"] else torch.nn.CrossEntropyLoss(reduction="sum")
    n_train = torch.LongTensor([eng.train_mask.numel()])
    comm.all_reduce_sum(n_train)
    n_train = int(n_train.item())
    state = {"epoch": 0}

    def step(feats, labels):
        state["epoch"] += 1
        return train_for_one_epoch(state["epoch"] + 1, eng.graph, tr.model, feats, labels, opt, crit, n_train, eng.train_mask)

    def timed(nsteps, fn):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        extras = [fn() for _ in range(nsteps)]
        b.record()
        torch.cuda.synchronize()
        dist.barrier()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), extras

    sampler = ClockSampler(dev.index or 0)      # samples under load from the warm-up on
    sampler.start()
    adaptive = args.assign_scheme == "adaptive" and tr.assigner.is_tracing
    if adaptive:
        # adaptive re-assignment (runtime_util.py:86-93) happens when epoch % assign_cycle == 1: trace two
        # epochs at the uniform warm-up width, let the third warm-up epoch solve + re-allocate (untimed),
        # then freeze the assignment for the timed region
        tr.assigner.assign_cycle = 3
    for _ in range(max(args.warmup, 3)):
        step(eng.feats, eng.labels)
    if adaptive:
        tr.assigner.assign_cycle = 10 ** 9
        for _ in range(2):
            step(eng.feats, eng.labels)
    exposed, agg_s = [], []

    def dev_step():
        _, _, traced, _ = step(eng.feats, eng.labels)
        exposed.append(eng.last_exposed_comm_ms)
        agg_s.append(traced[3] + traced[4] + traced[5])
        return traced

    p2p = comm.ctx.comm_buffer.p2p
    if p2p is not None:
        p2p.profile = True
    from adaqp_b200 import dense, fused
    gemm_before = {**dense.LAUNCHES, **fused.LAUNCHES}
    ms, traced_all = timed(args.steps, dev_step)
    gemm_launches = {k: {**dense.LAUNCHES, **fused.LAUNCHES}[k] - gemm_before[k] for k in gemm_before}
    exch_times = None
    if p2p is not None:
        exch_times = p2p.kernel_times_ms()
        p2p.profile = False
    # ---- end-to-end: host inputs -> device every step, loss read back
    feats_host = eng.feats.cpu().pin_memory()
    labels_host = eng.labels.cpu().pin_memory()
    # Input pipeline as a user would run it: the pinned-host -> device copy of step i+1 is issued on a
    # copy stream while step i computes (two device buffers), so every timed step still pays for one
    # full H2D copy of its inputs and one D2H read of its loss, but the copy hides behind the epoch.
    copy_stream = torch.cuda.Stream()
    bufs = [(torch.empty_like(eng.feats), torch.empty_like(eng.labels)) for _ in range(2)]
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"i": 0}

    def prefetch(i):
        f, l = bufs[i % 2]
        copy_stream.wait_stream(torch.cuda.current_stream())   # the buffer's previous reader (step i-2) is done
        with torch.cuda.stream(copy_stream):
            f.copy_(feats_host, non_blocking=True)
            l.copy_(labels_host, non_blocking=True)
            copied[i % 2].record(copy_stream)

    prefetch(0)

    def e2e_step():
        i = e2e_state["i"]
        e2e_state["i"] += 1
        torch.cuda.current_stream().wait_event(copied[i % 2])
        prefetch(i + 1)
        f, l = bufs[i % 2]
        _, loss, _, _ = step(f, l)
        return float(loss.item())

    e2e_step()
    ms_e2e, losses = timed(args.steps, e2e_step)
    sampler.stop_flag.set()
    torch.cuda.synchronize()
    if comm.ctx.comm_buffer.p2p is not None:
        comm.ctx.comm_buffer.p2p.check_status()
    # ---- roofline of the dominant kernel (CSR SpMM) from the live event timings of the timed region
    alg_bytes, launches = spmm_algorithmic_bytes(eng, dims, eng.use_parallel)
    agg_ms = 1e3 * float(np.mean(agg_s))
    gl = torch.tensor([alg_bytes, agg_ms, float(np.mean(exposed))], dtype=torch.float64)
    dist.all_reduce(gl, op=dist.ReduceOp.MAX)
    peaks, peak_kind = measured_peaks()
    achieved = alg_bytes / (agg_ms * 1e-3) / 1e9
    # dram__bytes_read + dram__bytes_write per launch from an `ncu --set full` capture of THIS workload at THIS N
    # (profiles/spmm_traffic.json, keyed by n_gpus; per-shape figures next to it); null where none was taken
    traffic = traffic_detail = None
    tpath = os.path.join(ROOT, "profiles", "spmm_traffic.json")
    if os.path.exists(tpath) and args.dataset == "ogbn-products" and args.scale == 1.0:
        ent = json.load(open(tpath)).get("by_n_gpus", {}).get(str(world))
        if ent:
            traffic, traffic_detail = ent.get("dram_bytes_per_launch_mean"), ent.get("per_shape")
    n_exch = 5 if world > 1 else 0
    # our kernels per epoch: aggregation launches (the marginal rows take two passes -- local then halo
    # segment -- unless ADAQP_MARGINAL_SPLIT=0) + per exchange: send, flag wait, receive (quant) or
    # send, flag wait, ack (fp32)
    split_extra = 5 if (eng.use_parallel and eng.num_marginal > 0 and os.environ.get("ADAQP_MARGINAL_SPLIT", "1") != "0") else 0
    launches_per_epoch = launches + split_extra + n_exch * 3 + sum(gemm_launches.values()) // max(args.steps, 1)
    out = {
        "metric": "epochs_per_sec", "value": args.steps / (ms / 1e3), "unit": "epochs/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "exposed_comm_ms": float(gl[2].item()),
        "config": workload_config(args, cfg, world, dims),
        "rank0": {"n_inner": eng.num_inner, "n_central": eng.num_central, "n_halo": eng.num_remove,
                  "nnz": int(eng.layout.indptr[-1])},
        "clocks": sampler.summary(),
        "e2e": {"value": args.steps / (ms_e2e / 1e3), "unit": "epochs/s",
                "h2d_bytes_per_step": int(feats_host.numel() * 4 + labels_host.numel() * labels_host.element_size()),
                "d2h_bytes_per_step": 4,
                "api": "train_for_one_epoch on inputs copied from pinned host memory every step (double-buffered prefetch on a copy stream), loss.item()"},
        "gpu_launches": int(launches_per_epoch * args.steps),
        "roofline": {"kernel": "spmm_csr_kernel", "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "traffic_per_shape": traffic_detail, "peak_kind": f"of {peak_kind}",
                     "algorithmic_bytes_per_epoch": alg_bytes, "launches_per_epoch": launches, "spmm_ms_per_epoch": agg_ms,
                     "dram_rate": dram_rate(traffic, launches, agg_ms, peaks["hbm_gbs"]),
                     "gather_bound": {"note": "no-reuse bound 4*F*nnz: what a random gather must move when the source matrix exceeds L2",
                                      "achieved_GBps": sum(4 * F * int(eng.layout.indptr[-1]) for F in [dims[0], dims[1], dims[2], dims[2], dims[1]]) / (agg_ms * 1e-3) / 1e9}},
        "dense_gemm": {"kernels_per_epoch": {k: v // max(args.steps, 1) for k, v in gemm_launches.items()},
                       "arithmetic": "tcgen05 kind::tf32 x3 (error-compensated split, fp32 accumulate)" if dense.enabled() else "torch.matmul fp32",
                       "layer_norm_relu": "fused (csrc/norm.cu)" if fused.enabled() else "torch"},
        "final_loss": losses[-1],
        "exchange": exchange_stats(comm.ctx.comm_buffer.p2p, eng, traced_all) if world > 1 else None,
        "roofline_exchange": exchange_roofline(comm.ctx.comm_buffer.p2p, eng, exch_times, args.steps, peaks["hbm_gbs"]) if world > 1 else None,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_port(eng, dims)
    hook = getattr(args, "before_teardown", None)     # tools/run_configs.py: hand the layout / assignment to the reference arm
    if hook is not None:
        hook(tr)
    if not args.no_verify:
        # outside the timed region: what the peers wrote into my slab over NVLink vs the CPU oracle
        # (bit-exact), and one training step vs the reference flow around the reference's own kernels
        from tools import parity_check
        out["parity_check"] = parity_check.parity_check(tr, window_groups=args.verify_window)
    comm.ctx.delete_buffer()
    if rank == 0:
        print(json.dumps(out), flush=True)


def run_reference(args, rank, world):
    import yaml
    from oracle import ref_path
    cfg = yaml.safe_load(open(os.path.join(ROOT, "adaqp_b200", "config", f"{args.dataset}.yaml")))
    dims = [cfg["data"]["num_feats"]] + [cfg["model"]["hidden_dim"]] * (cfg["model"]["num_layers"] - 1)
    out = ref_path.bench(args, rank, world)
    if "config" in out:
        out["config"] = workload_config(args, cfg, world, dims)
    if rank == 0:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    a = parse()
    r, w = setup_env(a)
    if a.impl == "reference":
        run_reference(a, r, w)
    else:
        run_ours(a, r, w)
