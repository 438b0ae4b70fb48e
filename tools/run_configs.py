"""Run several BASELINE.json configurations -- product arm, then the reference-flow arm on the SAME
prepared partitions and bit assignment -- inside ONE process group (imports, CUDA context and NCCL
are paid once; GPU-minutes are the scarce resource).

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_configs.py \
        --configs products-sage-adaptive,yelp-gcn-random,amazon-sage-random --steps 5 --out profiles/bench/r02

    python tools/run_configs.py --spawn 2 --scale 0.01 ...     # self-launch W ranks (rank r on cuda:(r % #GPUs))

Each arm is exactly bench.py's code path (bench.run_ours / oracle.ref_path.bench): one JSON line per
arm, also written to <out>_<config>_{ours,reference}.json.  A failing configuration is reported and
the next one still runs."""
from __future__ import annotations

import argparse
import json
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# name -> bench arguments (BASELINE.json `configs`)
CONFIGS = {
    "products-gcn-random": dict(dataset="ogbn-products", model_name="gcn", mode="AdaQP", assign_scheme="random"),
    "reddit-gcn-uniform4": dict(dataset="reddit", model_name="gcn", mode="AdaQP", assign_scheme="uniform", assign_bits=4),
    "products-sage-adaptive": dict(dataset="ogbn-products", model_name="sage", mode="AdaQP", assign_scheme="adaptive"),
    "yelp-gcn-random": dict(dataset="yelp", model_name="gcn", mode="AdaQP", assign_scheme="random"),
    "yelp-gcn-adaptive": dict(dataset="yelp", model_name="gcn", mode="AdaQP", assign_scheme="adaptive"),
    "amazon-sage-random": dict(dataset="amazonProducts", model_name="sage", mode="AdaQP", assign_scheme="random"),
    "amazon-sage-adaptive": dict(dataset="amazonProducts", model_name="sage", mode="AdaQP", assign_scheme="adaptive"),
    "products-gcn-vanilla": dict(dataset="ogbn-products", model_name="gcn", mode="Vanilla", assign_scheme="uniform"),
}


def worker(a):
    import io
    from contextlib import redirect_stdout
    import torch
    import bench
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    os.environ.setdefault("ADAQP_SYNTHETIC", "1")
    os.environ.setdefault("ADAQP_SEED", "2024")
    for name in a.configs.split(","):
        base = CONFIGS[name]
        shared = {}

        def grab(tr, shared=shared):
            from adaqp_b200.helper import BitType
            from adaqp_b200.manager import GraphEngine as engine
            from tools.parity_check import current_assignment
            shared["layout"] = engine.ctx.layout
            cm = getattr(tr.assigner, "cost_model", None)
            if cm:
                k = sorted(cm)[0]
                shared["cost_model"] = {"channel": k, "alpha_ms_per_MB": float(cm[k][0]), "beta_ms": float(cm[k][1])}
            if engine.ctx.bit_type == BitType.QUANT:
                shared["assignment"] = current_assignment()
                bits = torch.cat([b for per in shared["assignment"].values() for b in per.values()])
                shared["bits_share"] = {str(b): float((bits == b).float().mean()) for b in (2, 4, 8)}

        for arm in ("ours", "reference"):
            if arm == "reference" and a.no_reference:
                continue
            kw = dict(gpus=world, steps=a.steps if arm == "ours" else a.ref_steps, warmup=3, impl=arm,
                      scale=a.scale, no_cpu_baseline=True, no_verify=not a.verify, verify_window=64,
                      assign_bits=None, before_teardown=grab if arm == "ours" else None,
                      layout=shared.get("layout"), assignment=shared.get("assignment"))
            kw.update(base)
            args = argparse.Namespace(**kw)
            os.environ["ADAQP_SYNTH_SCALE"] = str(a.scale)
            buf = io.StringIO()
            try:
                with redirect_stdout(buf):
                    (bench.run_ours if arm == "ours" else bench.run_reference)(args, rank, world)
                line = [l for l in buf.getvalue().splitlines() if l.startswith("{")]
                if rank == 0 and line:
                    rec = json.loads(line[-1])
                    rec["config_name"] = name
                    if arm == "ours" and "bits_share" in shared:
                        rec["assigned_bits_share_rank0"] = shared["bits_share"]
                    if arm == "ours" and "cost_model" in shared:
                        rec["cost_model_rank0"] = shared["cost_model"]
                    print(json.dumps(rec), flush=True)
                    if a.out:
                        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
                        with open(f"{a.out}_{name}_n{world}_{arm}.json", "w") as f:
                            json.dump(rec, f, indent=1)
            except Exception:
                print(json.dumps({"config_name": name, "impl": arm, "rank": rank, "error": traceback.format_exc()[-1500:]}), flush=True)
            torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="products-sage-adaptive,yelp-gcn-random,amazon-sage-random")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--ref-steps", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--verify", action="store_true")
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--spawn", type=int, default=0)
    a = ap.parse_args()
    if a.spawn:
        import socket
        import torch
        import torch.multiprocessing as mp
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        ngpu = max(torch.cuda.device_count(), 1)

        ctx = mp.get_context("spawn")
        procs = [ctx.Process(target=_spawn_entry, args=(r, a, port, ngpu)) for r in range(a.spawn)]
        for p in procs:
            p.start()
        for p in procs:
            p.join()
        sys.exit(max(p.exitcode or 0 for p in procs))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")
    os.environ.setdefault("LOCAL_RANK", "0")
    worker(a)


def _spawn_entry(r, a, port, ngpu):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(r),
                       "WORLD_SIZE": str(a.spawn), "LOCAL_RANK": str(r % ngpu)})
    worker(a)


if __name__ == "__main__":
    main()
