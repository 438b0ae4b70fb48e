"""Event-timed sweep of the aggregation kernel variants on an ogbn-products-shaped partition.

    python tools/bench_spmm.py [--scale 1.0] [--world 1] [--reps 10] [--json out.json]

Variants (library option `spmm_impl`, adaqp_b200/_lib.py): 1 = register-staged gather (default),
2 = cp.async lane-private ring, 3 = TMA `tile::gather4` ring (4 neighbour rows per request),
4 = one TMA bulk copy per neighbour row; plus the streaming hints of v1 (bit 0: st.global.cs for
the output rows, bit 1: ld.global.cs for the index stream) and rows per grab.  Every variant's
output is compared with v1's (max abs difference relative to the largest magnitude)."""
import argparse, json, os, sys
import numpy as np, torch, yaml
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--world", type=int, default=1)
    ap.add_argument("--dims", type=str, default="256,100")
    ap.add_argument("--json", type=str, default=None)
    ap.add_argument("--variants", type=str, default="1:0:0,1:1:0,1:2:0,1:3:0,3:0:1,3:0:2,3:0:4,3:0:8,4:0:4,2:0:0",
                    help="impl:hints:rows_per_grab (0 = default), comma separated")
    a = ap.parse_args()
    from adaqp_b200 import build
    build.build()
    from adaqp_b200 import _lib
    from adaqp_b200.manager.graph import LocalGraph, spmm
    from adaqp_b200.manager.layout import prepare_all_in_process
    from adaqp_b200.manager.partition_synth import spec_from_config
    cfg = yaml.safe_load(open(os.path.join(ROOT, "adaqp_b200", "config", "ogbn-products.yaml")))
    spec = spec_from_config(cfg, a.world, a.scale)
    L = prepare_all_in_process(spec)[0]
    dev = torch.device("cuda:0")
    g = LocalGraph(L.indptr, L.indices, L.in_degrees, L.out_degrees, L.n_inner, L.n_halo, dev)
    nnz = int(L.indptr[-1])
    results = []
    for F in [int(x) for x in a.dims.split(",")]:
        xl = torch.randn(L.n_inner, F, device=dev)
        xh = torch.randn(max(L.n_halo, 1), F, device=dev) if L.n_halo else None
        out = torch.empty(L.n_inner, F, device=dev)
        ref = None
        for var in a.variants.split(","):
            impl, hints, grab = (int(x) for x in var.split(":"))
            _lib.set_option("spmm_impl", impl)
            _lib.set_option("spmm_hints", hints)
            _lib.set_option("spmm_rows_per_grab", grab)
            for _ in range(3):
                spmm(g, xl, xh, g.norm["out_-0.5"], g.norm["in_-0.5"], out=out)
            ts = []
            for _ in range(a.reps):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); spmm(g, xl, xh, g.norm["out_-0.5"], g.norm["in_-0.5"], out=out); e.record()
                torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
            if ref is None:
                ref = out.clone()
            diff = float((out - ref).abs().max() / ref.abs().max())
            ms = float(np.median(ts))
            comp = 4 * nnz + 8 * (L.n_inner + 1) + 4 * F * (2 * L.n_inner + L.n_halo) + 4 * (2 * L.n_inner + L.n_halo)
            results.append({"impl": impl, "hints": hints, "rows_per_grab": grab, "F": F, "rows": L.n_inner, "nnz": nnz,
                            "ms": ms, "ms_min": float(min(ts)), "no_reuse_GBps": 4 * F * nnz / ms / 1e6,
                            "compulsory_GBps": comp / ms / 1e6, "max_rel_diff_vs_v1": diff})
            print(json.dumps(results[-1]), flush=True)
    if a.json:
        with open(a.json, "w") as f:
            json.dump({"scale": a.scale, "world": a.world, "results": results}, f, indent=1)


if __name__ == "__main__":
    main()
