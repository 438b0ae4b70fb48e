"""Time the aggregation kernel alone on an ogbn-products-shaped single partition.
ADAQP_SPMM=1 selects v1 (register gather), default v2 (TMA ring); ADAQP_SPMM_HINTS=0 drops the
L2 eviction hints.   python tools/bench_spmm.py [--scale 0.25] [--reps 10]"""
import argparse, json, os, sys
import numpy as np, torch, yaml
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.25)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--world", type=int, default=1)
    a = ap.parse_args()
    from adaqp_b200 import build
    build.build()
    from adaqp_b200.manager.graph import LocalGraph, spmm
    from adaqp_b200.manager.layout import prepare_all_in_process
    from adaqp_b200.manager.partition_synth import spec_from_config
    cfg = yaml.safe_load(open(os.path.join(ROOT, "adaqp_b200", "config", "ogbn-products.yaml")))
    spec = spec_from_config(cfg, a.world, a.scale)
    L = prepare_all_in_process(spec)[0]
    dev = torch.device("cuda:0")
    g = LocalGraph(L.indptr, L.indices, L.in_degrees, L.out_degrees, L.n_inner, L.n_halo, dev)
    nnz = int(L.indptr[-1])
    for F in (256, 100):
        xl = torch.randn(L.n_inner, F, device=dev)
        xh = torch.randn(max(L.n_halo, 1), F, device=dev) if L.n_halo else None
        out = torch.empty(L.n_inner, F, device=dev)
        for _ in range(3):
            spmm(g, xl, xh, g.norm["out_-0.5"], g.norm["in_-0.5"], out=out)
        ts = []
        for _ in range(a.reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); spmm(g, xl, xh, g.norm["out_-0.5"], g.norm["in_-0.5"], out=out); e.record()
            torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
        ms = float(np.median(ts))
        comp = 4 * nnz + 8 * (L.n_inner + 1) + 4 * F * (2 * L.n_inner + L.n_halo) + 4 * (2 * L.n_inner + L.n_halo)
        print(json.dumps({"impl": os.environ.get("ADAQP_SPMM", "2"), "hints": os.environ.get("ADAQP_SPMM_HINTS", "1"), "F": F,
                          "rows": L.n_inner, "nnz": nnz, "ms": ms, "no_reuse_GBps": 4 * F * nnz / ms / 1e6,
                          "compulsory_GBps": comp / ms / 1e6, "checksum": float(out.double().abs().sum())}), flush=True)


if __name__ == "__main__":
    main()
