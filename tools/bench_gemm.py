"""Event-timed comparison of the tcgen05 3xTF32 GEMM with torch.matmul (fp32 SIMT) at the epoch's shapes."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from adaqp_b200 import build
    build.build()
    from adaqp_b200 import dense
    dev = torch.device("cuda:0")
    out = []
    for M in (2449029, 306129):
        for K, N in ((100, 256), (256, 256), (256, 47)):
            x = torch.randn(M, K, device=dev)
            w = torch.randn(N, K, device=dev) * 0.1
            b = torch.randn(N, device=dev)
            wt = w.t().contiguous()
            res = {"M": M, "K": K, "N": N}
            from adaqp_b200 import _lib

            def with_kb(kb):
                def run():
                    _lib.set_option("gemm_block_k", kb)
                    return dense.gemm_nt(x, w, b)
                return run
            for name, fn in (("tcgen05_3xtf32", with_kb(32)), ("tcgen05_3xtf32_k16", with_kb(16)), ("torch_fp32", lambda: torch.addmm(b, x, wt))):
                for _ in range(3):
                    fn()
                ts = []
                for _ in range(10):
                    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
                res[name + "_ms"] = float(np.median(ts))
            byt = 4 * (M * K + M * N)
            res["hbm_GBps_tcgen05"] = byt / res["tcgen05_3xtf32_ms"] / 1e6
            res["speedup"] = res["torch_fp32_ms"] / res["tcgen05_3xtf32_ms"]
            got, want = dense.gemm_nt(x[:4096], w, b), (x[:4096].double() @ w.double().t() + b.double())
            res["max_rel_err_vs_f64"] = float(((got.double() - want).abs().max() / want.abs().max()).item())
            out.append(res)
            print(json.dumps(res), flush=True)
    for M in (2449029, 306129):
        for N, K in ((256, 100), (256, 256)):
            dy = torch.randn(M, N, device=dev) * 1e-3
            x = torch.relu(torch.randn(M, K, device=dev))
            res = {"op": "wgrad dY^T X", "M": M, "N": N, "K": K}
            for name, fn in (("tcgen05_3xtf32", lambda: dense.gemm_tn(dy, x)), ("torch_fp32", lambda: dy.t() @ x)):
                for _ in range(3):
                    fn()
                ts = []
                for _ in range(10):
                    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
                res[name + "_ms"] = float(np.median(ts))
            res["hbm_GBps_tcgen05"] = 4 * M * (N + K) / res["tcgen05_3xtf32_ms"] / 1e6
            res["speedup"] = res["torch_fp32_ms"] / res["tcgen05_3xtf32_ms"]
            want = dy.double().t() @ x.double()
            res["max_rel_err_vs_f64"] = float(((dense.gemm_tn(dy, x).double() - want).abs().max() / want.abs().max()).item())
            out.append(res)
            print(json.dumps(res), flush=True)
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
