"""bench.py host-side helpers (no GPU): algorithmic-byte accounting, clock summary, peaks."""
import importlib
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
bench = importlib.import_module("bench")


def test_spmm_algorithmic_bytes_matches_survey_formula():
    from adaqp_b200.manager.layout import prepare_all_in_process
    from adaqp_b200.manager.partition_synth import SynthSpec
    spec = SynthSpec(name="t", num_nodes=3000, num_edges=3000 * 12, num_parts=2, num_feats=20, num_classes=5,
                     cross_fraction=0.2, community_size=64, seed=1)
    L = prepare_all_in_process(spec)[0]
    eng = SimpleNamespace(layout=L)
    dims = [20, 256, 256]
    total, launches = bench.spmm_algorithmic_bytes(eng, dims, use_parallel=False)
    assert launches == 5
    nnz, rows, nsrc = int(L.indptr[-1]), L.n_inner, L.n_inner + L.n_halo
    want = sum(4 * nnz + 8 * (rows + 1) + 4 * F * (nsrc + rows) + 4 * (nsrc + rows) for F in (20, 256, 256, 256, 256))
    assert total == want
    total_p, launches_p = bench.spmm_algorithmic_bytes(eng, dims, use_parallel=True)
    assert launches_p == 10 and total_p < want + 5 * (8 + 4 * 256 * L.n_inner + 8 * L.n_inner)


def test_clock_summary_and_peaks():
    s = bench.ClockSampler(0)
    assert s.summary()["reasons"] == ["unsampled"]
    s.samples = [["1965", "1965", "Not Active", "Not Active", "Not Active", "Active"],
                 ["1900", "1965", "Not Active", "Not Active", "Not Active", "Not Active"]]
    out = s.summary()
    assert out["sm_max_mhz"] == 1965 and out["reasons"] == ["sw_power_cap"] and out["sm_mhz"] in (1900, 1965)
    peaks, kind = bench.measured_peaks()
    assert peaks["hbm_gbs"] > 1000 and kind in ("measured", "fallback")


def test_dram_rate_block():
    import bench
    r = bench.dram_rate(54702720400.0, 5, 56.1, 6480.5)
    assert abs(r["GBps"] - 4875.5) < 1.0 and abs(r["frac_of_peak"] - 0.7523) < 1e-3
    assert bench.dram_rate(None, 5, 56.1, 6480.5) is None and bench.dram_rate(1e9, 0, 1.0, 6480.5) is None
