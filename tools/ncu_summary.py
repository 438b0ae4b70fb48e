"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel:
launch count, total / mean duration and share of the listed GPU time.

    python tools/ncu_summary.py gpurun_out/launches.csv > profiles/r01_launches.md
"""
from __future__ import annotations

import csv
import io
import re
import sys
from collections import defaultdict


def load(path):
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    if start < 0:
        raise SystemExit("no ncu CSV header found")
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    out = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
        out.append((r["Kernel Name"], val * scale))
    return out


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    return name if len(name) <= 110 else name[:107] + "..."


def main():
    launches = load(sys.argv[1])
    agg = defaultdict(lambda: [0, 0.0])
    for k, ms in launches:
        agg[short(k)][0] += 1
        agg[short(k)][1] += ms
    total = sum(v[1] for v in agg.values())
    print(f"# ncu launch list summary ({len(launches)} launches, {total:.3f} ms of GPU time)\n")
    print("Per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes.\n")
    print("| kernel | launches | total ms | mean us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| `{k}` | {n} | {ms:.3f} | {1e3 * ms / n:.1f} | {100 * ms / total:.1f}% |")


if __name__ == "__main__":
    main()
