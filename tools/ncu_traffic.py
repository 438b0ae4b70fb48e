"""DRAM traffic per launch of a kernel from `ncu -i X.ncu-rep --page raw --csv` (stdin) ->
JSON consumed by bench.py (profiles/spmm_traffic.json)."""
import csv, json, sys
rows = list(csv.reader(sys.stdin))
h, units = rows[0], rows[1]
def col(name):
    return h.index(name)
def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
out = []
for r in rows[2:]:
    rd = to_bytes(r[col("dram__bytes_read.sum")], units[col("dram__bytes_read.sum")])
    wr = to_bytes(r[col("dram__bytes_write.sum")], units[col("dram__bytes_write.sum")])
    dur = float(r[col("gpu__time_duration.sum")].replace(",", ""))
    du = units[col("gpu__time_duration.sum")]
    dur_ms = dur * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "second": 1e3}.get(du, 1e-6)
    out.append({"kernel": r[col("Kernel Name")][:60], "dram_read": rd, "dram_write": wr, "ms_under_ncu": dur_ms,
                "dram_pct": float(r[col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")]),
                "l2_hit_pct": float(r[col("lts__t_sector_hit_rate.pct")]),
                "warps_active_pct": float(r[col("sm__warps_active.avg.pct_of_peak_sustained_active")])})
n = len(out)
print(json.dumps({"launches": n, "dram_bytes_per_launch": sum(o["dram_read"] + o["dram_write"] for o in out) / max(n, 1),
                  "per_launch": out, "how": sys.argv[1] if len(sys.argv) > 1 else ""}, indent=1))
