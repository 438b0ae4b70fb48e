"""Top source lines by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` (stdin)."""
import csv, sys, collections
rows = list(csv.reader(sys.stdin))
hdr = None
out = []
kernel = None
for r in rows:
    if not r:
        continue
    if r[0].startswith("Kernel Name") or (len(r) == 1 and "kernel" in r[0]):
        kernel = r
    if "Source" in r and any("Sampl" in c for c in r):
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        out.append((kernel, dict(zip(hdr, r))))
if not out:
    print("no source rows parsed; header candidates:", [r[:6] for r in rows[:5]])
    sys.exit(0)
samp_col = [c for c in hdr if c.startswith("# Samples") or "Sampling Data (All)" in c or c == "Warp Stall Sampling (All Samples)"]
samp_col = samp_col[0] if samp_col else [c for c in hdr if "Sampl" in c][0]
inst_col = [c for c in hdr if "Instructions Executed" in c]
inst_col = inst_col[0] if inst_col else None
tot = sum(float(d[samp_col] or 0) for _, d in out) or 1.0
agg = collections.OrderedDict()
for _, d in out:
    key = d.get("Source", "")[:150]
    a = agg.setdefault(key, [0.0, 0.0])
    a[0] += float(d[samp_col] or 0)
    if inst_col:
        a[1] += float(d[inst_col] or 0)
print(f"columns: samples={samp_col!r} inst={inst_col!r} total samples={tot:.0f}")
for k, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[1]) if len(sys.argv) > 1 else 30]:
    print(f"{100*s/tot:6.2f}%  inst={i:12.0f}  {k}")
