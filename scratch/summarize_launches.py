"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel summary csv.
usage: python scratch/summarize_launches.py gpurun_out/r2_launches_fused.csv profiles/r2_launches_fused_summary.csv"""
import csv, re, sys
from collections import defaultdict

src, dst = sys.argv[1], sys.argv[2]
rows = []
with open(src, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    head = name.split("(")[0].replace("void ", "")
    if head.startswith("at::"):
        short = head[:48]
    else:
        short = re.sub(r"<[^<>]*>$", "", head.replace("<unnamed>::", "")).split("::")[-1]
    unit = r["Metric Unit"]
    v = float(r["Metric Value"].replace(",", ""))
    ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    rows.append((short, ms))
agg = defaultdict(lambda: [0, 0.0])
for k, ms in rows:
    agg[k][0] += 1
    agg[k][1] += ms
tot = sum(v[1] for v in agg.values())
with open(dst, "w") as f:
    f.write("kernel,launches,total_ms,share_pct,avg_ms\n")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{n},{ms:.3f},{100 * ms / tot:.2f},{ms / n:.4f}\n")
print(f"{len(rows)} launches, {tot:.1f} ms")
