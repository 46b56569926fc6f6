"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (and grid for the mat-muls)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = re.sub(r"<.*", "", row["Kernel Name"]).replace("void ", "").replace("oh::", "").replace("unnamed>::", "").replace("(anonymous namespace)::", "")
    val = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    val = val / 1000 if unit in ("ns", "nsecond") else val * 1000 if unit in ("ms", "msecond") else val
    key = (name[:44], row.get("Grid Size", "") if "gemm" in name else "")
    agg[key][0] += 1
    agg[key][1] += val
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':44s} {'grid':16s} {'n':>6s} {'total_us':>11s} {'avg_us':>9s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]:44s} {k[1]:16s} {v[0]:6d} {v[1]:11.1f} {v[1] / v[0]:9.2f} {v[1] / tot:7.3f}")
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
