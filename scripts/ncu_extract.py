"""Append one row per captured kernel of an .ncu-rep to a profiles/ CSV that already has a header (columns = the metric
names of `ncu --page raw --csv`, first column `variant`).
Usage: python scripts/ncu_extract.py <report.ncu-rep> <profiles/file.csv> "<variant label>" """
import csv
import io
import subprocess
import sys

rep, out, label = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
with open(out) as f:
    want = next(csv.reader(f))
with open(out, "a", newline="") as f:
    w = csv.writer(f)
    for r in rows[2:]:
        w.writerow([label] + [r[hdr.index(c)] if c in hdr else "" for c in want[1:]])
print(f"appended {len(rows) - 2} row(s) to {out}")
