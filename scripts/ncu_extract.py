"""Extract the columns of an existing profiles/ CSV from an .ncu-rep (`ncu --page raw --csv` metric names).
  append mode (CSV whose first column is `variant`): one labelled row per captured kernel is appended
      python scripts/ncu_extract.py <report.ncu-rep> <profiles/file.csv> "<variant label>"
  rewrite mode (first column `Kernel Name`): the file is rewritten with the same columns, units row included
      python scripts/ncu_extract.py <report.ncu-rep> <profiles/file.csv>"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
label = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
with open(out) as f:
    want = next(csv.reader(f))


def pick(r, cols):
    return [r[hdr.index(c)] if c in hdr else "" for c in cols]


if want[0] == "variant":
    assert label, "append mode needs a variant label"
    with open(out, "a", newline="") as f:
        w = csv.writer(f)
        for r in rows[2:]:
            w.writerow([label] + pick(r, want[1:]))
    print(f"appended {len(rows) - 2} row(s) to {out}")
else:
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(want)
        w.writerow(pick(rows[1], want))
        for r in rows[2:]:
            w.writerow(pick(r, want))
    print(f"rewrote {out} with {len(rows) - 2} row(s)")
