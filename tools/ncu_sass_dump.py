"""Every SASS row of an ncu report's source page with its execution and stall-sample counts (run on the GPU box).
python tools/ncu_sass_dump.py report.ncu-rep out.csv"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
r = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True)
rows = list(csv.reader(io.StringIO(r.stdout)))
hi = next(i for i, row in enumerate(rows) if any("Sampl" in c for c in row))
h = rows[hi]
si = next(i for i, c in enumerate(h) if "Sampl" in c and "All" in c)
ei = next(i for i, c in enumerate(h) if c.strip() == "Instructions Executed")
with open(out, "w", newline="") as fh:
    w = csv.writer(fh)
    w.writerow(["address", "sass", "samples", "executed"])
    for row in rows[hi + 1:]:
        if len(row) == len(h):
            w.writerow([row[0], row[1].strip(), row[si], row[ei]])
