"""Top source lines / SASS instructions of an ncu report by warp-stall samples (run on the GPU box: the reports are too big
to bring back).   python tools/ncu_source_top.py report.ncu-rep out_prefix [N]
Writes <out_prefix>.cuda.txt (per CUDA source line) and <out_prefix>.sass.txt (per SASS instruction), top N rows each."""
import csv
import io
import subprocess
import sys

rep, out, n = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 70
for view in ("cuda", "sass"):
    r = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", view, "--csv"], capture_output=True, text=True)
    txt = r.stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr_i = next((i for i, row in enumerate(rows) if any("Sampl" in c for c in row)), None)
    with open(f"{out}.{view}.txt", "w") as fh:
        if hdr_i is None:
            fh.write("no sampling columns found\n" + txt[:4000] + "\nSTDERR\n" + r.stderr[:2000])
            continue
        hdr = rows[hdr_i]
        fh.write("columns: " + " | ".join(hdr) + "\n")
        si = next(i for i, c in enumerate(hdr) if "Sampl" in c and "All" in c) if any("Sampl" in c and "All" in c for c in hdr) else next(i for i, c in enumerate(hdr) if "Sampl" in c)
        ei = next((i for i, c in enumerate(hdr) if c.strip() == "Instructions Executed"), None)
        body = []
        for row in rows[hdr_i + 1:]:
            if len(row) != len(hdr):
                continue
            try:
                body.append((float(row[si].replace(",", "") or 0), row))
            except ValueError:
                pass
        tot = sum(b[0] for b in body) or 1.0
        fh.write(f"total samples {tot:.0f}, rows {len(body)}\n")
        for s, row in sorted(body, key=lambda b: -b[0])[:n]:
            keep = [row[i] for i in range(len(hdr)) if i in (0, 1, 2, si, ei)]
            fh.write(f"{100 * s / tot:6.2f}%  " + " | ".join(x[:160] for x in keep) + "\n")
