"""Top stalled SASS instructions of the first kernel in an .ncu-rep (ncu --page source --print-source sass)."""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# several kernels: sections start with a 'Kernel Name' row
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
lo = starts[k]
hi = starts[k + 1] if k + 1 < len(starts) else len(rows)
body = rows[lo + 2:hi]
tot = sum(float(r[2] or 0) for r in body)
execd = sum(float(r[5] or 0) for r in body)
print(rows[lo][1][:100], "samples", tot, "warp-instr executed", execd)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
for v, i, r in sorted(((float(r[2] or 0), i, r) for i, r in enumerate(body)), reverse=True)[:n]:
    print(f"{100 * v / tot:5.1f}%  #{i:5d} exec {r[5]:>8s}  {r[1].strip()[:100]}")
