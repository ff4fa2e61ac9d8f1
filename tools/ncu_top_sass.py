"""Top SASS instructions by warp-stall samples from `ncu --page source --csv` output.
usage: ncu_top_sass.py file.csv [min_fraction]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
h = next(i for i, r in enumerate(rows) if len(r) > 4 and r[0] == "Address")
hdr = rows[h]
si = hdr.index("# Samples")
data = [r for r in rows[h + 1:] if len(r) > si]
tot = sum(int(r[si]) for r in data if r[si].isdigit())
print("total samples", tot, "instructions", len(data))
stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
for n, r in enumerate(data):
    s = int(r[si]) if r[si].isdigit() else 0
    if s > tot * frac:
        top = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print(f"{n:5d} {s:6d} {100 * s / tot:5.1f}%  {r[1][:80]:80s} {top}")
