#!/usr/bin/env python
"""Top stall sites of an ncu source-page CSV (ncu -i rep --page source --csv): usage ncu_src_top.py file.csv [N] [column]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
col = sys.argv[3] if len(sys.argv) > 3 else "# Samples"
ci = hdr.index(col); si = hdr.index("Source"); ai = hdr.index("Address"); ei = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(float(r[ci] or 0) for r in body)
print("total", col, tot, "instructions executed", sum(float(r[ei] or 0) for r in body))
idx = sorted(range(len(body)), key=lambda i: -float(body[i][ci] or 0))[:n]
for i in sorted(idx):
    r = body[i]
    st = sorted(((float(r[j] or 0), hdr[j][6:]) for j in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {r[ai][-5:]} {float(r[ci] or 0):7.0f} {float(r[ei] or 0):9.0f}  {r[si][:70]:70s} " + " ".join(f"{b}={a:.0f}" for a, b in st if a > 0))
