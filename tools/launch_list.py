#!/usr/bin/env python
"""ncu launch-list CSV (--metrics gpu__time_duration.sum) -> profiles/ text: own kernels only, plus per-kernel totals and shares.
usage: launch_list.py launches.csv out.txt "command line the list was taken from" """
import csv, sys, collections
src, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ki, vi, gi, bi, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Grid Size", "Block Size", "ID"))
lib = ("void at::", "at::", "void cub", "void thrust", "ncclDev", "void gemv", "void cutlass", "void (anonymous")
own = [r for r in rows[1:] if not r[ki].startswith(lib) and "native::" not in r[ki] and "at_cuda_detail" not in r[ki] and "<unnamed>" not in r[ki]]
tot = collections.OrderedDict()
for r in own:
    k = r[ki].split("(")[0][:60]
    d = tot.setdefault(k, [0, 0.0]); d[0] += 1; d[1] += float(r[vi].replace(",", ""))
all_ns = sum(v[1] for v in tot.values())
with open(out, "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {cmd}\n")
    f.write("# cold-cache, serialised: compare SHARES, not absolute times.  Own kernels only.\n")
    f.write("# per kernel: launches, total ns, share of the own-kernel time\n")
    for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        f.write(f"#   {k:60s} {n:5d} {ns:14.0f} {ns / all_ns:7.3f}\n")
    f.write("# columns: id, kernel, grid, block, duration ns\n")
    for r in own:
        f.write(f"{r[ii]:>5s} {r[ki][:100]:100s} {r[gi]:>15s} {r[bi]:>15s} {r[vi].replace(',', ''):>12s}\n")
print("own launches", len(own), "of", len(rows) - 1)
