"""Tuning aid (CPU): per-role stall summary of k_estep_tc from an `ncu --set full --import-source on` report.
usage: python tools/et_src_profile.py gpurun_out/prof.ncu-rep [min_share]"""
import csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.006
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[1]; data = rows[2:]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
first = {}
for k, r in enumerate(data):
    for key in ('UBLKCP', 'UTCHMMA', 'F2FP', 'LDTM', 'MUFU.RCP'):
        if key in r[isrc] and key not in first: first[key] = k
print(first)
tot = sum(int(r[isamp]) for r in data); print('total samples', tot)
b = [0, first['UBLKCP'] - 120, first['UTCHMMA'] - 150, first['F2FP'] - 250, first['LDTM'] - 80, first['MUFU.RCP'] - 300, len(data)]
names = ['setup', 'loader', 'mma', 'conv', 'rec fwd', 'rec bwd']
for n, lo, hi in zip(names, b[:-1], b[1:]):
    ss = sum(int(r[isamp]) for r in data[lo:hi]); ex = sum(int(r[iex]) for r in data[lo:hi])
    print('%-8s [%4d,%4d) samples %6d (%.1f%%) warp-instr %d' % (n, lo, hi, ss, 100 * ss / tot, ex))
for k, r in enumerate(data):
    s = int(r[isamp])
    if s >= tot * thr:
        st = sorted(((hdr[i][6:], int(r[i])) for i in stall), key=lambda x: -x[1])[:2]
        print(k, s, r[iex], r[isrc].strip()[:64], st)
