"""Summary of a k_viterbi_v4 timeline (tools/v4_trace.py output file): usage v4_trace_stats.py [trace.txt] [first] [last]"""
import sys
import numpy as np
f = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/v4_trace.txt"
f0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
f1 = int(sys.argv[3]) if len(sys.argv) > 3 else 124
d = np.loadtxt(f).astype(np.int64)
role = lambda r: d[d[:, 0] == r][:, 2:]
m, c, r0, r15 = role(0), role(2), role(3), role(4)
print("period", (m[f1, 0] - m[f0, 0]) / (f1 - f0))
print("mma: barrier wait", np.mean(m[f0:f1, 2] - m[f0:f1, 0]), "issue", np.mean(m[f0:f1, 3] - m[f0:f1, 2]))
for name, r in (("rec0", r0), ("rec15", r15)):
    print(name, "wait", np.mean(r[f0:f1, 1] - r[f0:f1, 0]), "arith", np.mean(r[f0:f1, 2] - r[f0:f1, 1]), "ldwait+arrive", np.mean(r[f0:f1, 3] - r[f0:f1, 2]),
          "gap", np.mean(r[f0 + 1:f1 + 1, 0] - r[f0:f1, 3]), " start of frame f behind mma issued(f):", np.mean(r[f0:f1, 1] - m[f0:f1, 3]))
print("conv: lead of A_full arrive before mma wait:", np.mean(m[f0:f1, 0] - c[f0:f1, 3]))
print("mma wait by frame mod 4:", [float(np.mean((m[f0:f1, 2] - m[f0:f1, 0])[i::4])) for i in range(4)])
print("mma issue by frame mod 4:", [float(np.mean((m[f0:f1, 3] - m[f0:f1, 2])[i::4])) for i in range(4)])
print("rec0 gap by frame mod 4:", [float(np.mean((r0[f0 + 1:f1 + 1, 0] - r0[f0:f1, 3])[i::4])) for i in range(4)])
