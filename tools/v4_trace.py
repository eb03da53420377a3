"""Tuning aid (GPU): per-frame pipeline timeline of CTA 0 of k_viterbi_v4 (SAPR_V_TRACE): usage v4_trace.py [out] [first] [count]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "v4_trace.txt")
f0 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
nf = int(sys.argv[3]) if len(sys.argv) > 3 else 24
os.environ["SAPR_V_TRACE"] = out
import torch
from sapr_b200 import engine, synth
dev = torch.device("cuda", 0)
B = 148 * 128 * 2
X, offsets, labels, mu, sd = synth.device_corpus(B, 11, 8, 39, 200, 12345, dev)
A, means, var = synth.truth_models(mu, sd, 0.9)
m = engine.WordModels(11, 8, 39); m.set(means, var, A)
batch = engine.PackedBatch(X, offsets, 39, offsets.cpu().numpy(), labels)
for _ in range(2):
    m.viterbi(batch, None, engine.FP32, 0, want_scores=False, want_path=True)
torch.cuda.synchronize()
d = np.loadtxt(out).astype(np.int64)
t0 = d[(d[:, 0] == 0) & (d[:, 1] == f0), 2][0]
names = {0: "mma  [wait A_full | A_full ok | acc_empty ok | issued]", 1: "tma (per 4 frames) [ring stage free | issued | landed]",
         2: "conv w16 [start | raw ok (first of 4) | A_empty ok | A_full arrive]", 3: "rec w0 [wait acc_full | acc ok | acc_empty arrive | done]", 4: "rec w15"}
mm = d[d[:, 0] == 0]
a, b = mm[mm[:, 1] == 20][0], mm[mm[:, 1] == 220][0]
print("SM clock under the kernel: %.0f MHz (%d cycles in %d ns over 200 frames)" % ((b[2] - a[2]) * 1e3 / (b[3] - a[3]), b[2] - a[2], b[3] - a[3]))
for ro in range(5):
    print("role", ro, names.get(ro, ""))
    lo, hi = (f0 // 4, (f0 + nf) // 4 + 1) if ro == 1 else (f0, f0 + nf)
    for row in d[(d[:, 0] == ro) & (d[:, 1] >= lo) & (d[:, 1] < hi)]:
        ev = [int(v - t0) if v > 0 else -1 for v in row[2:]]
        print("  f%3d " % row[1] + " ".join("%7d" % v for v in ev))
