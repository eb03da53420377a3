"""Tuning aid (GPU): per-frame pipeline timeline of CTA 0 of the tensor-core Viterbi kernel (SAPR_TC_TRACE)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "tc_trace.txt")
os.environ["SAPR_TC_TRACE"] = out
import torch
from sapr_b200 import engine, synth
dev = torch.device("cuda", 0)
B = 148 * 128
X, offsets, labels, mu, sd = synth.device_corpus(B, 11, 8, 39, 200, 12345, dev)
A, means, var = synth.truth_models(mu, sd, 0.9)
m = engine.WordModels(11, 8, 39); m.set(means, var, A)
batch = engine.PackedBatch(X, offsets, 39, offsets.cpu().numpy(), labels)
for _ in range(2):
    m.viterbi(batch, None, engine.FP32, 0, want_scores=False, want_path=True)
torch.cuda.synchronize()
d = np.loadtxt(out).astype(np.int64)
t0 = d[:, 2:][d[:, 2:] > 0].min()
names = {0: "mma  [wait A_full | A_full ok | issued | acc done]", 1: "load [empty ok | issued | landed]",
         2: "w0   [conv start | raw ok | A arrive | acc wait | acc ok | rec done]", 3: "w6", 4: "w15"}
for ro in range(5):
    print("role", ro, names.get(ro, ""))
    for row in d[d[:, 0] == ro][:40]:
        ev = [int(v - t0) if v > 0 else -1 for v in row[2:]]
        print("  f%3d " % row[1] + " ".join("%7d" % v for v in ev))
