"""Tuning aid (GPU): per-frame pipeline timeline of CTA 0 of the tensor-core E-step kernel (SAPR_ET_TRACE)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "et_trace.txt")
os.environ["SAPR_ET_TRACE"] = out
import torch
from sapr_b200 import engine, synth
dev = torch.device("cuda", 0)
B = 148 * 128 * 4
X, offsets, labels, mu, sd = synth.device_corpus(B, 11, 8, 39, 200, 12345, dev)
A, means, var = synth.truth_models(mu, sd, 0.9)
m = engine.WordModels(11, 8, 39); m.set(means, var, A)
batch = engine.PackedBatch(X, offsets, 39, offsets.cpu().numpy(), labels)
order = engine.group_by_model(labels)
for _ in range(2):
    m.estep(batch, labels, order, engine.FP32)
torch.cuda.synchronize()
d = np.loadtxt(out).astype(np.int64)
t0 = d[:, 2:][d[:, 2:] > 0].min()
names = {0: "mma  [start | A_full ok | acc_empty ok | issued]", 1: "conv4 [start | raw ok | A_free ok | done]",
         2: "rec0 [acc wait | acc ok | loaded | -]", 3: "load [start | empty ok | issued | -] (row = stage*4+40)"}
for ro in range(4):
    print("role", ro, names[ro])
    for row in d[d[:, 0] == ro][:28]:
        ev = [int(v - t0) if v > 0 else -1 for v in row[2:]]
        print("  f%3d " % row[1] + " ".join("%7d" % v for v in ev))
