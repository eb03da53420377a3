"""Tuning aid (GPU): engine.train_words wall time per iteration, grouped fused E-step vs the general path, at several sizes."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from sapr_b200 import engine, synth
dev = torch.device("cuda", 0)
res = {}
for B, D, T in ((330, 13, 100), (330, 39, 100), (20000, 39, 200), (200000, 39, 200)):
    X, offsets, labels, mu, sd = synth.device_corpus(B, 11, 8, D, T, 5, dev)
    A, means, var = synth.truth_models(mu, sd, 0.9)
    batch = engine.PackedBatch(X, offsets, D, offsets.cpu().numpy(), labels)
    for mode in ("1", "0"):
        os.environ["SAPR_GROUPED"] = mode
        m = engine.WordModels(11, 8, D); m.set(means + 0.1, var, A)
        engine.train_words(m, batch, labels, 2, 1e-3, tol=0.0)
        torch.cuda.synchronize(); t = time.perf_counter()
        n = 8
        engine.train_words(m, batch, labels, n, 1e-3, tol=0.0)
        torch.cuda.synchronize()
        res[f"B{B}_D{D}_T{T}_{'grouped' if mode == '1' else 'general'}"] = round((time.perf_counter() - t) / n * 1e3, 3)
print(json.dumps(res))
