"""Profiling driver (GPU): a few launches of the cfg-2 Viterbi step at one tile per SM (or argv[1] utterances); meant to run
under ncu (-k regex:k_viterbi) after a plain run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sapr_b200 import engine, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128
dev = torch.device("cuda", 0)
X, offsets, labels, mu, sd = synth.device_corpus(B, 11, 8, 39, 200, 20241122, dev)
A, means, var = synth.truth_models(mu, sd, 0.9)
m = engine.WordModels(11, 8, 39); m.set(means, var, A)
batch = engine.PackedBatch(X, offsets, 39, offsets.cpu().numpy(), labels)
for _ in range(3):
    out = m.viterbi(batch, None, engine.FP32, 0, want_scores=False, want_path=True)
torch.cuda.synchronize()
print("ok", int(out["best_word"].sum().item()))
