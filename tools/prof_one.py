"""Profiling driver (GPU): three launches of the grouped fused E-step and of the TMA Viterbi kernel at one tile per SM
(or argv[1] utterances); meant to run under ncu (-k regex:...) after a plain run."""
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
gb = engine.GroupedBatch(batch, labels, 11)
for _ in range(3):
    st, ll = m.estep_grouped(gb.X, gb.T, gb.model_start)
for _ in range(3):
    out = m.viterbi(batch, None, engine.FP32, 0, want_scores=False, want_path=True)
torch.cuda.synchronize()
print("ok", float(ll.sum().item()), int(out["best_word"].sum().item()))
