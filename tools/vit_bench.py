"""Tuning aid (GPU): device-resident cfg-2 Viterbi step, kernel time from the context's event pairs.
usage: python tools/vit_bench.py [utts] [steps]   (env SAPR_TMA=0 selects the per-row bulk-copy kernel)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sapr_b200 import _lib, engine, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
X, offsets, labels, mu, sd = synth.device_corpus(B, 11, 8, 39, 200, 20241120, dev)
A, means, var = synth.truth_models(mu, sd, 0.9)
m = engine.WordModels(11, 8, 39); m.set(means, var, A)
batch = engine.PackedBatch(X, offsets, 39, offsets.cpu().numpy(), labels)
ctx = _lib.default_context()
for _ in range(3):
    out = m.viterbi(batch, None, engine.FP32, 0, want_scores=False, want_path=True)
torch.cuda.synchronize()
ctx.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    out = m.viterbi(batch, None, engine.FP32, 0, want_scores=False, want_path=True)
e1.record(); torch.cuda.synchronize()
k_ms, k_n = ctx.profile_read(0); f_ms, f_n = ctx.profile_read(1); r_ms, r_n = ctx.profile_read(6)
ctx.profile(False)
acc = float((out["best_word"] == labels).float().mean().item())
flagged = ctx.viterbi_flagged()
print(json.dumps({"tma": os.environ.get("SAPR_TMA", "1"), "utts": B, "step_ms": e0.elapsed_time(e1) / steps, "kernel_ms": k_ms / max(k_n, 1),
                  "finish_ms": f_ms / max(f_n, 1), "flagged": flagged, "redo_ms": r_ms / max(r_n, 1),  "frac": 31412 * B / (k_ms / max(k_n, 1) / 1e3) / 1e9 / 6552.3, "word_acc": acc,
                  "wsum": int(out["best_word"].sum().item()), "psum": int(out["path"].to(torch.int64).sum().item())}))
