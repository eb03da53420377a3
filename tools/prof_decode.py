import cProfile, pstats, io, sys, contextlib, time
sys.path.insert(0, '.')
import numpy as np, torch
from sapr_b200 import synth
from sapr_b200.custom_hmm import HMM
feats, lab, _, _ = synth.make_corpus(66, 11, 8, 13, 80, 120, seed=5)
with contextlib.redirect_stdout(io.StringIO()):
    hs = []
    for sem in ("standard", "sapr"):
        h = HMM(8, 13, feats, model_name="w", semantics=sem); h.baum_welch([f for f, w in zip(feats, lab) if w == 0], 2); hs.append(h)
for h in hs:
    for f in feats[:5]: h.decode(f)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for f in feats: h.decode(f)
    torch.cuda.synchronize()
    print(h.semantics, 'ms per decode', (time.perf_counter() - t) / len(feats) * 1e3)
    pr = cProfile.Profile(); pr.enable()
    for f in feats: h.decode(f)
    pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(14); print(s.getvalue()[:2600])
