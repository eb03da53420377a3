"""Diagnostic (GPU): error profile of the tensor-core emission tile against the float64 oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as orc
from sapr_b200 import engine as eng


def main():
    for name in ("rung1_d39", "rung1_d13"):
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"), allow_pickle=True))
        from conftest import split_features
        feats = split_features(g)
        M = g["means"].shape[0]
        m = eng.WordModels(M, int(g["N"]), int(g["D"]))
        m.set(g["means"], g["var"], g["A"])
        batch = eng.PackedBatch.from_features(feats)
        E = m.tc_emission(batch).cpu().numpy()
        offs = batch.offsets_host
        errs, refs = [], []
        for u in range(batch.B):
            for w in range(M):
                ref = orc.emission_diag(feats[u], g["means"][w], g["var"][w])[:, 1:-1]
                got = E[offs[u]:offs[u + 1], w * 8:(w + 1) * 8]
                errs.append((got - ref).ravel()); refs.append(ref.ravel())
        e = np.concatenate(errs); r = np.abs(np.concatenate(refs))
        print(name, "n", e.size, "nan", int(np.isnan(e).sum()), "max|err|", np.nanmax(np.abs(e)), "mean err", np.nanmean(e),
              "rms", np.sqrt(np.nanmean(e * e)))
        for lo, hi in ((0, 50), (50, 100), (100, 300), (300, 1000), (1000, 1e9)):
            k = (r >= lo) & (r < hi)
            if k.any():
                print(f"  |ref| in [{lo},{hi}): n={int(k.sum())} max|err|={np.abs(e[k]).max():.3e} rms={np.sqrt(np.mean(e[k]**2)):.3e} "
                      f"mean={e[k].mean():.3e}")


if __name__ == "__main__":
    main()
