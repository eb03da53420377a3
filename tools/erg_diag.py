"""Locate the first frame where the tensor-core ergodic forward departs from the float64 kernel (prefix scoring)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from sapr_b200.hmmlearn_hmm import GaussianHMM
S, D = 192, 39
rng = np.random.default_rng(S + D)
means = 2.0 * rng.standard_normal((S, D)); var = rng.uniform(0.5, 1.5, (S, D)) ** 2
tm = rng.dirichlet(np.ones(S), size=S); sp = rng.dirichlet(np.ones(S))
tm = 1e-4 * tm + (1 - 1e-4) * (0.6 * np.eye(S) + 0.4 * np.roll(np.eye(S), 1, axis=1))
tm[3] = 0.0; tm[3, 3] = 0.7; tm[3, 4] = 0.3
lengths = [int(x) for x in rng.integers(1, 90, size=150)]
lengths[0], lengths[1], lengths[140] = 1, 2, 120
states = np.empty(sum(lengths), dtype=np.int64)
o = 0
for T in lengths:
    st = rng.choice(S, p=sp)
    for t in range(T):
        states[o + t] = st
        st = rng.integers(0, S) if t % 7 == 3 and st != 3 else rng.choice(S, p=tm[st])
    o += T
X = (means[states] + np.sqrt(var[states]) * rng.standard_normal((sum(lengths), D))).astype(np.float32)
model = GaussianHMM(n_components=S, covariance_type="diag", n_iter=1, init_params="")
model.means_, model.covars_, model.transmat_, model.startprob_ = means, var, tm, sp
ref = model.score_each(X, lengths).cpu().numpy()
got = model.score_each(X, lengths, precision="tc").cpu().numpy()
err = np.abs(got - ref)
bad = np.argsort(-err)[:5]
print("worst utterances", bad, err[bad], np.asarray(lengths)[bad])
offs = np.concatenate([[0], np.cumsum(lengths)])
u = int(bad[0])
if err[u] > 0.1:
    xs = X[offs[u]:offs[u + 1]]
    T = len(xs)
    Xp = np.concatenate([xs[:k] for k in range(1, T + 1)])
    lp = list(range(1, T + 1))
    r = model.score_each(Xp, lp).cpu().numpy(); g = model.score_each(Xp, lp, precision="tc").cpu().numpy()
    d = g - r
    k = int(np.argmax(np.abs(d) > 0.1))
    print("first bad prefix length", k + 1, "diffs around", d[max(0, k - 2):k + 3])
    print("states around", states[offs[u] + max(0, k - 3):offs[u] + k + 2])
    print("per-frame ref increments", np.diff(r)[max(0, k - 3):k + 2], "tc", np.diff(g)[max(0, k - 3):k + 2])
