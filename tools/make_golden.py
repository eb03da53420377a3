#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE's own code.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tools/make_golden.py

It imports ``/root/reference/assignment2/custom_hmm.py`` unmodified (never copied)
and records, for seeded synth-v1 inputs, the outputs of every hot-path function
(SURVEY.md 8a rows a-1..a-12):

* Rung 0 -- ``custom_hmm.HMM`` exactly as written (sapr emission, D1-D6, D9);
* Rung 1 -- a subclass overriding ONLY ``compute_emission_matrix`` with a float64
  true diagonal Gaussian on frame-major input, so that the reference's untouched
  ``forward/backward/compute_gamma/compute_xi/update_A/update_B/baum_welch/decode``
  are the oracle for the diagonal-emission kernels.

Features are float32-valued but handed to the reference as float64 arrays, so that
the reference's float32 ``np.sum`` in ``calculate_means`` (custom_hmm.py:76) does not
put 1e-7 noise into every downstream number; the float32-dtype behaviour is
recorded separately (``init_f32_*``).
"""
import copy
import io
import os
import sys
import contextlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("SAPR_REF", "/root/reference/assignment2")
sys.path.insert(0, REF)

from custom_hmm import HMM as RefHMM  # noqa: E402  (the reference, unmodified)
from sapr_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
LOG2PI = np.log(2 * np.pi)


class Rung1HMM(RefHMM):
    """Reference recursions + true diagonal-Gaussian emission on (T, D) input."""

    def compute_emission_matrix(self, features):
        X = np.asarray(features, dtype=np.float64)          # (T, D) frame-major
        T = X.shape[0]
        E = np.full((T, self.total_states), -np.inf)
        for j in range(1, self.total_states - 1):
            var = np.diag(self.B["covariance"][j])
            diff = X - self.B["mean"][j]
            E[:, j] = -0.5 * (self.num_obs * LOG2PI + np.sum(np.log(var)) + np.sum(diff * diff / var, axis=1))
        return E


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def pack(feats):
    offs = np.zeros(len(feats) + 1, dtype=np.int64)
    offs[1:] = np.cumsum([f.shape[1] for f in feats])
    X = np.concatenate([f.T for f in feats], axis=0).astype(np.float32)
    return X, offs


def per_function(hmm, feat, prefix, out, rung1=False):
    arg = feat.T if rung1 else feat
    with np.errstate(all="ignore"):
        E = hmm.compute_emission_matrix(arg)
        alpha, scale = hmm.forward(E)
        beta = hmm.backward(E, scale)
        gamma = hmm.compute_gamma(alpha, beta)
        xi = hmm.compute_xi(alpha, beta, E)
    out[prefix + "E"] = E
    out[prefix + "alpha"] = alpha
    out[prefix + "scale"] = np.float64(scale)
    out[prefix + "beta"] = beta
    out[prefix + "gamma"] = gamma
    out[prefix + "xi"] = xi
    return E, alpha, beta, gamma, xi


def rung0():
    """cfg-1 shaped, small: D=13 MFCC-scale, N=8, 3 words x 4 utterances, ragged T."""
    N, D, M, B = 8, 13, 3, 24
    feats32, labels, mu, sd = synth.make_corpus(B, M, N, D, 30, 44, seed=20241118 + 1)
    feats = [f.astype(np.float64) for f in feats32]
    out = {}
    X, offs = pack(feats32)
    out.update(X=X, offsets=offs, labels=labels, N=np.int32(N), D=np.int32(D), M=np.int32(M))

    hmm = RefHMM(N, D, feats, model_name="w0")
    out.update(init_global_mean=hmm.global_mean, init_global_cov=hmm.global_covariance,
               init_A=hmm.A.copy(), init_mean=hmm.B["mean"].copy(), init_cov=hmm.B["covariance"].copy())
    h32 = RefHMM(N, D, feats32)
    out.update(init_f32_global_mean=h32.global_mean, init_f32_global_cov=h32.global_covariance)

    # per-function vectors at the flat start (utterance 0)
    per_function(hmm, feats[0], "flat_u0_", out)
    # flat-start decode: every state identical -> exact ties everywhere (first candidate wins)
    sc, path = hmm.decode(feats[0])
    out.update(flat_dec_score=np.float64(sc), flat_dec_path=np.asarray(path, dtype=np.int32))

    # per-word training trajectories: k = 1, 2, 3 iterations from the flat start
    for w in range(M):
        wf = [feats[i] for i in range(B) if labels[i] == w]
        for k in (1, 2, 3):
            h = copy.deepcopy(hmm)
            with np.errstate(all="ignore"):
                hist = quiet(h.baum_welch, wf, k)
            out[f"bw_w{w}_k{k}_hist"] = np.asarray(hist)
            out[f"bw_w{w}_k{k}_A"] = h.A.copy()
            out[f"bw_w{w}_k{k}_mean"] = h.B["mean"].copy()
            out[f"bw_w{w}_k{k}_cov"] = h.B["covariance"].copy()
            if k == 2 and w == 0:
                trained = h
    # per-function vectors with full-covariance parameters (after 2 iterations), utterance 1
    E, alpha, beta, gamma, xi = per_function(trained, feats[1], "trained_u1_", out)
    # direct update_A / update_B on those statistics
    h = copy.deepcopy(trained)
    agg_g = np.sum(gamma[:-1], axis=0); agg_x = np.sum(xi, axis=0)
    h.update_A(agg_x, agg_g)
    h.update_B([feats[1]], [gamma])
    out.update(upd_agg_gamma=agg_g, upd_agg_xi=agg_x, upd_A=h.A.copy(), upd_mean=h.B["mean"].copy(),
               upd_cov=h.B["covariance"].copy())
    # decode (walks the first D frames only) with the 3 trained models (k=2)
    models = []
    for w in range(M):
        h = copy.deepcopy(hmm)
        h.A = out[f"bw_w{w}_k2_A"].copy()
        h.B = {"mean": out[f"bw_w{w}_k2_mean"].copy(), "covariance": out[f"bw_w{w}_k2_cov"].copy()}
        models.append(h)
    scores = np.zeros((B, M)); paths = np.zeros((B, M, D), dtype=np.int32)
    with np.errstate(all="ignore"):
        for u in range(B):
            for w in range(M):
                sc, p = models[w].decode(feats[u])
                scores[u, w] = sc; paths[u, w] = p
    out.update(dec_scores=scores, dec_paths=paths)
    np.savez_compressed(os.path.join(OUT, "rung0_d13.npz"), **out)
    print("rung0_d13:", len(out), "arrays")


def rung1(name, N, D, M, B, T_lo, T_hi, seed, a_self, perturb=0.0, take=None):
    """perturb > 0: the models are moved off the generating parameters (means + perturb * sigma * N(0,1)), the situation of
    an early Baum-Welch iteration.  With D = 39 this makes exit-constrained paths so unlikely for some utterances that the
    reference's un-normalised xi (custom_hmm.py:270-316: exp(alpha + ... + beta - logsumexp(alpha[-1]))) underflows to 0 in
    float64 for EVERY arc, the `if np.sum(xi[t]) > 0` guard (:319) skips the normalisation and the utterance contributes
    nothing to the transition statistics -- recorded here as reference behaviour (quirk D10)."""
    feats32, labels, mu, sd = synth.make_corpus(B, M, N, D, T_lo, T_hi, seed=seed)
    if take is not None:                                    # a slice of a larger corpus (keeps the files small)
        feats32, labels, B = feats32[:take], labels[:take], take
    feats = [f.astype(np.float64) for f in feats32]
    A, means, var = synth.truth_models(mu, sd, a_self)
    if perturb > 0:
        prng = np.random.default_rng(seed + 17)
        means = means + perturb * np.sqrt(var) * prng.standard_normal(means.shape)
        means[:, 0] = 0.0; means[:, -1] = 0.0
    S = N + 2
    out = {}
    X, offs = pack(feats32)
    out.update(X=X, offsets=offs, labels=labels, N=np.int32(N), D=np.int32(D), M=np.int32(M),
               A=A, means=means, var=var)
    # flat-start global covariance is needed by update_B's variance floor
    base = Rung1HMM(N, D, feats, model_name="base")
    out.update(global_cov=base.global_covariance, global_mean=base.global_mean)

    def model(w):
        h = copy.deepcopy(base)
        h.A = A[w].copy()
        cov = np.zeros((S, D, D))
        for j in range(S):
            cov[j] = np.diag(var[w, j])
        h.B = {"mean": means[w].copy(), "covariance": cov}
        return h

    models = [model(w) for w in range(M)]
    # per-function vectors, utterance 0 with its own model and with a wrong model
    per_function(models[labels[0]], feats[0], "u0_own_", out, rung1=True)
    per_function(models[(labels[0] + 1) % M], feats[0], "u0_other_", out, rung1=True)
    # E-step quantities for every utterance against its own model
    Tmax = max(f.shape[1] for f in feats)
    ll = np.zeros(B); G = np.zeros((B, S)); Xs = np.zeros((B, S)); occ = np.zeros((B, S))
    gam = np.zeros((int(offs[-1]), S))
    for u in range(B):
        h = models[labels[u]]
        E = h.compute_emission_matrix(feats[u].T)
        al, sc = h.forward(E); be = h.backward(E, sc)
        g = h.compute_gamma(al, be); x = h.compute_xi(al, be, E)
        ll[u] = np.logaddexp.reduce(al[-1])
        G[u] = g[:-1].sum(axis=0); occ[u] = g.sum(axis=0)
        Xs[u] = np.einsum("tii->i", x)
        if perturb > 0:
            out.setdefault("es_xi_rows_zero", np.zeros(B, dtype=np.int32))[u] = int(np.sum(x.sum(axis=(1, 2)) == 0.0))
        gam[offs[u]:offs[u + 1]] = g
    out.update(es_loglik=ll, es_G=G, es_xi_self=Xs, es_occ=occ, es_gamma=gam)
    # one Baum-Welch iteration per word (E-step + M-step) through the reference loop
    newA = np.zeros_like(A); newmean = np.zeros_like(means); newcov = np.zeros((M, S, D, D)); hist = np.zeros(M)
    for w in range(M):
        wf = [feats[i].T for i in range(B) if labels[i] == w]     # Rung-1 takes (T, D)
        h = copy.deepcopy(models[w])
        # update_B expects (D, T) arrays; the Rung-1 emission expects (T, D): wrap
        class _H(type(h)):
            def update_B(self, fl, gl):
                return RefHMM.update_B(self, [f.T for f in fl], gl)
        h.__class__ = _H
        hh = quiet(h.baum_welch, wf, 1)
        hist[w] = hh[0]; newA[w] = h.A; newmean[w] = h.B["mean"]; newcov[w] = h.B["covariance"]
    out.update(bw1_hist=hist, bw1_A=newA, bw1_mean=newmean, bw1_cov=newcov)
    # decoding: every utterance against every model (all frames), decoder.py:42-47 argmax
    scores = np.zeros((B, M)); paths = np.full((B, M, Tmax), -1, dtype=np.int8)
    best = np.zeros(B, dtype=np.int32)
    for u in range(B):
        bs, bw = float("-inf"), -1
        for w in range(M):
            sc, p = models[w].decode(feats[u].T)
            scores[u, w] = sc; paths[u, w, :len(p)] = p
            if sc > bs:
                bs, bw = sc, w
        best[u] = bw
    out.update(dec_scores=scores, dec_paths=paths, dec_best=best)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name + ":", len(out), "arrays; recognition accuracy",
          float(np.mean(best == labels)))


def edge_cases():
    """Short / degenerate inputs through the reference: T = N+1, T < N, T = 2."""
    N, D, M = 8, 13, 2
    feats32, labels, mu, sd = synth.make_corpus(4, M, N, D, 20, 24, seed=99)
    A, means, var = synth.truth_models(mu, sd, 0.8)
    S = N + 2
    out = dict(A=A, means=means, var=var, N=np.int32(N), D=np.int32(D))
    base = Rung1HMM(N, D, [f.astype(np.float64) for f in feats32])
    base.A = A[0].copy()
    cov = np.zeros((S, D, D))
    for j in range(S):
        cov[j] = np.diag(var[0, j])
    base.B = {"mean": means[0].copy(), "covariance": cov}
    for T in (2, 5, 8, 9, 10):
        x = feats32[0][:, :T].astype(np.float64)
        out[f"T{T}_x"] = x.T.astype(np.float32)
        with np.errstate(all="ignore"):
            E, al, be, g, xi = per_function(base, x, f"T{T}_", out, rung1=True)
            sc, p = base.decode(x.T)
        out[f"T{T}_dec_score"] = np.float64(sc)
        out[f"T{T}_dec_path"] = np.asarray(p, dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "edge_cases.npz"), **out)
    print("edge_cases:", len(out), "arrays")


def ref_pickle():
    """A model pickle exactly as the reference's train.py:74-78 writes it (pickle.dump of the reference's own
    ``custom_hmm.HMM`` object after baum_welch), plus what the reference's decode returns with that object."""
    import pickle
    N, D, B = 8, 13, 6
    feats32, labels, mu, sd = synth.make_corpus(B, 1, N, D, 30, 40, seed=20241118 + 7)
    feats = [f.astype(np.float64) for f in feats32]
    hmm = RefHMM(N, D, feats, model_name="heed")
    with np.errstate(all="ignore"):
        quiet(hmm.baum_welch, feats, 2)
    with open(os.path.join(OUT, "ref_heed_custom_2.pkl"), "wb") as f:
        pickle.dump(hmm, f)
    X, offs = pack(feats32)
    scores = np.zeros(B); paths = np.zeros((B, D), dtype=np.int32)
    with np.errstate(all="ignore"):
        for u in range(B):
            sc, p = hmm.decode(feats[u])
            scores[u] = sc; paths[u] = p
    np.savez_compressed(os.path.join(OUT, "ref_pickle.npz"), X=X, offsets=offs, dec_scores=scores, dec_paths=paths,
                        A=hmm.A, mean=hmm.B["mean"], cov=hmm.B["covariance"])
    print("ref_pickle: state keys", sorted(hmm.__dict__))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "ref_pickle":
        ref_pickle()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "mismatch":
        rung1("rung1_mismatch_d39", N=8, D=39, M=11, B=700, T_lo=40, T_hi=40, seed=1200, a_self=0.9, perturb=0.3, take=22)
        sys.exit(0)
    rung0()
    rung1("rung1_d39", N=8, D=39, M=11, B=22, T_lo=40, T_hi=60, seed=20241118 + 2, a_self=0.9)
    rung1("rung1_d13", N=8, D=13, M=11, B=33, T_lo=24, T_hi=40, seed=20241118 + 3, a_self=0.85)
    rung1("rung1_mismatch_d39", N=8, D=39, M=11, B=700, T_lo=40, T_hi=40, seed=1200, a_self=0.9, perturb=0.3, take=22)
    edge_cases()
    ref_pickle()
