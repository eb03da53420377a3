"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Wraps ``oracle/sapr_oracle.c`` (a float64 restatement of the reference's
``assignment2/custom_hmm.py``; every C function cites the reference lines it
follows).  Only ``tests/``, ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs and ``__graft_entry__.smoke()`` may import this module; the
product package ``sapr_b200`` never does.

Features are handed over the way the reference holds them -- a list of ``(D, T)``
arrays -- and are transposed to the frame-major layout the C code uses.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc, seconds)."""
    src = os.path.join(_HERE, "sapr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_decode.restype = C.c_double
        _lib.orc_hl_forward.restype = C.c_double
        _lib.orc_hl_viterbi.restype = C.c_double
        _lib.orc_hl_estep.restype = C.c_double
        _lib.orc_baum_welch.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def pack(features: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """list of (D, T_u) -> frame-major (sum T, D) float64 + int64 offsets[B+1]."""
    lens = [f.shape[1] for f in features]
    offs = np.zeros(len(features) + 1, dtype=np.int64)
    offs[1:] = np.cumsum(lens)
    X = np.concatenate([np.asarray(f, dtype=np.float64).T for f in features], axis=0)
    return np.ascontiguousarray(X), offs


def num_threads() -> int:
    return int(lib().orc_num_threads())


# --------------------------------------------------------------------------- #
# per-function wrappers (argument meaning = the reference method of the same name)
# --------------------------------------------------------------------------- #
def init_parameters(features, N: int, var_floor_factor: float = 0.001):
    X, offs = pack(features)
    D = X.shape[1]
    S = N + 2
    gm = np.zeros(D); gc = np.zeros((D, D)); A = np.zeros((S, S))
    mean = np.zeros((S, D)); cov = np.zeros((S, D, D))
    lib().orc_init_parameters(_d(X), offs.ctypes.data_as(_lp), C.c_int(len(features)), C.c_int(D),
                              C.c_int(N), C.c_double(var_floor_factor), _d(gm), _d(gc), _d(A),
                              _d(mean), _d(cov))
    return gm, gc, A, mean, cov


def emission_sapr(feature_DT, mean, cov) -> np.ndarray:
    X = _f64(np.asarray(feature_DT).T)
    T, D = X.shape
    S = mean.shape[0]
    E = np.empty((T, S))
    mean = _f64(mean); cov = _f64(cov)
    lib().orc_emission_sapr(_d(X), C.c_int(T), C.c_int(D), C.c_int(S), _d(mean), _d(cov), _d(E))
    return E


def emission_diag(feature_DT, mean, var, all_emit: bool = False) -> np.ndarray:
    X = _f64(np.asarray(feature_DT).T)
    T, D = X.shape
    S = mean.shape[0]
    E = np.empty((T, S))
    mean = _f64(mean); var = _f64(var)
    lib().orc_emission_diag(_d(X), C.c_int(T), C.c_int(D), C.c_int(S), _d(mean), _d(var),
                            C.c_int(int(all_emit)), _d(E))
    return E


def forward(E, A):
    E = _f64(E); A = _f64(A)
    T, S = E.shape
    alpha = np.empty((T, S)); scale = C.c_double(0.0)
    lib().orc_forward(_d(E), C.c_int(T), C.c_int(S), _d(A), _d(alpha), C.byref(scale))
    return alpha, scale.value


def backward(E, A, scale: float):
    E = _f64(E); A = _f64(A)
    T, S = E.shape
    beta = np.empty((T, S))
    lib().orc_backward(_d(E), C.c_int(T), C.c_int(S), _d(A), C.c_double(scale), _d(beta))
    return beta


def gamma(alpha, beta):
    alpha = _f64(alpha); beta = _f64(beta)
    T, S = alpha.shape
    g = np.empty((T, S))
    lib().orc_gamma(_d(alpha), _d(beta), C.c_int(T), C.c_int(S), _d(g))
    return g


def xi(alpha, beta, E, A):
    alpha = _f64(alpha); beta = _f64(beta); E = _f64(E); A = _f64(A)
    T, S = alpha.shape
    x = np.zeros((max(T - 1, 0), S, S))
    lib().orc_xi(_d(alpha), _d(beta), _d(E), C.c_int(T), C.c_int(S), _d(A), _d(x))
    return x


def update_A(agg_xi, agg_gamma, A):
    A = _f64(A).copy(); agg_xi = _f64(agg_xi); agg_gamma = _f64(agg_gamma)
    lib().orc_update_A(_d(agg_xi), _d(agg_gamma), C.c_int(A.shape[0]), _d(A))
    return A


def update_B(features, gamma_per_seq, global_cov, var_floor_factor: float):
    X, offs = pack(features)
    D = X.shape[1]
    G = _f64(np.concatenate(gamma_per_seq, axis=0))
    S = G.shape[1]
    mean = np.zeros((S, D)); cov = np.zeros((S, D, D))
    gc = _f64(global_cov)
    lib().orc_update_B(_d(X), offs.ctypes.data_as(_lp), C.c_int(len(features)), C.c_int(D), C.c_int(S),
                       _d(G), _d(gc), C.c_double(var_floor_factor), _d(mean), _d(cov))
    return mean, cov


def decode(E, A, T_eff: int | None = None):
    """custom_hmm.decode on a given emission matrix; T_eff = rows walked (D3)."""
    E = _f64(E); A = _f64(A)
    T, S = E.shape
    Te = T if T_eff is None else int(T_eff)
    assert 1 <= Te <= T
    path = np.zeros(Te, dtype=np.int32)
    sc = lib().orc_decode(_d(E), C.c_int(Te), C.c_int(S), _d(A), path.ctypes.data_as(_ip))
    return float(sc), path


def baum_welch(features, N, A, mean, cov, global_cov, emission_mode: int, max_iter: int = 15,
               tol: float = 1e-4, var_floor_factor: float = 0.001):
    """Returns (history, A, mean, cov) after training; inputs are not modified."""
    X, offs = pack(features)
    D = X.shape[1]
    A = _f64(A).copy(); mean = _f64(mean).copy(); cov = _f64(cov).copy(); gc = _f64(global_cov)
    hist = np.zeros(max_iter)
    n = lib().orc_baum_welch(_d(X), offs.ctypes.data_as(_lp), C.c_int(len(features)), C.c_int(D),
                             C.c_int(N), C.c_int(emission_mode), C.c_int(max_iter), C.c_double(tol),
                             C.c_double(var_floor_factor), _d(gc), _d(A), _d(mean), _d(cov), _d(hist))
    return hist[:n].tolist(), A, mean, cov


# --------------------------------------------------------------------------- #
# batched legs (diag emission + sapr topology), X frame-major (sum T, D)
# --------------------------------------------------------------------------- #
def viterbi_batch(X, offsets, A, mean, var, first_frames: int = 0, want_scores=True, want_path=True,
                  nthreads: int = 0):
    X = _f64(X); offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    A = _f64(A); mean = _f64(mean); var = _f64(var)
    M, S, D = mean.shape
    B = len(offsets) - 1
    bw = np.zeros(B, dtype=np.int32); bs = np.zeros(B)
    sc = np.zeros((B, M)) if want_scores else None
    bp = np.zeros(int(offsets[-1]), dtype=np.int32) if want_path else None
    lib().orc_viterbi_batch(_d(X), offsets.ctypes.data_as(_lp), C.c_int(B), C.c_int(D), C.c_int(S),
                            C.c_int(M), _d(A), _d(mean), _d(var), C.c_int(first_frames),
                            bw.ctypes.data_as(_ip), _d(bs), _d(sc) if want_scores else None,
                            bp.ctypes.data_as(_ip) if want_path else None, C.c_int(nthreads))
    return bw, bs, sc, bp


def stats_stride(S: int, D: int) -> int:
    return 3 * S + 2 * S * D


def estep_batch(X, offsets, model_of_utt, A, mean, var, nthreads: int = 0):
    X = _f64(X); offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    mou = np.ascontiguousarray(model_of_utt, dtype=np.int32)
    A = _f64(A); mean = _f64(mean); var = _f64(var)
    M, S, D = mean.shape
    B = len(offsets) - 1
    stats = np.zeros((M, stats_stride(S, D))); ll = np.zeros(B)
    lib().orc_estep_batch(_d(X), offsets.ctypes.data_as(_lp), C.c_int(B), C.c_int(D), C.c_int(S),
                          C.c_int(M), mou.ctypes.data_as(_ip), _d(A), _d(mean), _d(var), _d(stats),
                          _d(ll), C.c_int(nthreads))
    return stats, ll


# --------------------------------------------------------------------------- #
# hmmlearn-style restatement (parity unpinned, see sapr_oracle.c header)
# --------------------------------------------------------------------------- #
def _logz(a):
    with np.errstate(divide="ignore"):
        return np.log(_f64(a))


def hl_forward(lf, startprob, transmat):
    lf = _f64(lf); T, S = lf.shape
    fwd = np.empty((T, S)); lpi = _logz(startprob); lA = _logz(transmat)
    lp = lib().orc_hl_forward(_d(lf), C.c_int(T), C.c_int(S), _d(lpi), _d(lA), _d(fwd))
    return float(lp), fwd


def hl_backward(lf, transmat):
    lf = _f64(lf); T, S = lf.shape
    bwd = np.empty((T, S)); lA = _logz(transmat)
    lib().orc_hl_backward(_d(lf), C.c_int(T), C.c_int(S), _d(lA), _d(bwd))
    return bwd


def hl_viterbi(lf, startprob, transmat):
    lf = _f64(lf); T, S = lf.shape
    path = np.zeros(T, dtype=np.int32); lpi = _logz(startprob); lA = _logz(transmat)
    lp = lib().orc_hl_viterbi(_d(lf), C.c_int(T), C.c_int(S), _d(lpi), _d(lA), path.ctypes.data_as(_ip))
    return float(lp), path


def hl_estep(X, offsets, startprob, transmat, mean, var):
    X = _f64(X); offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    mean = _f64(mean); var = _f64(var); sp = _f64(startprob); tm = _f64(transmat)
    S, D = mean.shape
    B = len(offsets) - 1
    start = np.zeros(S); trans = np.zeros((S, S)); post = np.zeros(S)
    obs = np.zeros((S, D)); obs2 = np.zeros((S, D)); ll = np.zeros(B)
    tot = lib().orc_hl_estep(_d(X), offsets.ctypes.data_as(_lp), C.c_int(B), C.c_int(D), C.c_int(S),
                             _d(sp), _d(tm), _d(mean), _d(var), _d(start), _d(trans), _d(post),
                             _d(obs), _d(obs2), _d(ll))
    return dict(logprob=float(tot), start=start, trans=trans, post=post, obs=obs, obs2=obs2, loglik=ll)


def hl_mstep(st, startprob, transmat, covars_prior=1e-2, covars_weight=1.0):
    """hmmlearn 0.3.3 _do_mstep for params='stmc', diag."""
    sp = np.where(np.asarray(startprob) == 0, 0.0, np.maximum(st["start"], 0))
    sp = sp / sp.sum()
    tm = np.where(np.asarray(transmat) == 0, 0.0, np.maximum(st["trans"], 0))
    rs = tm.sum(axis=1, keepdims=True)
    rs[rs == 0] = 1.0
    tm = tm / rs
    denom = st["post"][:, None]
    means = st["obs"] / denom
    num = covars_prior + st["obs2"] - 2 * means * st["obs"] + means ** 2 * denom
    cv = num / np.maximum(denom + max(covars_weight - 1, 0), 1e-5)
    return sp, tm, means, cv


class OracleHMM:
    """Mirror of the reference ``custom_hmm.HMM`` surface on top of the C oracle,
    so parity tests read like the reference's own tests.  ``emission`` selects
    "sapr" (as written, D1/D2) or "diag" (Rung-1 ladder)."""

    def __init__(self, num_states, num_obs, feature_set=None, model_name=None,
                 var_floor_factor=0.001, emission="sapr"):
        assert num_states > 0 and num_obs > 0
        self.num_states, self.num_obs = num_states, num_obs
        self.total_states = num_states + 2
        self.model_name, self.var_floor_factor, self.emission = model_name, var_floor_factor, emission
        self.pi = np.zeros(self.total_states); self.pi[0] = 1.0
        if feature_set is not None:
            assert all(f.shape[0] == num_obs for f in feature_set)
            gm, gc, A, mean, cov = init_parameters(feature_set, num_states, var_floor_factor)
            self.global_mean, self.global_covariance, self.A = gm, gc, A
            self.B = {"mean": mean, "covariance": cov}

    def compute_emission_matrix(self, features):
        if self.emission == "sapr":
            return emission_sapr(features, self.B["mean"], self.B["covariance"])
        var = np.ascontiguousarray(np.diagonal(self.B["covariance"], axis1=1, axis2=2))
        return emission_diag(features, self.B["mean"], var)

    def forward(self, E):
        return forward(E, self.A)

    def backward(self, E, scale):
        return backward(E, self.A, scale)

    def compute_gamma(self, alpha, beta):
        return gamma(alpha, beta)

    def compute_xi(self, alpha, beta, E):
        return xi(alpha, beta, E, self.A)

    def update_A(self, agg_xi, agg_gamma):
        self.A = update_A(agg_xi, agg_gamma, self.A)

    def update_B(self, features_list, gamma_per_seq):
        m, c = update_B(features_list, gamma_per_seq, self.global_covariance, self.var_floor_factor)
        self.B = {"mean": m, "covariance": c}

    def baum_welch(self, features_list, max_iter=15, tol=1e-4):
        hist, A, mean, cov = baum_welch(features_list, self.num_states, self.A, self.B["mean"],
                                        self.B["covariance"], self.global_covariance,
                                        0 if self.emission == "sapr" else 1, max_iter, tol,
                                        self.var_floor_factor)
        self.A, self.B = A, {"mean": mean, "covariance": cov}
        return hist

    def decode(self, features):
        T_eff = features.shape[0]            # custom_hmm.py:466
        E = self.compute_emission_matrix(features)   # raises for the wrong orientation, like the reference
        if T_eff > E.shape[0]:
            raise IndexError("index out of bounds (reference: T_frames < D)")
        sc, path = decode(E, self.A, T_eff)
        return sc, [int(p) for p in path]
