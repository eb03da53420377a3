"""hmmlearn-style surface of the reference's second implementation (assignment2/hmmlearn_hmm.py).

``GaussianHMM`` mirrors the subset of ``hmmlearn.hmm.GaussianHMM`` the reference touches
(hmmlearn_hmm.py:27-43, :103-104; train.py:116-117; decoder.py:43; visualize.py:124): constructor kwargs,
settable ``means_ / covars_ / transmat_ / startprob_``, ``fit / score / decode / predict`` and
``monitor_.history``.  hmmlearn 0.3.3 itself is not vendored in the reference and not installable here, so
the arithmetic follows SURVEY.md Appendix B (parity unpinned by the reference); it runs on the DENSE-topology
float64 kernels of csrc/hmmlearn.cu.  ``HMMLearnModel`` is the reference's wrapper class, same signature.
"""
from __future__ import annotations

import logging
from collections import deque
from typing import List

import numpy as np

from ._lib import EMIT_DIAG, TOPO_DENSE, ptr
from .engine import PackedBatch, WordModels


def _torch():
    import torch
    return torch


class ConvergenceMonitor:
    """hmmlearn.base.ConvergenceMonitor: history of per-iteration total log-probability."""

    def __init__(self, tol, n_iter, verbose=False):
        self.tol, self.n_iter, self.verbose = tol, n_iter, verbose
        self.history = deque()
        self.iter = 0

    def _reset(self):
        self.iter = 0
        self.history.clear()

    def report(self, log_prob):
        if self.verbose:
            delta = log_prob - self.history[-1] if self.history else np.nan
            print(f"{self.iter + 1:>10d} {log_prob:>16.8f} {delta:>+16.8f}")
        self.history.append(log_prob)
        self.iter += 1

    @property
    def converged(self):
        return (self.iter == self.n_iter or
                (len(self.history) >= 2 and self.history[-1] - self.history[-2] < self.tol))


class GaussianHMM:
    def __init__(self, n_components=1, covariance_type="diag", min_covar=1e-3, startprob_prior=1.0,
                 transmat_prior=1.0, means_prior=0, means_weight=0, covars_prior=1e-2, covars_weight=1,
                 algorithm="viterbi", random_state=None, n_iter=10, tol=1e-2, verbose=False, params="stmc",
                 init_params="stmc", implementation="log"):
        if covariance_type != "diag":
            raise NotImplementedError("only covariance_type='diag' is built (the reference uses nothing else)")
        if implementation != "log":
            raise NotImplementedError("only implementation='log' is built (hmmlearn_hmm.py:32)")
        self.n_components, self.covariance_type, self.min_covar = n_components, covariance_type, min_covar
        self.startprob_prior, self.transmat_prior = startprob_prior, transmat_prior
        self.means_prior, self.means_weight = means_prior, means_weight
        self.covars_prior, self.covars_weight = covars_prior, covars_weight
        self.algorithm, self.random_state, self.n_iter, self.tol = algorithm, random_state, n_iter, tol
        self.verbose, self.params, self.init_params, self.implementation = verbose, params, init_params, implementation
        self.monitor_ = ConvergenceMonitor(tol, n_iter, verbose)
        self._covars = None
        self._dev = None

    # hmmlearn stores (S, D) for "diag" and the getter expands to (S, D, D)
    @property
    def covars_(self):
        return np.array([np.diag(c) for c in self._covars])

    @covars_.setter
    def covars_(self, v):
        v = np.asarray(v, dtype=np.float64)
        self._covars = np.array([np.diag(c) for c in v]) if v.ndim == 3 else v.copy()

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_dev"] = None
        st.pop("_dev_key", None)
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._dev = None
        self._dev_key = None

    def _check(self):
        sp = np.asarray(self.startprob_, dtype=np.float64)
        tm = np.asarray(self.transmat_, dtype=np.float64)
        S = self.n_components
        if sp.shape != (S,) or not np.isclose(sp.sum(), 1.0):
            raise ValueError("startprob_ must have length n_components and sum to 1")
        if tm.shape != (S, S) or not np.allclose(tm.sum(axis=1), 1.0):
            raise ValueError("transmat_ rows must sum to 1")
        self.means_ = np.asarray(self.means_, dtype=np.float64)
        if self._covars is None or self._covars.shape != self.means_.shape or (self._covars <= 0).any():
            raise ValueError("'diag' covars must be positive and shaped like means_")
        self.n_features = self.means_.shape[1]

    def _models(self) -> WordModels:
        self._check()
        S, D = self.means_.shape
        if self._dev is None or (self._dev.S, self._dev.D) != (S, D):
            self._dev = WordModels(1, S, D, EMIT_DIAG, TOPO_DENSE)
        # decode / score are called once per utterance and model by decoder.py: upload only when the parameters changed
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (self.means_, self._covars, self.transmat_, self.startprob_)]
        key = (id(self._dev),) + tuple(hash(a.tobytes()) for a in arrs)
        if getattr(self, "_dev_key", None) != key:
            self._dev.set(*arrs)
            self._dev_key = key
        return self._dev

    @staticmethod
    def _batch(X, lengths):
        return PackedBatch.from_frames(np.asarray(X), lengths)

    def score_each(self, X, lengths=None, precision="float64"):
        """Per-sequence log-probabilities (device float64 tensor).  ``precision="float64"`` is the log-domain
        verification mode; ``"tc"`` runs the fp32 tensor-core scaled forward of csrc/ergodic_tc.cu (dense models
        with 64/128/192/256 states, BASELINE cfg 4)."""
        torch = _torch()
        m = self._models()
        b = self._batch(X, lengths)
        lp = torch.zeros(b.B, dtype=torch.float64, device=b.X.device)
        if precision not in ("float64", "tc"):
            raise ValueError("precision must be 'float64' or 'tc'")
        if precision == "float64":
            m.ctx.check(m.lib.sapr_hl_score(m.ctx.h, m.h, 0, ptr(b.X), b.ldx, ptr(b.offsets), b.B, b.total_frames, ptr(lp)))
        else:
            m.ctx.check(m.lib.sapr_ergodic_score(m.ctx.h, m.h, 0, ptr(b.X), b.ldx, ptr(b.offsets), b.B, b.max_T, ptr(lp)))
        return lp

    def score(self, X, lengths=None, precision="float64"):
        return float(self.score_each(X, lengths, precision).cpu().numpy().sum())

    def decode(self, X, lengths=None, algorithm=None):
        torch = _torch()
        m = self._models()
        b = self._batch(X, lengths)
        lp = torch.zeros(b.B, dtype=torch.float64, device=b.X.device)
        path = torch.zeros(b.total_frames, dtype=torch.int32, device=b.X.device)
        m.ctx.check(m.lib.sapr_hl_decode(m.ctx.h, m.h, 0, ptr(b.X), b.ldx, ptr(b.offsets), b.B, b.total_frames,
                                         ptr(lp), ptr(path)))
        return float(lp.cpu().numpy().sum()), path.cpu().numpy().astype(np.int64)

    def predict(self, X, lengths=None):
        return self.decode(X, lengths)[1]

    def _estep(self, m, b):
        torch = _torch()
        S, D = self.means_.shape
        n = int(m.lib.sapr_hl_stats_len(S, D))
        stats = torch.zeros(n, dtype=torch.float64, device=b.X.device)
        lp = torch.zeros(b.B, dtype=torch.float64, device=b.X.device)
        m.ctx.check(m.lib.sapr_hl_estep(m.ctx.h, m.h, 0, ptr(b.X), b.ldx, ptr(b.offsets), b.B, b.total_frames,
                                        ptr(stats), ptr(lp)))
        st = stats.cpu().numpy()
        o = 0
        out = {}
        for name, size, shape in (("start", S, (S,)), ("trans", S * S, (S, S)), ("post", S, (S,)),
                                  ("obs", S * D, (S, D)), ("obs2", S * D, (S, D))):
            out[name] = st[o:o + size].reshape(shape); o += size
        return out, float(lp.cpu().numpy().sum())

    def _do_mstep(self, st):
        """hmmlearn 0.3.3 BaseHMM._do_mstep + GaussianHMM._do_mstep for 'diag' (SURVEY Appendix B)."""
        if "s" in self.params:
            sp = np.maximum(self.startprob_prior - 1 + st["start"], 0)
            sp = np.where(self.startprob_ == 0, 0, sp)
            self.startprob_ = sp / sp.sum()
        if "t" in self.params:
            tm = np.maximum(self.transmat_prior - 1 + st["trans"], 0)
            tm = np.where(self.transmat_ == 0, 0, tm)
            rs = tm.sum(axis=1, keepdims=True)
            rs[rs == 0] = 1
            self.transmat_ = tm / rs
        denom = st["post"][:, None]
        if "m" in self.params:
            self.means_ = (self.means_weight * self.means_prior + st["obs"]) / (self.means_weight + denom)
        if "c" in self.params:
            meandiff = self.means_ - self.means_prior
            c_n = (self.means_weight * meandiff ** 2 + st["obs2"] - 2 * self.means_ * st["obs"] + self.means_ ** 2 * denom)
            c_d = max(self.covars_weight - 1, 0) + denom
            self._covars = (self.covars_prior + c_n) / np.maximum(c_d, 1e-5)

    def fit(self, X, lengths=None):
        b = self._batch(X, lengths)
        self._check()
        self.monitor_ = ConvergenceMonitor(self.tol, self.n_iter, self.verbose)
        self.monitor_._reset()
        for _ in range(self.n_iter):
            m = self._models()
            st, logprob = self._estep(m, b)
            self._do_mstep(st)
            self.monitor_.report(logprob)
            if self.monitor_.converged:
                break
        return self


class _hmm_ns:
    GaussianHMM = GaussianHMM


hmm = _hmm_ns   # `from hmmlearn import hmm` look-alike used by the reference wrapper


class HMMLearnModel:
    """assignment2/hmmlearn_hmm.py:11-108, same constructor and methods; ``feature_set`` may be passed
    directly instead of being re-loaded from ./feature_set (hmmlearn_hmm.py:23)."""

    def __init__(self, num_states: int = 8, model_name: str = None, n_iter: int = 15, min_covar: float = 0.01,
                 feature_set=None):
        self.model_name = model_name
        self.num_states = num_states
        self.total_states = num_states + 2
        if feature_set is None:
            from .mfcc_extract import load_mfccs
            feature_set = load_mfccs("feature_set")
        self.all_features = feature_set
        self.global_mean = self.calc_global_mean(self.all_features)
        self.global_cov = self.calc_global_cov(self.all_features)
        self.model = GaussianHMM(n_components=self.total_states, covariance_type="diag", n_iter=n_iter, params="stmc",
                                 implementation="log", min_covar=min_covar, init_params="")
        self.model.means_ = np.tile(self.global_mean, (self.total_states, 1))
        self.model.covars_ = np.tile(self.global_cov, (self.total_states, 1))
        self.model.transmat_ = self.initialize_transmat()
        self.model.startprob_ = np.zeros(self.total_states)
        self.model.startprob_[0] = 1.0

    def initialize_transmat(self) -> np.ndarray:
        total_frames = sum(f.shape[1] for f in self.all_features)
        avg_frames_per_state = total_frames / len(self.all_features) / self.num_states
        aii = np.exp(-1 / (avg_frames_per_state - 1))
        transmat = np.zeros((self.total_states, self.total_states))
        transmat[0, 1] = 1.0
        for i in range(1, self.num_states + 1):
            transmat[i, i] = aii
            transmat[i, i + 1] = 1 - aii
        transmat[self.num_states + 1, self.num_states + 1] = 1.0
        return transmat

    def prepare_data(self, feature_set: List[np.ndarray]) -> np.ndarray:
        return np.concatenate([f.T for f in feature_set], axis=0)

    def _global_stats(self, feature_set):
        # global mean / population variance (hmmlearn_hmm.py:83-94) from the GPU flat-start sums
        from .engine import init_flat_start
        b = PackedBatch.from_features(feature_set)
        gmean, var, _, _ = init_flat_start(b, self.num_states, var_floor_factor=0.0)
        return gmean, var

    def calc_global_mean(self, feature_set):
        return self._global_stats(feature_set)[0]

    def calc_global_cov(self, feature_set):
        return self._global_stats(feature_set)[1]

    def fit(self, feature_set: List[np.ndarray]):
        logging.info(f"Training {self.model_name} HMM using hmmlearn in {self.model.n_iter} iterations...")
        X = self.prepare_data(feature_set)
        lengths = [f.shape[1] for f in feature_set]
        try:
            self.model.fit(X, lengths)
            log_likelihood = self.model.score(X, lengths)
            return self.model, log_likelihood
        except Exception as e:   # the reference swallows and returns None (hmmlearn_hmm.py:107-108)
            logging.error(f"Error occurred while training {self.model_name} HMM: {e}")
