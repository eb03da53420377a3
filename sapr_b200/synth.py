"""synth-v1: the synthetic isolated-word corpus every test and benchmark uses.

The reference ships no data (assignment2/.gitignore:2-8); its models are 11 words x
8-state left-to-right HMMs on 13-dim MFCCs (assignment2/train.py:89-92,129-131).
synth-v1 draws, per word and state, a diagonal Gaussian and emits
left-to-right utterances from it.  Host (numpy) generation is used for parity
tests; ``device_corpus`` generates benchmark-sized corpora directly in HBM with
torch (tensor plumbing only) so their content does not depend on the GPU count.
"""
from __future__ import annotations

import numpy as np

VOCAB = ["heed", "hid", "head", "had", "hard", "hud", "hod", "hoard", "hood", "whod", "heard"]

# per-coefficient scale of real MFCCs recorded by the reference's own test run
# (assignment2/pytest_results/initialization_results.txt:28-41)
MFCC_VAR = np.array([27861.58, 4428.24, 1204.46, 1009.87, 645.36, 412.71, 239.74, 162.52,
                     184.36, 125.25, 108.62, 108.56, 87.97])


def dim_scale(D: int) -> np.ndarray:
    return np.sqrt(MFCC_VAR) if D == 13 else np.ones(D)


def ground_truth(M: int, N: int, D: int, seed: int):
    """Per word/state means and standard deviations: mu ~ N(0, 2^2) s_d, sigma ~ U(.5, 1.5) s_d."""
    rng = np.random.default_rng(seed)
    s = dim_scale(D)
    mu = rng.normal(0.0, 2.0, size=(M, N, D)) * s
    sd = rng.uniform(0.5, 1.5, size=(M, N, D)) * s
    return mu, sd


def make_corpus(B: int, M: int, N: int, D: int, T_lo: int, T_hi: int, seed: int):
    """Returns (features, labels, mu, sd): features is a list of (D, T_u) float32 arrays --
    the orientation the reference stores (assignment2/mfcc_extract.py:41-42)."""
    mu, sd = ground_truth(M, N, D, seed)
    rng = np.random.default_rng(seed + 7919)
    feats, labels = [], []
    for u in range(B):
        w = u % M
        T = int(rng.integers(T_lo, T_hi + 1))
        cuts = np.sort(rng.choice(np.arange(1, T), size=N - 1, replace=False))
        state = np.searchsorted(cuts, np.arange(T), side="right")
        x = mu[w, state] + sd[w, state] * rng.standard_normal((T, D))
        feats.append(np.ascontiguousarray(x.T.astype(np.float32)))
        labels.append(w)
    return feats, np.asarray(labels, dtype=np.int32), mu, sd


def truth_models(mu: np.ndarray, sd: np.ndarray, a_self: float = 0.9):
    """Word models in the reference's parameterisation: S = N+2 states with a
    non-emitting entry (0) and exit (S-1) (assignment2/custom_hmm.py:94-116)."""
    M, N, D = mu.shape
    S = N + 2
    means = np.zeros((M, S, D)); var = np.ones((M, S, D)); A = np.zeros((M, S, S))
    means[:, 1:-1] = mu
    var[:, 1:-1] = sd ** 2
    A[:, 0, 1] = 1.0
    for i in range(1, N + 1):
        A[:, i, i] = a_self
        A[:, i, i + 1] = 1.0 - a_self
    A[:, S - 1, S - 1] = 1.0
    return A, means, var


def pack_frame_major(features, d_pad: int | None = None):
    """list of (D, T_u) -> float32 (sum T, Dpad) frame-major + int64 offsets[B+1].
    Dpad = D rounded up to a multiple of 4 floats (16-byte rows for 128-bit loads / TMA)."""
    D = features[0].shape[0]
    dp = d_pad or ((D + 3) // 4 * 4)
    lens = [f.shape[1] for f in features]
    offs = np.zeros(len(features) + 1, dtype=np.int64)
    offs[1:] = np.cumsum(lens)
    X = np.zeros((int(offs[-1]), dp), dtype=np.float32)
    for f, o in zip(features, offs[:-1]):
        X[o:o + f.shape[1], :D] = f.T
    return X, offs


def device_corpus(B: int, M: int, N: int, D: int, T: int, seed: int, device, block: int = 8192):
    """Fixed-length corpus generated in HBM: returns (X[B*T, Dpad] f32, offsets[B+1] i64,
    labels[B] i32, mu, sd).  Content depends only on (seed, utterance id)."""
    import torch

    mu, sd = ground_truth(M, N, D, seed)
    dp = (D + 3) // 4 * 4
    mu_t = torch.tensor(mu, dtype=torch.float32, device=device)
    sd_t = torch.tensor(sd, dtype=torch.float32, device=device)
    X = torch.zeros((B * T, dp), dtype=torch.float32, device=device)
    labels = (torch.arange(B, device=device) % M).to(torch.int32)
    for b0 in range(0, B, block):
        nb = min(block, B - b0)
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + b0)
        # N-1 distinct cut points in [1, T-1]: rank of random keys
        keys = torch.rand((nb, T - 1), generator=g, device=device)
        cuts = keys.argsort(dim=1)[:, : N - 1] + 1
        cuts, _ = cuts.sort(dim=1)
        t_idx = torch.arange(T, device=device).view(1, T, 1)
        state = (t_idx >= cuts.view(nb, 1, N - 1)).sum(dim=2)          # (nb, T) in [0, N-1]
        w = labels[b0:b0 + nb].long().view(nb, 1).expand(nb, T)
        eps = torch.randn((nb, T, D), generator=g, device=device)
        x = mu_t[w, state] + sd_t[w, state] * eps
        X[b0 * T:(b0 + nb) * T, :D] = x.reshape(nb * T, D)
    offsets = torch.arange(B + 1, device=device, dtype=torch.int64) * T
    return X, offsets, labels, mu, sd
