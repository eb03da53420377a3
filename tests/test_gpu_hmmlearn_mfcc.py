"""hmmlearn-style class and the MFCC kernel against the CPU oracle restatements (both PARITY UNPINNED by
the reference: hmmlearn / librosa are not vendored and cannot be installed here)."""
import numpy as np
import pytest

from conftest import assert_close, split_features
from oracle import oracle as orc
from oracle import mfcc_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _setup(g, w=0):
    from sapr_b200.hmmlearn_hmm import GaussianHMM
    feats = [f for f, l in zip(split_features(g), g["labels"]) if l == w]
    S = 10
    rng = np.random.default_rng(3)
    means = g["means"][w].copy()
    means[0] = means[1] + 0.1 * rng.standard_normal(means.shape[1])
    means[-1] = means[-2] + 0.1 * rng.standard_normal(means.shape[1])
    var = g["var"][w].copy()
    var[0] = var[1]; var[-1] = var[-2]
    tm = g["A"][w].copy()
    sp = np.zeros(S); sp[0] = 1.0
    model = GaussianHMM(n_components=S, covariance_type="diag", n_iter=3, params="stmc", implementation="log",
                        min_covar=0.01, init_params="")
    model.means_, model.covars_, model.transmat_, model.startprob_ = means, var, tm, sp
    X = np.concatenate([f.T for f in feats], axis=0)
    lengths = [f.shape[1] for f in feats]
    return model, X, lengths, (means, var, tm, sp)


def test_score_decode_vs_oracle(cuda, rung1_d13):
    model, X, lengths, (means, var, tm, sp) = _setup(rung1_d13)
    offs = np.concatenate([[0], np.cumsum(lengths)])
    tot, tot_v, paths = 0.0, 0.0, []
    for a, b in zip(offs[:-1], offs[1:]):
        lf = orc.emission_diag(X[a:b].T, means, var, all_emit=True)
        tot += orc.hl_forward(lf, sp, tm)[0]
        lp, p = orc.hl_viterbi(lf, sp, tm)
        tot_v += lp; paths.append(p)
    assert abs(model.score(X, lengths) - tot) <= 1e-10 * abs(tot)
    lp, path = model.decode(X, lengths)
    assert abs(lp - tot_v) <= 1e-10 * abs(tot_v)
    assert np.array_equal(path, np.concatenate(paths))
    assert np.array_equal(model.predict(X, lengths), path)
    assert model.covars_.shape == (10, 13, 13)      # hmmlearn getter expands 'diag' to full matrices


def test_large_ergodic_score_decode_vs_oracle(cuda):
    """BASELINE cfg 4 shape at test size: N = 256 fully connected states (rows ~ Dirichlet(1)), D = 39 -- score and
    Viterbi decode through the one-CTA-per-utterance kernels against the oracle restatement, then one Baum-Welch
    iteration (CTA-per-utterance backward / posteriors / transition statistics) against the oracle's E-step + M-step."""
    from sapr_b200.hmmlearn_hmm import GaussianHMM
    rng = np.random.default_rng(4)
    S, D = 256, 39
    means = 2.0 * rng.standard_normal((S, D)); var = rng.uniform(0.5, 1.5, (S, D)) ** 2
    tm = rng.dirichlet(np.ones(S), size=S); sp = rng.dirichlet(np.ones(S))
    lengths = [37, 5, 1, 64]
    states = rng.integers(0, S, size=sum(lengths))
    X = (means[states] + np.sqrt(var[states]) * rng.standard_normal((sum(lengths), D))).astype(np.float32)
    model = GaussianHMM(n_components=S, covariance_type="diag", n_iter=1, init_params="")
    model.means_, model.covars_, model.transmat_, model.startprob_ = means, var, tm, sp
    offs = np.concatenate([[0], np.cumsum(lengths)])
    tot, tot_v, paths = 0.0, 0.0, []
    for a, b in zip(offs[:-1], offs[1:]):
        lf = orc.emission_diag(X[a:b].T, means, var, all_emit=True)
        tot += orc.hl_forward(lf, sp, tm)[0]
        lp, p = orc.hl_viterbi(lf, sp, tm)
        tot_v += lp; paths.append(p)
    assert abs(model.score(X, lengths) - tot) <= 1e-10 * abs(tot)
    lp, path = model.decode(X, lengths)
    assert abs(lp - tot_v) <= 1e-10 * abs(tot_v)
    assert np.array_equal(path, np.concatenate(paths))
    st = orc.hl_estep(X, offs.astype(np.int64), sp, tm, means, var)
    sp2, tm2, means2, var2 = orc.hl_mstep(st, sp, tm)
    model.tol = -np.inf
    model.fit(X, lengths)
    assert_close(np.array(model.monitor_.history), np.array([st["logprob"]]), 1e-10, what="history")
    occupied = st["post"] > 1e-3                                   # states nobody visited keep prior-dominated values
    assert_close(model.means_[occupied], means2[occupied], 1e-8, what="means")
    assert_close(model._covars[occupied], var2[occupied], 1e-7, what="covars")
    assert_close(model.transmat_, tm2, 1e-8, atol=1e-12, what="transmat")
    assert_close(model.startprob_, sp2, 1e-10, atol=1e-15, what="startprob")
    big = GaussianHMM(n_components=1025, covariance_type="diag", n_iter=1, init_params="")
    big.means_, big.covars_ = np.zeros((1025, D)), np.ones((1025, D))
    big.transmat_, big.startprob_ = np.full((1025, 1025), 1 / 1025), np.full(1025, 1 / 1025)
    with pytest.raises(Exception):                                 # beyond 1024 states: refused loudly
        big.score(X, lengths)


@pytest.mark.parametrize("S,D,peaked", [(256, 39, False), (128, 13, False), (64, 39, False), (192, 39, True), (256, 39, True)])
def test_ergodic_tensor_core_score_vs_float64(cuda, S, D, peaked):
    """BASELINE cfg 4: the fp32 tensor-core scaled forward (csrc/ergodic_tc.cu) against the float64 log-domain kernel on
    the same device batch -- more than one 128-utterance tile, ragged lengths down to one frame, a sparse row in the
    transition matrix.  ``peaked``: near-deterministic transitions scored on frames drawn from unrelated states, so the
    predicted mass on the observed state collapses and the kernel's exact-renormalisation path runs (jumps the model
    gives probability zero are left out: the fp16 operands drop posterior mass below 2^-39 of a frame's total, so a
    likelihood carried only by such tails is outside this kernel's contract, see include/sapr_b200.h).  Tolerance (include/sapr_b200.h): fp16 transition weights (2^-12 relative, one entry dominates a
    peaked sum) + fp16 hi/lo emission operands give |d logP| <= 5e-6 |logP| + 2e-4 T."""
    from sapr_b200.hmmlearn_hmm import GaussianHMM
    rng = np.random.default_rng(S + D)
    means = 2.0 * rng.standard_normal((S, D)); var = rng.uniform(0.5, 1.5, (S, D)) ** 2
    tm = rng.dirichlet(np.ones(S), size=S); sp = rng.dirichlet(np.ones(S))
    if peaked:
        tm = 1e-4 * tm + (1 - 1e-4) * (0.6 * np.eye(S) + 0.4 * np.roll(np.eye(S), 1, axis=1))
    tm[3] = 0.0; tm[3, 3] = 0.7; tm[3, 4] = 0.3                    # a left-to-right style row inside the dense matrix
    lengths = [int(x) for x in rng.integers(1, 90, size=150)]
    lengths[0], lengths[1], lengths[140] = 1, 2, 120
    states = np.empty(sum(lengths), dtype=np.int64)
    o = 0
    for T in lengths:                                             # hidden paths drawn from the chain itself
        st = rng.choice(S, p=sp)
        for t in range(T):
            states[o + t] = st
            st = rng.integers(0, S) if peaked and t % 7 == 3 and st != 3 else rng.choice(S, p=tm[st])
        o += T
    X = (means[states] + np.sqrt(var[states]) * rng.standard_normal((sum(lengths), D))).astype(np.float32)
    model = GaussianHMM(n_components=S, covariance_type="diag", n_iter=1, init_params="")
    model.means_, model.covars_, model.transmat_, model.startprob_ = means, var, tm, sp
    ref = model.score_each(X, lengths).cpu().numpy()
    got = model.score_each(X, lengths, precision="tc").cpu().numpy()
    assert np.isfinite(got).all()
    err = np.abs(got - ref)
    tol = 5e-6 * np.abs(ref) + 2e-4 * np.asarray(lengths)
    print("ergodic tc: max err", err.max(), "max err/tol", (err / tol).max())
    assert (err <= tol).all(), (err.max(), (err / tol).max())
    assert abs(model.score(X, lengths, precision="tc") - ref.sum()) <= 5e-6 * abs(ref.sum())


@pytest.mark.parametrize("S,D", [(256, 39), (64, 13)])
def test_ergodic_tensor_core_score_vs_oracle(cuda, S, D):
    """sapr_ergodic_score (tensor-core forward, csrc/ergodic_tc.cu) checked DIRECTLY against the CPU oracle's
    hmmlearn-style forward (orc.emission_diag + orc.hl_forward), not against another kernel of this repository.
    Same tolerance as the kernel's contract in include/sapr_b200.h: |d logP| <= 5e-6 |logP| + 2e-4 T."""
    from sapr_b200.hmmlearn_hmm import GaussianHMM
    rng = np.random.default_rng(1000 + S)
    means = 2.0 * rng.standard_normal((S, D)); var = rng.uniform(0.5, 1.5, (S, D)) ** 2
    tm = rng.dirichlet(np.ones(S), size=S); sp = rng.dirichlet(np.ones(S))
    lengths = [1, 2, 17, 40, 63, 5, 33, 90] + [int(x) for x in rng.integers(3, 50, size=140)]   # > one 128-row tile
    states = rng.integers(0, S, size=sum(lengths))
    X = (means[states] + np.sqrt(var[states]) * rng.standard_normal((sum(lengths), D))).astype(np.float32)
    model = GaussianHMM(n_components=S, covariance_type="diag", n_iter=1, init_params="")
    model.means_, model.covars_, model.transmat_, model.startprob_ = means, var, tm, sp
    got = model.score_each(X, lengths, precision="tc").cpu().numpy()
    offs = np.concatenate([[0], np.cumsum(lengths)])
    ref = np.array([orc.hl_forward(orc.emission_diag(X[a:b].T, means, var, all_emit=True), sp, tm)[0]
                    for a, b in zip(offs[:-1], offs[1:])])
    err = np.abs(got - ref)
    tol = 5e-6 * np.abs(ref) + 2e-4 * np.asarray(lengths)
    assert np.isfinite(got).all() and (err <= tol).all(), (err.max(), (err / tol).max())


def test_fit_vs_oracle(cuda, rung1_d13):
    model, X, lengths, (means, var, tm, sp) = _setup(rung1_d13, w=2)
    offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    hist = []
    for _ in range(3):
        st = orc.hl_estep(X, offs, sp, tm, means, var)
        hist.append(st["logprob"])
        sp, tm, means, var = orc.hl_mstep(st, sp, tm)
    model.tol = -np.inf                        # run all three iterations
    model.fit(X, lengths)
    assert_close(np.array(model.monitor_.history), np.array(hist), 1e-10, what="history")
    assert_close(model.means_, means, 1e-9, what="means")
    assert_close(model._covars, var, 1e-8, what="covars")
    assert_close(model.transmat_, tm, 1e-9, atol=1e-12, what="transmat")
    assert_close(model.startprob_, sp, 1e-12, what="startprob")


def test_hmmlearn_wrapper(cuda):
    from sapr_b200 import synth
    from sapr_b200.hmmlearn_hmm import HMMLearnModel
    feats, labels, _, _ = synth.make_corpus(22, 11, 8, 13, 40, 60, seed=11)
    w = HMMLearnModel(num_states=8, model_name="heed", n_iter=2, feature_set=feats)
    X = np.concatenate([f.T for f in feats], axis=0)
    assert_close(w.global_mean, X.astype(np.float64).mean(axis=0), 1e-10, what="global mean")
    assert_close(w.global_cov, X.astype(np.float64).var(axis=0), 1e-9, what="global var")
    model, ll = w.fit([f for f, l in zip(feats, labels) if l == 0])
    assert np.isfinite(ll) and len(model.monitor_.history) == 2


@pytest.mark.parametrize("which", ["librosa", "cfg5"])
def test_mfcc_vs_oracle(cuda, which):
    from sapr_b200 import mfcc_extract as mx
    rng = np.random.default_rng(0)
    sr = 22050 if which == "librosa" else 16000
    p = mx.librosa_params(sr) if which == "librosa" else mx.cfg5_params()
    sigs = []
    for n in (sr, int(0.37 * sr), 5000):
        t = np.arange(n) / sr
        y = sum(a * np.sin(2 * np.pi * f * t + ph) for a, f, ph in zip((0.5, 0.3, 0.2), rng.uniform(200, 3000, 3), rng.uniform(0, 6, 3)))
        sigs.append((y + 0.01 * rng.standard_normal(n)).astype(np.float32))
    audio = cuda.as_tensor(np.concatenate(sigs), device="cuda")
    so = np.concatenate([[0], np.cumsum([len(s) for s in sigs])]).astype(np.int64)
    feats, fo = mx.mfcc_batch(audio, so, p)
    F = feats.cpu().numpy()
    for u, y in enumerate(sigs):
        ref = mfcc_oracle.mfcc(y, p.sample_rate, p.n_fft, p.win_length, p.hop_length, p.n_mels, p.n_mfcc, p.center,
                               p.mel_slaney, p.log_db, p.top_db, p.preemph, p.fmin, p.fmax)
        got = F[fo[u]:fo[u + 1], :p.n_mfcc].T
        assert got.shape == ref.shape
        if which == "librosa":
            assert ref.shape == (13, 1 + len(y) // 220)          # tests/test_mfcc_extract.py:31-34 + Appendix C
        # fp32 FFT/log vs float64: absolute tolerance on dB/log-scaled cepstra
        assert np.max(np.abs(got - ref)) < (5e-2 if which == "librosa" else 5e-3), np.max(np.abs(got - ref))
    one = mx.mfcc_from_samples(sigs[0], sr, p)
    assert one.shape[0] == 13 and one.dtype == np.float32


def test_cfg5_audio_to_words_on_device(cuda):
    """BASELINE cfg 5 at test size: synthetic 16 kHz audio -> fused MFCC kernel (25 ms / 10 ms Hamming, 512-pt FFT, 26 mel,
    DCT-13, pre-emphasis) -> batched Viterbi against 11 word models, features never leaving the GPU.  The recognised
    words / paths must equal the oracle's decoding of the very same feature tensor (float64 mode: bit-exact)."""
    from sapr_b200 import engine, mfcc_extract as mx, synth
    rng = np.random.default_rng(9)
    sr, B = 16000, 24
    sigs = []
    for u in range(B):
        n = int(rng.integers(sr // 2, sr))
        t = np.arange(n) / sr
        y = sum(a * np.sin(2 * np.pi * f * t + ph) for a, f, ph in zip((0.5, 0.3, 0.2), rng.uniform(200, 3000, 3), rng.uniform(0, 6, 3)))
        sigs.append((y + 0.05 * rng.standard_normal(n)).astype(np.float32))
    audio = cuda.as_tensor(np.concatenate(sigs), device="cuda")
    so = np.concatenate([[0], np.cumsum([len(s) for s in sigs])]).astype(np.int64)
    feats, fo = mx.mfcc_batch(audio, so, mx.cfg5_params())
    assert feats.shape[1] == 16 and fo[-1] == feats.shape[0]
    # word models around the corpus statistics (any fixed model set exercises the path)
    F = feats[:, :13].cpu().numpy().astype(np.float64)
    mu, sd = F.mean(0), F.std(0) + 1e-3
    means = np.zeros((11, 10, 13)); var = np.ones((11, 10, 13))
    means[:, 1:-1] = mu + sd * rng.standard_normal((11, 8, 13)); var[:, 1:-1] = (sd * rng.uniform(0.7, 1.3, (11, 8, 13))) ** 2
    A = synth.truth_models(np.zeros((11, 8, 13)), np.ones((11, 8, 13)), 0.9)[0]
    m = engine.WordModels(11, 8, 13)
    m.set(means, var, A)
    batch = engine.PackedBatch(feats, cuda.as_tensor(fo, device="cuda"), 13, fo)
    bw, bs, sc, bp = orc.viterbi_batch(F, fo, A, means, var)
    o64 = m.viterbi(batch, None, engine.FP64, 0, want_scores=True)
    assert np.array_equal(o64["best_word"].cpu().numpy(), bw)
    assert np.array_equal(o64["path"].cpu().numpy().astype(np.int32), bp)
    o32 = m.viterbi(batch, None, engine.FP32, 0, want_scores=True)      # tensor-core path, D = 13
    assert_close(o32["scores"].cpu().numpy(), sc, 1e-6, what="scores32")
    assert np.mean(o32["best_word"].cpu().numpy() == bw) >= 0.95


def test_large_batch_thread_per_utterance_vs_oracle(cuda, rung1_d13):
    """2500 short utterances: enough for the thread-per-utterance kernels (small batches run one CTA per utterance) and for
    many chunks in the frame-parallel observation statistics; score / decode / one fit iteration against the oracle."""
    model, X0, lengths0, (means, var, tm, sp) = _setup(rung1_d13)
    rng = np.random.default_rng(11)
    lengths = rng.integers(10, 19, size=2500)
    offs = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    X = (means[rng.integers(1, 9, size=int(offs[-1]))] + np.sqrt(var[1]) * rng.standard_normal((int(offs[-1]), 13))).astype(np.float32).astype(np.float64)
    tot, tot_v, paths = 0.0, 0.0, []
    for a, b in zip(offs[:-1], offs[1:]):
        lf = orc.emission_diag(X[a:b].T, means, var, all_emit=True)
        tot += orc.hl_forward(lf, sp, tm)[0]
        lp, p = orc.hl_viterbi(lf, sp, tm)
        tot_v += lp; paths.append(p)
    assert abs(model.score(X, list(lengths)) - tot) <= 1e-10 * abs(tot)
    lp, path = model.decode(X, list(lengths))
    assert abs(lp - tot_v) <= 1e-10 * abs(tot_v)
    assert np.array_equal(path, np.concatenate(paths))
    st = orc.hl_estep(X, offs, sp, tm, means, var)
    sp2, tm2, means2, var2 = orc.hl_mstep(st, sp, tm)
    model.n_iter = 1
    model.fit(X, list(lengths))
    assert_close(np.array(model.monitor_.history), np.array([st["logprob"]]), 1e-10, what="history")
    assert_close(model.means_, means2, 1e-9, what="means")
    assert_close(model._covars, var2, 1e-8, what="covars")
    assert_close(model.transmat_, tm2, 1e-9, atol=1e-12, what="transmat")
