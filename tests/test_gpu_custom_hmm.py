"""The drop-in ``custom_hmm.HMM`` class on the CUDA path against vectors produced by the reference's own
``custom_hmm.HMM`` (Rung 0: as written; Rung 1: diagonal emission), and the reference's own test
assertions (assignment2/tests/*.py) re-run against the drop-in class on a synthetic feature_set."""
import pickle

import numpy as np
import pytest

from conftest import assert_close, split_features

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def HMM():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from sapr_b200.custom_hmm import HMM
    return HMM


def _model(HMM, g, pre, **kw):
    h = HMM(int(g["N"]), int(g["D"]), **kw)
    h.A = g[pre + "A"].copy()
    h.B = {"mean": g[pre + "mean"].copy(), "covariance": g[pre + "cov"].copy()}
    h.global_covariance = g["init_global_cov"].copy()
    return h


def test_init_parameters(HMM, rung0):
    feats = split_features(rung0)
    h = HMM(8, 13, feats)
    assert_close(h.global_mean, rung0["init_global_mean"], 1e-12, what="global_mean")
    assert_close(h.global_covariance, rung0["init_global_cov"], 1e-11, what="global_cov")
    assert_close(h.A, rung0["init_A"], 1e-13, what="A")
    assert_close(h.B["mean"], rung0["init_mean"], 1e-12, what="mean")
    assert_close(h.B["covariance"], rung0["init_cov"], 1e-11, what="cov")
    with pytest.raises(AssertionError):
        HMM(0, 13)
    with pytest.raises(AssertionError):
        HMM(8, 13, [np.zeros((12, 30), dtype=np.float32)])


@pytest.mark.parametrize("pre,params,utt", [("flat_u0_", "init_", 0), ("trained_u1_", "bw_w0_k2_", 1)])
def test_per_function_as_written(HMM, rung0, pre, params, utt):
    feats = split_features(rung0)
    h = _model(HMM, rung0, params)
    E = h.compute_emission_matrix(feats[utt])
    assert_close(E, rung0[pre + "E"], 1e-9, what="E")
    Eg = rung0[pre + "E"]
    alpha, scale = h.forward(Eg)
    assert_close(alpha, rung0[pre + "alpha"], 1e-9, what="alpha")
    assert abs(scale - float(rung0[pre + "scale"])) <= 1e-9 * max(1.0, abs(scale))
    beta = h.backward(Eg, float(rung0[pre + "scale"]))
    assert_close(beta, rung0[pre + "beta"], 1e-9, what="beta")
    assert_close(h.compute_gamma(rung0[pre + "alpha"], rung0[pre + "beta"]), rung0[pre + "gamma"], 0, 1e-10, what="gamma")
    assert_close(h.compute_xi(rung0[pre + "alpha"], rung0[pre + "beta"], Eg), rung0[pre + "xi"], 0, 1e-10, what="xi")


def test_update_A_B(HMM, rung0):
    feats = split_features(rung0)
    h = _model(HMM, rung0, "bw_w0_k2_")
    h.update_A(rung0["upd_agg_xi"], rung0["upd_agg_gamma"])
    h.update_B([feats[1]], [rung0["trained_u1_gamma"]])
    assert_close(h.A, rung0["upd_A"], 1e-12, what="A")
    assert_close(h.B["mean"], rung0["upd_mean"], 1e-10, what="mean")
    assert_close(h.B["covariance"], rung0["upd_cov"], 1e-9, what="cov")


@pytest.mark.parametrize("w", [0, 1, 2])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_baum_welch_trajectory_as_written(HMM, rung0, w, k):
    feats = split_features(rung0)
    wf = [f for f, l in zip(feats, rung0["labels"]) if l == w]
    h = HMM(8, 13, feats, model_name=f"w{w}")
    hist = h.baum_welch(wf, k)
    ref = rung0[f"bw_w{w}_k{k}_hist"]
    assert len(hist) == len(ref)
    rt = {1: 1e-10, 2: 1e-8, 3: 1e-6}[k]     # the as-written emission is chaotic
    assert_close(np.array(hist), ref, rt, what="history")
    assert_close(h.A, rung0[f"bw_w{w}_k{k}_A"], rt, what="A")
    assert_close(h.B["mean"], rung0[f"bw_w{w}_k{k}_mean"], rt, what="mean")
    assert_close(h.B["covariance"], rung0[f"bw_w{w}_k{k}_cov"], rt * 10, what="cov")


def test_decode_as_written(HMM, rung0):
    feats = split_features(rung0)
    h = HMM(8, 13, feats)
    sc, path = h.decode(feats[0])
    assert path == rung0["flat_dec_path"].tolist()
    assert abs(sc - float(rung0["flat_dec_score"])) <= 1e-9 * abs(sc)
    assert len(path) == feats[0].shape[0] == 13                 # tests/test_decode.py:36-38
    for w in range(3):
        m = _model(HMM, rung0, f"bw_w{w}_k2_")
        for u in range(0, len(feats), 3):
            sc, path = m.decode(feats[u])
            assert path == rung0["dec_paths"][u, w].tolist(), (u, w)
            assert_close(sc, rung0["dec_scores"][u, w], 1e-7, what="score")
    with pytest.raises(IndexError):
        h.decode(feats[0][:, :10])           # T_frames < D
    with pytest.raises(ValueError):
        h.decode(feats[0].T)                 # (T, D) orientation


def test_standard_semantics_vs_rung1(HMM, rung1_d13):
    """semantics='standard': diagonal emission, decode walks every frame, fused kernels."""
    g = rung1_d13
    feats = split_features(g)
    S, D = 10, 13
    for prec, rs in (("fp64", 1e-12), ("fp32", 1e-6)):
        for w in (0, 3):
            h = HMM(8, 13, semantics="standard", precision=prec)
            h.A = g["A"][w].copy()
            cov = np.zeros((S, D, D))
            cov[:, np.arange(D), np.arange(D)] = g["var"][w]
            h.B = {"mean": g["means"][w].copy(), "covariance": cov}
            h.global_covariance = g["global_cov"]
            E = h.compute_emission_matrix(feats[0])
            if w == g["labels"][0]:
                assert_close(E, g["u0_own_E"], 1e-12, what="E")
            for u in (0, 5, 17):
                sc, path = h.decode(feats[u])
                T = feats[u].shape[1]
                assert path == g["dec_paths"][u, w, :T].tolist()
                assert_close(sc, g["dec_scores"][u, w], rs, what="score")
            wf = [f for f, l in zip(feats, g["labels"]) if l == w]
            hist = h.baum_welch(wf, 1)
            assert_close(np.array(hist), g["bw1_hist"][w:w + 1], rs, what="hist")
            assert_close(h.A, g["bw1_A"][w], 1e-10 if prec == "fp64" else 1e-5, atol=1e-10 if prec == "fp64" else 1e-5, what="A")
            sig = np.sqrt(g["var"][w])
            assert np.max(np.abs(h.B["mean"] - g["bw1_mean"][w]) / sig) < (1e-9 if prec == "fp64" else 1e-4)


def test_pickle_roundtrip(HMM, rung0, tmp_path):
    feats = split_features(rung0)
    h = HMM(8, 13, feats, model_name="heed")
    h.decode(feats[0])                       # creates device handles
    p = tmp_path / "heed_custom_15.pkl"      # train.py:108 naming
    with open(p, "wb") as f:
        pickle.dump(h, f)
    with open(p, "rb") as f:
        h2 = pickle.load(f)
    assert h2.decode(feats[0]) == h.decode(feats[0])


def test_reference_written_pickle_decodes(HMM, monkeypatch):
    """tests/golden/ref_heed_custom_2.pkl was written by the REFERENCE's custom_hmm.HMM through pickle.dump as in
    train.py:74-78 (tools/make_golden.py::ref_pickle); it resolves ``custom_hmm.HMM`` through the INTEGRATION.md shim,
    has no semantics / precision attributes, and must decode like the reference did (decoder.py:26-27, :42-47)."""
    import os
    import sys
    import types
    from conftest import GOLDEN, load_golden
    import sapr_b200.custom_hmm as shim_target
    shim = types.ModuleType("custom_hmm")
    shim.HMM = shim_target.HMM
    monkeypatch.setitem(sys.modules, "custom_hmm", shim)
    monkeypatch.delenv("SAPR_SEMANTICS", raising=False)
    with open(os.path.join(GOLDEN, "ref_heed_custom_2.pkl"), "rb") as f:
        h = pickle.load(f)
    assert isinstance(h, HMM) and h.semantics == "sapr" and h.model_name == "heed"
    g = load_golden("ref_pickle")
    feats = split_features(g)
    for u, x in enumerate(feats):
        sc, path = h.decode(x)
        assert path == g["dec_paths"][u].tolist()
        assert_close(sc, g["dec_scores"][u], 1e-9, what="score")
    hist = h.baum_welch(feats, 1)              # the unpickled object trains, too
    assert np.isfinite(hist[0])
    buf = pickle.dumps(h)                      # and re-pickles without device handles
    assert pickle.loads(buf).decode(feats[0]) == h.decode(feats[0])


# ------------------------------------------------------------------------------------------------
# the reference's own assertions (assignment2/tests/test_foward_backward.py, test_training.py,
# test_initialization.py, test_decode.py) against the drop-in class
@pytest.fixture(scope="module")
def feature_set():
    from sapr_b200 import synth
    feats, labels, _, _ = synth.make_corpus(33, 11, 8, 13, 60, 80, seed=5)
    return feats


@pytest.fixture
def hmm_model(HMM, feature_set):
    return HMM(8, 13, feature_set)


def test_ref_emission_matrix(hmm_model, feature_set):          # tests/test_foward_backward.py:17-37
    B_probs = hmm_model.compute_emission_matrix(feature_set[0])
    assert B_probs.shape == (feature_set[0].shape[1], hmm_model.total_states)
    assert np.all(B_probs <= 0)
    assert np.all(np.isfinite(B_probs[B_probs != -np.inf]))
    assert np.all(B_probs[:, 0] == -np.inf) and np.all(B_probs[:, -1] == -np.inf)
    assert np.std(B_probs[:, 1:-1][0, :]) < 1e-10


def test_ref_fb_probabilities(hmm_model, feature_set):         # tests/test_foward_backward.py:43-132
    E = hmm_model.compute_emission_matrix(feature_set[0])
    alpha, scale = hmm_model.forward(E)
    beta = hmm_model.backward(E, scale)
    T = E.shape[0]
    assert alpha.shape == (T, hmm_model.total_states) and beta.shape == (T, hmm_model.total_states)
    assert np.all(alpha[1:, 0] == -np.inf)
    assert np.all(beta[:-1, -1] == -np.inf) and beta[-1, -1] == 0
    assert np.all(alpha[alpha != -np.inf] <= 0) and np.all(beta[beta != -np.inf] <= 0)
    mid = T // 2
    post = alpha[mid] + beta[mid]
    post = post - np.logaddexp.reduce(post)
    gamma = hmm_model.compute_gamma(alpha, beta)
    gm = gamma[mid][1:-1]
    nz = gm > 0
    assert np.allclose(post[1:-1][nz], np.log(gm[nz]), atol=1e-5)


def test_ref_gamma_xi_update(hmm_model, feature_set):          # tests/test_training.py:111-275
    feats = feature_set[:3]
    gammas, agg_g, agg_x = [], 0, 0
    for f in feats:
        E = hmm_model.compute_emission_matrix(f)
        a, s = hmm_model.forward(E)
        b = hmm_model.backward(E, s)
        g = hmm_model.compute_gamma(a, b)
        x = hmm_model.compute_xi(a, b, E)
        assert g.shape == E.shape and x.shape == (E.shape[0] - 1, 10, 10)
        assert np.all((g >= 0) & (g <= 1)) and np.all((x >= 0) & (x <= 1))
        gammas.append(g); agg_g = agg_g + g[:-1].sum(0); agg_x = agg_x + x.sum(0)
    hmm_model.update_A(agg_x, agg_g)
    A = hmm_model.A
    assert A[0, 1] == 1.0 and A[-1, -1] == 1.0
    assert np.allclose(A[1:-1].sum(axis=1), 1.0, atol=1e-10)
    hmm_model.update_B(feats, gammas)
    mean, cov = hmm_model.B["mean"], hmm_model.B["covariance"]
    assert mean.shape == (10, 13) and cov.shape == (10, 13, 13)
    assert np.all(mean[0] == 0) and np.all(mean[-1] == 0) and np.all(cov[0] == 0) and np.all(cov[-1] == 0)
    floor = hmm_model.var_floor_factor * np.mean(np.diag(hmm_model.global_covariance))
    for j in range(1, 9):
        assert np.allclose(cov[j], cov[j].T)
        assert np.all(np.linalg.eigvalsh(cov[j]) > -1e-10 * np.abs(cov[j]).max())
        assert np.all(np.diag(cov[j]) >= floor)
    assert np.all(np.isfinite(mean))


def test_ref_decode_length(HMM, feature_set):                  # tests/test_decode.py:26-38
    h = HMM(8, 13, feature_set)
    heed = feature_set[0::11]
    h.baum_welch(heed, max_iter=2)
    log_prob, path = h.decode(heed[0])
    assert len(path) == heed[0].shape[0]
