"""Pins the CPU oracle (oracle/sapr_oracle.c) to vectors produced by the reference's own
custom_hmm.py (tools/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import assert_close, split_features
from oracle import oracle as orc

TOL = 1e-9  # float64 restatement vs numpy/BLAS summation order


def _model(g, pre):
    h = orc.OracleHMM(int(g["N"]), int(g["D"]))
    h.A = g[pre + "A"].copy()
    h.B = {"mean": g[pre + "mean"].copy(), "covariance": g[pre + "cov"].copy()}
    return h


def test_init_parameters(rung0):
    feats = split_features(rung0)
    h = orc.OracleHMM(8, 13, feats)
    assert_close(h.global_mean, rung0["init_global_mean"], 1e-12, what="global_mean")
    assert_close(h.global_covariance, rung0["init_global_cov"], 1e-12, what="global_cov")
    assert_close(h.A, rung0["init_A"], 1e-14, what="A")
    assert_close(h.B["mean"], rung0["init_mean"], 1e-12, what="mean")
    assert_close(h.B["covariance"], rung0["init_cov"], 1e-12, what="cov")
    # the reference sums float32 features in float32 (custom_hmm.py:76): 1e-6 agreement only
    assert_close(h.global_mean, rung0["init_f32_global_mean"], 2e-6, what="global_mean f32")


@pytest.mark.parametrize("pre,params,utt", [("flat_u0_", "init_", 0), ("trained_u1_", "bw_w0_k2_", 1)])
def test_per_function_sapr(rung0, pre, params, utt):
    feats = split_features(rung0)
    h = _model(rung0, params)
    E = h.compute_emission_matrix(feats[utt])
    assert_close(E, rung0[pre + "E"], TOL, what="E")
    Eg = rung0[pre + "E"]
    alpha, scale = h.forward(Eg)
    assert_close(alpha, rung0[pre + "alpha"], TOL, what="alpha")
    assert abs(scale - float(rung0[pre + "scale"])) <= TOL * max(1, abs(scale))
    beta = h.backward(Eg, float(rung0[pre + "scale"]))
    assert_close(beta, rung0[pre + "beta"], TOL, what="beta")
    gam = h.compute_gamma(rung0[pre + "alpha"], rung0[pre + "beta"])
    assert_close(gam, rung0[pre + "gamma"], 0, 1e-12, what="gamma")
    xi = h.compute_xi(rung0[pre + "alpha"], rung0[pre + "beta"], Eg)
    assert_close(xi, rung0[pre + "xi"], 0, 1e-12, what="xi")


def test_update_A_B(rung0):
    feats = split_features(rung0)
    h = _model(rung0, "bw_w0_k2_")
    h.global_covariance = rung0["init_global_cov"]
    h.update_A(rung0["upd_agg_xi"], rung0["upd_agg_gamma"])
    h.update_B([feats[1]], [rung0["trained_u1_gamma"]])
    assert_close(h.A, rung0["upd_A"], 1e-12, what="A")
    assert_close(h.B["mean"], rung0["upd_mean"], 1e-11, what="mean")
    assert_close(h.B["covariance"], rung0["upd_cov"], 1e-10, what="cov")


@pytest.mark.parametrize("w", [0, 1, 2])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_baum_welch_trajectory_sapr(rung0, w, k):
    feats = split_features(rung0)
    wf = [f for f, l in zip(feats, rung0["labels"]) if l == w]
    h = orc.OracleHMM(8, 13, feats)
    hist = h.baum_welch(wf, k)
    ref = rung0[f"bw_w{w}_k{k}_hist"]
    assert len(hist) == len(ref)
    # the as-written emission is chaotic: tolerance widens with the iteration
    rt = {1: 1e-10, 2: 1e-8, 3: 1e-6}[k]
    assert_close(np.array(hist), ref, rt, what="history")
    assert_close(h.A, rung0[f"bw_w{w}_k{k}_A"], rt, what="A")
    assert_close(h.B["mean"], rung0[f"bw_w{w}_k{k}_mean"], rt, what="mean")
    assert_close(h.B["covariance"], rung0[f"bw_w{w}_k{k}_cov"], rt * 10, what="cov")


def test_decode_sapr_first_D_frames(rung0):
    feats = split_features(rung0)
    h = orc.OracleHMM(8, 13, feats)
    sc, path = h.decode(feats[0])
    assert path == rung0["flat_dec_path"].tolist()          # exact ties: first candidate wins
    assert abs(sc - float(rung0["flat_dec_score"])) <= TOL * abs(sc)
    assert len(path) == 13                                   # SURVEY D3, tests/test_decode.py:36-38
    for w in range(3):
        m = _model(rung0, f"bw_w{w}_k2_")
        for u in range(len(feats)):
            sc, path = m.decode(feats[u])
            assert path == rung0["dec_paths"][u, w].tolist(), (u, w)
            assert_close(sc, rung0["dec_scores"][u, w], 1e-7, what="score")  # cond(cov) up to 1e7 after 2 as-written iterations


def test_decode_errors_like_reference(rung0):
    feats = split_features(rung0)
    h = orc.OracleHMM(8, 13, feats)
    with pytest.raises(IndexError):
        h.decode(feats[0][:, :10])       # T_frames < D


@pytest.mark.parametrize("name", ["rung1_d39", "rung1_d13", "rung1_mismatch_d39"])
def test_rung1_per_function_and_estep(name, request):
    g = request.getfixturevalue(name)
    feats = split_features(g)
    labels = g["labels"]
    A, means, var = g["A"], g["means"], g["var"]
    for pre, w in (("u0_own_", labels[0]), ("u0_other_", (labels[0] + 1) % int(g["M"]))):
        E = orc.emission_diag(feats[0], means[w], var[w])
        assert_close(E, g[pre + "E"], 1e-12, what="E")
        al, sc = orc.forward(E, A[w])
        assert_close(al, g[pre + "alpha"], 1e-12, what="alpha")
        be = orc.backward(E, A[w], sc)
        assert_close(be, g[pre + "beta"], 1e-12, what="beta")
        assert_close(orc.gamma(al, be), g[pre + "gamma"], 0, 1e-12, what="gamma")
        assert_close(orc.xi(al, be, E, A[w]), g[pre + "xi"], 0, 1e-12, what="xi")
    # batched E-step leg vs the per-utterance reference quantities
    X, offs = orc.pack(feats)
    stats, ll = orc.estep_batch(X, offs, labels, A, means, var)
    assert_close(ll, g["es_loglik"], 1e-12, what="loglik")
    M, S, D = means.shape
    for w in range(M):
        sel = labels == w
        emit = (np.arange(S) > 0) & (np.arange(S) < S - 1)   # only emitting states are consumed (custom_hmm.py:359-361)
        assert_close(stats[w, :S], g["es_G"][sel].sum(0) * emit, 0, 1e-10, what="G")
        assert_close(stats[w, S:2 * S], g["es_xi_self"][sel].sum(0) * emit, 0, 1e-10, what="xi_self")
        assert_close(stats[w, 2 * S:3 * S], g["es_occ"][sel].sum(0) * emit, 0, 1e-10, what="occ")


@pytest.mark.parametrize("name", ["rung1_d39", "rung1_d13", "rung1_mismatch_d39"])
def test_rung1_baum_welch_one_iteration(name, request):
    g = request.getfixturevalue(name)
    feats = split_features(g)
    labels = g["labels"]; M, S, D = g["means"].shape
    for w in range(M):
        wf = [f for f, l in zip(feats, labels) if l == w]
        cov = np.zeros((S, D, D))
        for j in range(S):
            cov[j] = np.diag(g["var"][w, j])
        hist, A, mean, cov2 = orc.baum_welch(wf, S - 2, g["A"][w], g["means"][w], cov, g["global_cov"], 1, 1)
        assert_close(np.array(hist), g["bw1_hist"][w:w + 1], 1e-12, what="hist")
        assert_close(A, g["bw1_A"][w], 1e-10, what="A")
        assert_close(mean, g["bw1_mean"][w], 1e-10, what="mean")
        assert_close(cov2, g["bw1_cov"][w], 1e-9, atol=1e-9, what="cov")


@pytest.mark.parametrize("name", ["rung1_d39", "rung1_d13"])
def test_rung1_decode_words_and_paths(name, request):
    g = request.getfixturevalue(name)
    feats = split_features(g)
    X, offs = orc.pack(feats)
    bw, bs, sc, bp = orc.viterbi_batch(X, offs, g["A"], g["means"], g["var"])
    assert np.array_equal(bw, g["dec_best"])
    assert_close(sc, g["dec_scores"], 1e-12, what="scores")
    for u in range(len(feats)):
        T = feats[u].shape[1]
        assert np.array_equal(bp[offs[u]:offs[u + 1]], g["dec_paths"][u, bw[u], :T]), u
        for w in range(int(g["M"])):          # every (utterance, model) path, not just the winner
            E = orc.emission_diag(feats[u], g["means"][w], g["var"][w])
            _, p = orc.decode(E, g["A"][w])
            assert np.array_equal(p, g["dec_paths"][u, w, :T]), (u, w)


@pytest.mark.parametrize("T", [2, 5, 8, 9, 10])
def test_edge_cases_short_utterances(edge, T):
    """T < N+1 makes gamma rows NaN in the reference and leaves the exit state unreachable in
    decode (path = zeros + [S-1]); the oracle must reproduce the exact inf/NaN pattern."""
    A, means, var = edge["A"][0], edge["means"][0], edge["var"][0]
    x = edge[f"T{T}_x"].astype(np.float64).T     # (D, T)
    pre = f"T{T}_"
    with np.errstate(all="ignore"):
        E = orc.emission_diag(x, means, var)
        assert_close(E, edge[pre + "E"], 1e-12, what="E")
        al, sc = orc.forward(E, A)
        assert_close(al, edge[pre + "alpha"], 1e-12, what="alpha")
        be = orc.backward(E, A, sc)
        assert_close(be, edge[pre + "beta"], 1e-12, what="beta")
        assert_close(orc.gamma(al, be), edge[pre + "gamma"], 0, 1e-12, what="gamma")
        assert_close(orc.xi(al, be, E, A), edge[pre + "xi"], 0, 1e-12, what="xi")
        s, p = orc.decode(E, A)
    assert np.array_equal(p, edge[pre + "dec_path"])
    assert_close(s, edge[pre + "dec_score"], 1e-12, what="score")
