"""train.py / eval.py orchestration: whole-vocabulary training from a feature_set directory of
(13, T) float32 .npy files, model pickles named <word>_<impl>_<n_iter>.pkl, evaluation dictionaries, and the batched
recognition path against the per-sequence loop of decoder.py:42-47."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WORDS = ["heed", "hid", "head", "had", "hard", "hud", "hod", "hoard", "hood", "whod", "heard"]


@pytest.fixture(scope="module")
def corpus_dir(tmp_path_factory):
    from sapr_b200 import synth
    root = tmp_path_factory.mktemp("corpus")
    fs = root / "feature_set"; fs.mkdir()
    feats, labels, mu, sd = synth.make_corpus(11 * 8, 11, 8, 13, 40, 60, seed=4242)
    for u, (x, w) in enumerate(zip(feats, labels)):
        np.save(fs / f"sp{u // 11:02d}a_w{u:03d}_{WORDS[w]}.npy", x)       # mfcc_extract.py:41-42: (13, T) float32; word = last "_" field
    return root, feats, labels


def test_train_eval_custom_standard(corpus_dir, monkeypatch):
    from sapr_b200.train import train_hmm
    from sapr_b200.eval import eval_hmm
    from sapr_b200.decoder import Decoder
    root, feats, labels = corpus_dir
    hmms = train_hmm("custom", 8, 13, n_iter=4, feature_set_path=str(root / "feature_set"),
                     models_dir=str(root / "trained_models"), semantics="standard")
    assert list(hmms) == WORDS
    for w in WORDS:
        assert (root / "trained_models" / "custom" / f"{w}_custom_4.pkl").exists()      # train.py:108
        h = train_hmm.histories[w]
        assert 1 <= len(h) <= 4 and np.all(np.isfinite(h))
        assert all(b >= a - 1e-6 * abs(a) for a, b in zip(h, h[1:]))                     # EM with a true diagonal Gaussian is monotone
        assert np.allclose(hmms[w].A.sum(axis=1), 1.0)
    monkeypatch.setenv("SAPR_DECODE_LOOP", "1")        # the reference's utterances x models loop (decoder.py:42-47) ...
    seq = eval_hmm("custom", str(root / "feature_set"), model_iter=4, models_dir=str(root / "trained_models"), vocab_order=WORDS)
    monkeypatch.delenv("SAPR_DECODE_LOOP")             # ... against the default route (one fused launch per word) and the whole-set launch
    dflt = eval_hmm("custom", str(root / "feature_set"), model_iter=4, models_dir=str(root / "trained_models"), vocab_order=WORDS)
    assert dflt["predicted_labels"] == seq["predicted_labels"] and dflt["confusion_matrix"].equals(seq["confusion_matrix"])
    bat = eval_hmm("custom", str(root / "feature_set"), model_iter=4, models_dir=str(root / "trained_models"), vocab_order=WORDS,
                   batched=True)
    # no accuracy bar here: the reference's exit state is free once t >= N, so a custom-model score only
    # covers a short prefix of the utterance and recognition is weak by construction -- parity, not accuracy, is the gate
    assert seq["accuracy"] == pytest.approx(np.mean([a == b for a, b in zip(seq["true_labels"], seq["predicted_labels"])]))
    assert seq["true_labels"] == bat["true_labels"]
    assert seq["predicted_labels"] == bat["predicted_labels"]                             # one fused launch == utterances x models loop
    assert seq["confusion_matrix"].equals(bat["confusion_matrix"])
    assert int(seq["confusion_matrix"].values.sum()) == len(feats)
    for w in WORDS:
        for a, b in zip(seq["results"][w], bat["results"][w]):
            assert a["state_sequence"] == b["state_sequence"]
            assert abs(a["log_likelihood"] - b["log_likelihood"]) <= 1e-6 * abs(a["log_likelihood"])
    # ... and both against the CPU oracle (custom_hmm.py:462-514 x all models + decoder.py:42-47) on the trained parameters, and
    # the metrics against sklearn exactly as eval.py:28-38 calls it
    from oracle import oracle as orc
    from sklearn.metrics import accuracy_score, confusion_matrix
    from sapr_b200.mfcc_extract import load_mfccs_by_word
    ev_feats = [f for w in WORDS for f in load_mfccs_by_word(str(root / "feature_set"), w)]    # eval walks the vocabulary word by word, files in listdir order
    X, offs = orc.pack(ev_feats)
    Am = np.stack([hmms[w].A for w in WORDS]); mm = np.stack([hmms[w].B["mean"] for w in WORDS])
    vm = np.stack([np.diagonal(hmms[w].B["covariance"], axis1=1, axis2=2) for w in WORDS])
    bw, bs, sc, bp = orc.viterbi_batch(X, offs, Am, mm, vm)
    pred_idx = np.array([WORDS.index(w) for w in bat["predicted_labels"]])
    sc_sorted = np.sort(sc, axis=1)
    clear = (sc_sorted[:, -1] - sc_sorted[:, -2]) > 1e-6 * np.abs(sc_sorted[:, -1])            # fp32 production mode: exact-score near-ties aside
    assert np.array_equal(pred_idx[clear], bw[clear]) and clear.mean() > 0.95
    flat_paths = [p for w in WORDS for r in bat["results"][w] for p in r["state_sequence"]]
    agree = np.mean(np.array(flat_paths)[np.repeat(clear, np.diff(offs))] == bp[np.repeat(clear, np.diff(offs))])
    assert agree > 0.999, agree
    t_idx = [WORDS.index(w) for w in bat["true_labels"]]
    assert np.array_equal(bat["confusion_matrix"].values, confusion_matrix(t_idx, pred_idx))
    assert bat["accuracy"] == pytest.approx(accuracy_score(t_idx, pred_idx))
    # the reference's per-sequence entry point still works on the pickles
    d = Decoder(models_dir=str(root / "trained_models"), implementation="custom", n_iter=4, vocab_order=WORDS)
    word, score, states = d.decode_sequence(feats[3].T)
    assert word in WORDS and len(states) == feats[3].shape[1] and np.isfinite(score)


def test_train_eval_hmmlearn_style(corpus_dir, monkeypatch):
    from sapr_b200.train import train_hmm
    from sapr_b200.eval import eval_hmm
    root, feats, labels = corpus_dir
    monkeypatch.chdir(root)                      # HMMLearnModel re-loads ./feature_set for its global statistics (hmmlearn_hmm.py:23)
    hmms = train_hmm("hmmlearn", 8, 13, n_iter=3, feature_set_path=str(root / "feature_set"),
                     models_dir=str(root / "trained_models"))
    assert all(len(train_hmm.histories[w]) >= 1 for w in WORDS)
    res = eval_hmm("hmmlearn", str(root / "feature_set"), model_iter=3, models_dir=str(root / "trained_models"), vocab_order=WORDS)
    assert res["accuracy"] >= 0.9
    assert set(res) == {"results", "accuracy", "confusion_matrix", "true_labels", "predicted_labels"}   # eval.py:130-136


def test_confusion_on_device_vs_sklearn():
    """sapr_confusion (eval.py:28-38 on the device) against sklearn.metrics, incl. unreachable predictions (-1) and a word
    that never occurs."""
    import torch
    from sklearn.metrics import confusion_matrix
    from sapr_b200.engine import confusion_on_device
    rng = np.random.default_rng(3)
    M, B = 11, 100_003
    t = rng.integers(0, M - 1, B).astype(np.int32)           # word M-1 never occurs as truth
    p = np.where(rng.random(B) < 0.8, t, rng.integers(-1, M, B)).astype(np.int32)
    cm, acc = confusion_on_device(torch.as_tensor(t).cuda(), torch.as_tensor(p).cuda(), M)
    cm = cm.cpu().numpy()
    ok = p >= 0
    ref = confusion_matrix(t[ok], p[ok], labels=list(range(M)))
    assert np.array_equal(cm[:, :M], ref)
    assert np.array_equal(cm[:, M], np.bincount(t[~ok], minlength=M))
    assert acc == pytest.approx(float(np.mean(t == p)))
