"""CPU-only tests: the C-ABI library loads and exports every symbol include/sapr_b200.h declares, the
product fails loudly without a GPU, sharding and the world_size-2 statistics all-reduce (gloo)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_so():
    return os.path.exists(os.path.join(ROOT, "sapr_b200", "libsaprb200.so"))


def test_library_exports_every_declared_symbol():
    if not _have_so():
        import __graft_entry__ as ge
        ge.build()
    hdr = open(os.path.join(ROOT, "include", "sapr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sapr_[A-Za-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = ctypes.CDLL(os.path.join(ROOT, "sapr_b200", "libsaprb200.so"))
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    from sapr_b200 import _lib
    assert set(_lib.SIGNATURES) == declared, set(_lib.SIGNATURES) ^ declared
    assert _lib.load().sapr_version() >= 100
    assert _lib.load().sapr_stats_stride(8, 39) == 3 * 10 + 2 * 10 * 39


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sapr_b200 import _lib
    from sapr_b200.custom_hmm import HMM
    with pytest.raises(RuntimeError):
        _lib.Context()
    with pytest.raises(RuntimeError):
        HMM(8, 13, [np.zeros((13, 20), dtype=np.float32)])
    h = ctypes.c_void_p()
    assert _lib.load().sapr_ctx_create(0, None, ctypes.byref(h)) != 0


def test_reference_written_pickle_loads(monkeypatch):
    """A pickle written by the reference's train.py (tests/golden/ref_heed_custom_2.pkl, made by
    tools/make_golden.py::ref_pickle from the reference's own class) unpickles through the custom_hmm shim into a
    usable drop-in object: attributes of custom_hmm.py:10-33 kept, mode attributes defaulted."""
    import pickle
    import types
    import sapr_b200.custom_hmm as target
    shim = types.ModuleType("custom_hmm")
    shim.HMM = target.HMM
    monkeypatch.setitem(sys.modules, "custom_hmm", shim)
    monkeypatch.delenv("SAPR_SEMANTICS", raising=False)
    monkeypatch.delenv("SAPR_FP64_VERIFY", raising=False)
    with open(os.path.join(ROOT, "tests", "golden", "ref_heed_custom_2.pkl"), "rb") as f:
        h = pickle.load(f)
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_pickle.npz"))
    assert isinstance(h, target.HMM)
    assert (h.semantics, h.precision, h._dev, h._dev_key) == ("sapr", "fp32", None, None)
    assert (h.num_states, h.num_obs, h.total_states, h.model_name) == (8, 13, 10, "heed")
    assert np.array_equal(h.A, g["A"]) and np.array_equal(h.B["mean"], g["mean"])
    assert h.B["covariance"].shape == (10, 13, 13) and h.pi[0] == 1.0
    assert h._prec() == target.FP32
    # a bare reference-shaped state dict (no pickle) takes the same route
    st = {k: v for k, v in h.__dict__.items() if k not in ("semantics", "precision", "_dev", "_dev_key")}
    h2 = target.HMM.__new__(target.HMM)
    monkeypatch.setenv("SAPR_SEMANTICS", "standard")
    h2.__setstate__(st)
    assert h2.semantics == "standard" and h2._dev is None
    assert set(pickle.loads(pickle.dumps(h2)).__dict__) >= set(st)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sapr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_shard_bounds_balance_by_frames():
    from sapr_b200.dist import shard_bounds
    rng = np.random.default_rng(0)
    lens = rng.integers(10, 400, size=1000)
    offs = np.concatenate([[0], np.cumsum(lens)])
    for world in (1, 2, 4, 8):
        b = shard_bounds(offs, world)
        assert b[0][0] == 0 and b[-1][1] == 1000
        assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        frames = [offs[e] - offs[s] for s, e in b]
        assert max(frames) - min(frames) <= 2 * lens.max()
    assert shard_bounds(np.array([0, 5]), 4) == [(0, 0), (0, 0), (0, 0), (0, 1)] or sum(e - s for s, e in shard_bounds(np.array([0, 5]), 4)) == 1


def test_synth_is_deterministic_and_left_to_right():
    from sapr_b200 import synth
    f1, l1, mu, sd = synth.make_corpus(22, 11, 8, 39, 40, 60, seed=1)
    f2, l2, _, _ = synth.make_corpus(22, 11, 8, 39, 40, 60, seed=1)
    assert all(np.array_equal(a, b) for a, b in zip(f1, f2)) and np.array_equal(l1, l2)
    assert f1[0].dtype == np.float32 and f1[0].shape[0] == 39
    X, offs = synth.pack_frame_major(f1)
    assert X.shape[1] == 40 and X.dtype == np.float32 and np.all(X[:, 39] == 0)
    assert np.array_equal(X[offs[3]:offs[4], :39], f1[3].T)
    A, means, var = synth.truth_models(mu, sd)
    assert np.allclose(A.sum(axis=2)[:, :-1], 1.0) and A.shape == (11, 10, 10)


_WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["SAPR_ROOT"])
import torch
from sapr_b200.dist import Dist, shard_bounds
from sapr_b200 import synth
from oracle import oracle as orc
d = Dist(backend="gloo")
feats, labels, mu, sd = synth.make_corpus(44, 11, 8, 13, 30, 50, seed=3)
A, means, var = synth.truth_models(mu, sd, 0.85)
X, offs = orc.pack(feats)
s, e = shard_bounds(offs, d.world)[d.rank]
Xs, offs_s = orc.pack(feats[s:e])
stats, ll = orc.estep_batch(Xs, offs_s, labels[s:e], A, means, var)     # oracle stands in for the GPU E-step
ll_m = np.zeros(11); np.add.at(ll_m, labels[s:e], ll)
packed = torch.from_numpy(np.concatenate([stats.reshape(-1), ll_m]))
d.allreduce_(packed)
full, ll_full = orc.estep_batch(X, offs, labels, A, means, var)
ll_fm = np.zeros(11); np.add.at(ll_fm, labels, ll_full)
ref = np.concatenate([full.reshape(-1), ll_fm])
err = float(np.max(np.abs(packed.numpy() - ref) / np.maximum(1.0, np.abs(ref))))
assert err < 1e-12, err
mx = torch.tensor([float(d.rank)]); d.max_(mx); assert mx.item() == d.world - 1
d.barrier(); d.shutdown()
print("RANK_OK", d.rank, err)
'''


def test_world_size_2_gloo_stats_allreduce(tmp_path):
    """The N>1 path on CPU: shard by frames, E-step per shard, ONE all-reduce of the packed buffer,
    result equals the unsharded statistics."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, SAPR_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", WORLD_SIZE="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "RANK_OK" in o, o


def test_mfcc_oracle_basic_properties():
    from oracle import mfcc_oracle
    sr = 22050
    t = np.arange(sr) / sr
    y = np.sin(2 * np.pi * 440 * t)          # the reference's own test signal (tests/test_mfcc_extract.py:10-28)
    m = mfcc_oracle.mfcc(y, sr, 2048, 661, 220, 128, 13)
    assert m.shape == (13, 1 + sr // 220) and np.all(np.isfinite(m))
    W = mfcc_oracle.mel_filterbank(sr, 2048, 128, 0.0, sr / 2, True)
    assert W.shape == (128, 1025) and np.all(W >= 0) and np.all(W.sum(axis=1) > 0)


def test_eval_metrics_match_sklearn():
    """eval.py:28-38 calls sklearn's confusion_matrix / accuracy_score without labels=: rows are the labels that occur."""
    sk = pytest.importorskip("sklearn.metrics")
    from sapr_b200.eval import calculate_metrics, extract_labels
    vocab = ["heed", "hid", "head", "had", "hard", "hud", "hod", "hoard", "hood", "whod", "heard"]
    rng = np.random.default_rng(3)
    for present in (vocab, vocab[:4] + vocab[6:7]):                       # a case where some words never occur
        t = [present[i] for i in rng.integers(0, len(present), 200)]
        p = [present[i] for i in rng.integers(0, len(present), 200)]
        cm, acc = calculate_metrics(t, p, vocab)
        ti = [vocab.index(w) for w in t]; pi = [vocab.index(w) for w in p]
        assert np.array_equal(cm, sk.confusion_matrix(ti, pi))
        assert acc == pytest.approx(sk.accuracy_score(ti, pi))
    res = {"heed": [{"true_word": "heed", "predicted_word": "hid"}], "hid": [{"true_word": "hid", "predicted_word": "hid"}]}
    assert extract_labels(res) == (["heed", "hid"], ["hid", "hid"])


def test_committed_bench_lines_follow_the_contract():
    """The bench lines kept under profiles/ (written by bench.py on a B200) carry every key of the measurement contract."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    line = json.loads(open(os.path.join(root, "profiles", "r01_bench_s4.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in line, k
    assert line["n_gpus"] == 1 and line["warmup"] >= 3 and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert "workload" in line["config"] and "model" not in line["config"]
    r = line["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] is not None
    c = line["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = line["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < line["value"]
    assert line["gpu_launches"] > 0
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    ref = json.loads(open(os.path.join(root, "profiles", "r01_bench_ref_s4.json")).read().strip().splitlines()[-1])
    assert ref["impl"] == "reference" and ref["metric"] == line["metric"] and ref["unit"] == line["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["value"] == ref["value"]


def test_cpulist_parser_and_numa_bind_is_harmless_without_topology():
    from sapr_b200.dist import _parse_cpulist, bind_to_gpu_numa_node
    assert _parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert _parse_cpulist("5") == [5]
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node(0) is None          # no GPU / no sysfs entry here: nothing changes
    assert os.sched_getaffinity(0) == before
