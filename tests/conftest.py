import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def split_features(g):
    """golden X (sum T, D) f32 + offsets -> list of (D, T_u) float64 arrays (reference orientation)."""
    X, offs = g["X"], g["offsets"]
    return [np.ascontiguousarray(X[offs[u]:offs[u + 1]].T.astype(np.float64)) for u in range(len(offs) - 1)]


@pytest.fixture(scope="session")
def rung0():
    return load_golden("rung0_d13")


@pytest.fixture(scope="session")
def rung1_d39():
    return load_golden("rung1_d39")


@pytest.fixture(scope="session")
def rung1_d13():
    return load_golden("rung1_d13")


@pytest.fixture(scope="session")
def rung1_mismatch_d39():
    """Models moved off the generating parameters (an early Baum-Welch iteration): contains utterances whose xi underflows in
    the reference's float64 arithmetic and therefore drop out of the transition statistics."""
    return load_golden("rung1_mismatch_d39")


@pytest.fixture(scope="session")
def edge():
    return load_golden("edge_cases")


def assert_close(a, b, rtol, atol=0.0, what=""):
    """Finite cells within tolerance; the +-inf / NaN pattern must match exactly."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    fa, fb = np.isfinite(a), np.isfinite(b)
    assert np.array_equal(fa, fb), f"{what}: finite pattern differs"
    assert np.array_equal(np.isnan(a), np.isnan(b)), f"{what}: NaN pattern differs"
    assert np.array_equal(a[~fa & ~np.isnan(a)], b[~fb & ~np.isnan(b)]), f"{what}: inf sign differs"
    if fa.any():
        err = np.abs(a[fa] - b[fa]); tol = atol + rtol * np.maximum(1.0, np.abs(b[fb]))
        worst = float(np.max(err - tol))
        assert worst <= 0, f"{what}: max abs err {float(err.max()):.3e} exceeds tolerance by {worst:.3e}"
