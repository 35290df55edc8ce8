import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_err(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor) elementwise, NaNs must coincide."""
    a = np.asarray(a)
    b = np.asarray(b)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), "NaN pattern differs"
    ok = ~nan_b
    if not ok.any():
        return 0.0
    den = np.maximum(np.abs(b[ok]), floor if floor > 0 else np.finfo(float).tiny)
    return float(np.max(np.abs(a[ok] - b[ok]) / den))


def rt_err(a, b, zero_floor=1e-13):
    """Worst relative error of a per-channel probability array [..., n] against the golden one.

    Elementwise |a-b| / max(|b|, zero_floor * max_sigma |b[..., :]|): entries that are zero by
    symmetry (forbidden spin channels) carry round-off noise on both sides, so they are measured
    against the largest channel of the same point/source instead of against themselves.  Tunnelling
    values (1e-24) are still measured relative to themselves.
    """
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    assert a.shape == b.shape, (a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), "NaN pattern differs"
    scale = np.nanmax(np.abs(b), axis=-1, keepdims=True)
    den = np.maximum(np.abs(b), zero_floor * scale)
    den = np.where(den == 0, 1.0, den)
    err = np.abs(a - b) / den
    err = np.where(nan_b, 0.0, err)
    return float(err.max()) if err.size else 0.0
