"""pytest configuration: the ``gpu`` marker and shared helpers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def relerr(a, b):
    """max-norm error of ``a`` against reference ``b`` relative to max|b| (>= tiny)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin), "finite/non-finite pattern differs"
    if not fin.any():
        return 0.0
    scale = max(float(np.max(np.abs(b[fin]))), 1e-300)
    return float(np.max(np.abs(a[fin] - b[fin]))) / scale


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load
