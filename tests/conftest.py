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


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def assert_close(actual, desired, rtol=1e-5, atol_scale=1e-6, msg=""):
    """allclose(rtol, atol = atol_scale * max|desired|) -- BASELINE.md section 3.6."""
    actual = np.asarray(actual, dtype=np.float64)
    desired = np.asarray(desired, dtype=np.float64)
    atol = atol_scale * max(1e-30, float(np.max(np.abs(desired)))) if desired.size else 0.0
    np.testing.assert_allclose(actual, desired, rtol=rtol, atol=atol, err_msg=msg)
