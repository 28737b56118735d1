import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _device_count():
    from ndsm_b200 import load_library
    return load_library().ndsm_b200_device_count()


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library; GPU tests are skipped (not silently passed) when no device is visible."""
    from ndsm_b200 import load_library
    lib = load_library()  # raises if the .so is missing: no fallback
    if lib.ndsm_b200_device_count() <= 0:
        pytest.skip("no CUDA device visible")
    return lib


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


def rel_err(a, b):
    """max|a-b| / max|b| (0 when both are identically zero)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.abs(b).max()
    d = np.abs(a - b).max()
    return 0.0 if d == 0.0 else d / (scale if scale > 0 else 1.0)


def aniso_mesh(shape, stretch=(1.0, 1.3, 0.8), origin=(0.0, 0.25, -0.5)):
    """Uniform mesh vectors with different spacing per dimension; shape = (nx, ny[, nz])."""
    h = 1.0 / (shape[0] - 1)
    return [origin[d] + np.arange(n) * h * stretch[d] for d, n in enumerate(shape)]
