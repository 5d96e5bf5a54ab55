import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

# The unchanged reference wrapper calls np.float(dt) (parament.py:271), removed in numpy >= 1.24.
if not hasattr(np, "float"):
    np.float = float


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count():
    import ctypes
    try:
        cuda = ctypes.CDLL("libcuda.so.1")
    except OSError:
        return 0
    n = ctypes.c_int(0)
    if cuda.cuInit(0) != 0 or cuda.cuDeviceGetCount(ctypes.byref(n)) != 0:
        return 0
    return n.value


@pytest.fixture(scope="session")
def gpu_count():
    return _cuda_device_count()


def pytest_collection_modifyitems(config, items):
    # gpu tests are selected with -m gpu by the driver; when somebody runs the whole suite on a CPU box,
    # skip them with a clear reason instead of failing inside Parament_create (code 30).
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
