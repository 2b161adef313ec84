import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden_matrix(name):
    """Decoded copy of the reference's data/<name>.RData (tools/decode_rdata.py)."""
    return np.loadtxt(os.path.join(ROOT, "tests", "golden", name + ".txt"), skiprows=1, dtype=np.int32)


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def datasets():
    return {n: load_golden_matrix(n) for n in ("K2_N100_P5", "K2_N1000_P5", "K3_N1000_P5")}


def gpu_available():
    try:
        from bmm_mcmc_b200 import _lib
        return _lib.lib().bmm_device_count() > 0
    except Exception:
        return False
