import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU checker (oracle/): test infrastructure only."""
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "reference_cases.npz")
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def qpde():
    """The product library through its C ABI (ctypes)."""
    from quantpde_b200 import capi
    capi.lib()
    return capi


@pytest.fixture(scope="session")
def handle(qpde):
    try:
        h = qpde.Handle(0)   # raises without a B200: there is no CPU fallback
    except qpde.QpdeError as e:
        if e.code == qpde.ERR_NO_DEVICE:
            pytest.skip("no sm_100 device: %s" % e)
        raise
    yield h
    h.close()
