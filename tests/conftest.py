import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def ksn():
    """The product library through its ctypes binding (C-ABI)."""
    from kspace_neutrinos_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        capi.build()
    return capi.lib()


@pytest.fixture(scope="session")
def gpu(ksn):
    if not ksn.ksn_device_available():
        pytest.fail("test marked gpu but no CUDA device is visible")
    from kspace_neutrinos_b200 import capi
    capi.check(ksn.ksn_init(-1), "ksn_init")
    ksn.ksn_set_quiet(1)
    return ksn
