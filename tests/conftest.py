import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O

    O.build()
    O.lib()
    return O


@pytest.fixture(scope="session")
def pyref():
    from oracle import pyref as P

    return P


@pytest.fixture(scope="session")
def engine():
    """The product package (ctypes over libmsm_b200.so)."""
    import ec_gpu_b200 as m

    m.load_library()
    return m
