import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def hz():
    import __graft_entry__ as ge
    return ge.load_package()


@pytest.fixture(scope="session")
def codec(hz):
    """One codec context on cuda:0.  Fails loudly (no CPU fallback) when the library or GPU is missing."""
    c = hz.Codec(0)
    yield c
    c.close()
