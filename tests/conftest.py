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


@pytest.fixture
def knob(codec, monkeypatch):
    """Set a developer knob (HZ_* environment variable) for one test: the library reads them at context creation,
    so the session's context is told to re-read them, and again when the test is over."""
    def set_knob(name, value):
        monkeypatch.setenv(name, value)
        codec.reload_knobs()
    yield set_knob
    monkeypatch.undo()
    codec.reload_knobs()


@pytest.fixture(params=["fused", "legacy"])
def decmode(request, knob):
    """Both decoders on the same cases: the fused single-walk kernel and the multi-pass kernels (the library picks
    by chunk size; HZ_DEC forces one)."""
    knob("HZ_DEC", request.param)
    return request.param
