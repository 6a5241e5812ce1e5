import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG = "a-simple-stereo-slam-system-with-deep-loop-closing_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def pkg():
    return importlib.import_module(PKG)


@pytest.fixture(scope="session")
def synth():
    return importlib.import_module(PKG + ".synth")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ref():
    """oracle/_ref: the reference's own src/ORBextractor.cpp, compiled unmodified (oracle/Makefile `ref`)."""
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref is not built and /root/reference is absent")
    R.lib()
    return R
