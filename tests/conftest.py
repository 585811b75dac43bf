import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def rslib():
    """The CUDA library through its C ABI; GPU tests fail loudly if it is missing."""
    from roadsurf_b200 import lib
    lib.load()
    if lib.load().roadsurf_device_count() < 1:
        pytest.fail("no CUDA device visible to libroadsurf_b200.so")
    return lib
