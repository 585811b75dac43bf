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


@pytest.fixture(scope="session")
def exact(rslib):
    """Bit-identity with the oracle needs the host libm to be the one roadsurf_b200/csrc/rs_libm.h mirrors
    (glibc's exp / log, FMA code path).  On a box where it is not, the exact tests FAIL with this message
    -- they do not skip: a silent skip would hide a regression to the 0.5-1.3 % flip regime."""
    bad = rslib.selftest_libm(2_000_000)
    if bad != [0, 0]:
        pytest.fail(f"host libm differs from the one rs_libm.h mirrors (exp/log mismatches {bad} of 2e6): "
                    "the CUDA path cannot be bit-identical to the oracle on this machine; regenerate "
                    "rs_libm_tables.h with scripts/gen_libm_tables.py for this glibc")
    return True
