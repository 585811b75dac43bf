"""The oracle against its committed golden fixtures (regression pins; see make_golden.py)."""
import numpy as np
import pytest

from golden_io import CASE_NAMES, load_case


@pytest.mark.parametrize("name", CASE_NAMES)
def test_oracle_reproduces_golden_outputs(oracle, name):
    arrays, settings, params, golden, status, stride, steps = load_case(name)
    st, executed = oracle.run_batch(arrays, settings, params, nthreads=2)
    assert np.array_equal(st, status)
    assert executed == steps
    for k, want in golden.items():
        got = arrays.out[k][:, ::stride]
        assert got.shape == want.shape
        # same source, same flags: differences can only come from the host libm
        assert np.max(np.abs(got - want)) < 1e-9, k


def test_golden_covers_coupling_and_sky_view():
    arrays, settings, params, golden, status, stride, steps = load_case("coupled")
    assert settings.use_coupling == 1 and settings.use_relaxation == 1
    assert steps > arrays.npoints * arrays.sim_len  # coupling re-runs were executed
    assert any(arrays.local[p].sky_view < 1.0 for p in range(arrays.npoints))
    assert all(arrays.local[p].couplingIndexI == 720 for p in range(arrays.npoints))
    assert (status & 8).all()
