"""The C++ restatement (oracle/) against the UNMODIFIED reference built by oracle/build_ref.sh.

No Fortran compiler exists in this image, so oracle/_ref cannot be built and these tests SKIP --
loudly, with the reason printed in the pytest summary (`-rs`); they never pass silently.  The moment
`gfortran` (or another F2008 compiler, FC=...) is on PATH, or a prebuilt oracle/_ref/libroadsurf_ref*.so
is dropped into place, they run and pin the oracle: the strict build (-O2, no fast-math) must be
reproduced to 1e-9 (same IEEE operations, libm aside), the reference's own -Ofast build within the
north-star tolerance (1e-3 K / 1e-3 mm) except for threshold flips, which are counted.
"""
import numpy as np
import pytest

from golden_io import CASE_NAMES, load_case
from parity import S_TOL, T_TOL, compare


def _ref_or_skip(oracle, strict):
    try:
        return oracle.load_ref(strict=strict)
    except oracle.ReferenceUnavailable as e:
        pytest.skip(f"oracle/_ref not available -- PARITY STAYS UNPINNED: {e}")


@pytest.mark.parametrize("name", CASE_NAMES)
def test_restatement_equals_strict_reference_build(oracle, name):
    _ref_or_skip(oracle, strict=True)
    arrays, settings, params, golden, status, stride, steps = load_case(name)
    ref = arrays.copy()
    oracle.run_batch(arrays, settings, params, nthreads=2)
    oracle.run_batch(ref, settings, params, nthreads=2, backend="ref_strict")
    for k in arrays.out:
        assert np.max(np.abs(arrays.out[k] - ref.out[k])) < 1e-9, k
    # the reference mutates its inputs in place (VZ(1), SW_dir, sky-view SW/LW): so must the restatement
    for k in ("VZ", "SW", "SW_dir", "LW"):
        assert np.array_equal(getattr(arrays, k), getattr(ref, k)), k


@pytest.mark.parametrize("name", CASE_NAMES)
def test_restatement_within_tolerance_of_reference_flag_build(oracle, name):
    _ref_or_skip(oracle, strict=False)
    arrays, settings, params, golden, status, stride, steps = load_case(name)
    ref = arrays.copy()
    oracle.run_batch(arrays, settings, params, nthreads=2)
    oracle.run_batch(ref, settings, params, nthreads=2, backend="ref")
    r = compare(arrays.out, ref.out)
    print(f"{name}: oracle vs gfortran -Ofast reference: {r}")
    assert r["max_dT_matching"] <= T_TOL and r["max_dS_matching"] <= S_TOL, r
    assert r["mismatch_fraction"] <= 0.35, r   # threshold flips between two differently rounded builds


def test_build_recipe_reports_why_it_cannot_build(oracle):
    """The recipe itself is always exercised: it either builds or exits 3 / 4 with a reason."""
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(oracle.__file__))
    r = subprocess.run(["sh", os.path.join(here, "build_ref.sh")], capture_output=True, text=True)
    assert r.returncode in (0, 3, 4), r.stderr
    if r.returncode != 0:
        assert "build_ref:" in r.stderr
