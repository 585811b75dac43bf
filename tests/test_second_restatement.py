"""N-version check: the C++ oracle against a second restatement written independently from the Fortran
(oracle/py_restatement.py, plain Python, one point at a time) on the golden cases -- plain forecast with
sky-view points, and analysis + forecast with coupling and relaxation.  Both are IEEE double without FMA on
the same libm, so they must agree to rounding noise (1e-9 is the bar; in practice they agree exactly)."""
import numpy as np
import pytest

from golden_io import CASE_NAMES, load_case
from oracle import py_restatement


@pytest.mark.parametrize("name", CASE_NAMES)
def test_two_restatements_agree(oracle, name):
    arrays, settings, params, golden, status, stride, steps = load_case(name)
    ref = arrays.copy()
    st, executed = oracle.run_batch(ref, settings, params, nthreads=2)
    total, worst = 0, 0.0
    for p in range(arrays.npoints):
        out, nsteps, failed = py_restatement.run_point(arrays, settings, params, p)
        total += nsteps
        assert failed == bool(st[p] & 1)
        for k, v in out.items():
            d = np.abs(v - ref.out[k][p])
            assert np.array_equal(np.isnan(v), np.isnan(ref.out[k][p])), (p, k)
            worst = max(worst, float(np.nanmax(d)))
            assert np.nanmax(d) < 1e-9, (name, p, k, int(np.nanargmax(d)), float(np.nanmax(d)))
    assert total == executed          # same number of executed steps: the coupling re-runs agree too
    print(f"{name}: max |difference| between the two restatements = {worst:.3e}")


def test_second_restatement_layer_depths_and_sun():
    """The pieces with REAL(4) arithmetic inside: layer depths (src/Initialization.f90:217-235) and the
    Julian-day fraction (src/SunPosition.f90:245-253)."""
    arrays, settings, params, *_ = load_case("plain")
    pt = py_restatement.Point({k: getattr(arrays, v)[0] for k, v in
                               dict(Tair="tair", Tdew="tdew", VZ="VZ", Rhz="Rhz", prec="prec", SW="SW", LW="LW", SW_dir="SW_dir",
                                    LW_net="LW_net", TSurfObs="TSurfObs", PrecPhase="PrecPhase", Depth="Depth").items()}
                              | {n: arrays.time[k] for k, n in enumerate(("year", "month", "day", "hour", "minute", "second"))}
                              | {"local_horizons": arrays.local_horizons[0]}, settings, params, arrays.local[0])
    pt.initialization()
    # layer depths for N = 15 (m), four decimals
    want = [0, 0.0303, 0.0647, 0.1049, 0.1532, 0.2127, 0.2881, 0.3857, 0.5143, 0.6863, 0.9191, 1.2370, 1.6741, 2.2781,
            3.1156, 4.2801]
    assert np.allclose(pt.Z[1:17], want, atol=5e-5)
    from oracle import pyoracle
    z = np.zeros(16)
    from roadsurf_b200 import abi
    pyoracle.load().oracle_layer_depths(15, z.ctypes.data_as(abi.c_double_p))
    assert np.array_equal(np.array(pt.Z[1:17]), z)


CASES = [
    # name, make_case kwargs, post-processing of (arrays, settings)
    ("nlayers 4, fixed output depth", dict(npoints=2, hours=3, seed=41, nlayers=4, settings_kw=dict(tsurf_output_depth=0.05)), None),
    ("nlayers 12, output depth below the grid", dict(npoints=2, hours=3, seed=42, nlayers=12,
                                                     settings_kw=dict(tsurf_output_depth=50.0)), None),
    ("nlayers 20, depth 0 = top layer", dict(npoints=2, hours=3, seed=43, nlayers=20, settings_kw=dict(tsurf_output_depth=0.0)), None),
    ("per-step depth array", dict(npoints=2, hours=3, seed=44), "depth"),
    ("force_tsurf with observations throughout", dict(npoints=2, hours=2, seed=45, analysis_hours=2,
                                                      settings_kw=dict(force_tsurf=1)), "obs_everywhere"),
    ("coupling with a short window, relaxation, sky view", dict(npoints=3, hours=2, seed=46, analysis_hours=2, use_coupling=1,
                                                              use_relaxation=1, sky_view_fraction=1.0,
                                                              settings_kw=dict(coupling_minutes=30)), None),
    ("coupling whose window starts at step 1", dict(npoints=2, hours=1, seed=47, analysis_hours=1, use_coupling=1,
                                                    settings_kw=dict(coupling_minutes=600)), None),
    ("bad input in the middle, NaN input", dict(npoints=3, hours=2, seed=48), "bad"),
    ("summer day with shadows (sun well above the horizon)", dict(npoints=3, hours=6, seed=49, sky_view_fraction=1.0), "summer"),
    ("time step 60 s, coupling + relaxation", dict(npoints=2, hours=3, seed=50, analysis_hours=3, use_coupling=1, use_relaxation=1,
                                                  dt=60.0, settings_kw=dict(coupling_minutes=60)), None),
    ("time step 20 s, plain forecast with sky view", dict(npoints=2, hours=2, seed=51, dt=20.0, sky_view_fraction=1.0), None),
]


@pytest.mark.parametrize("name,kw,post", CASES, ids=[c[0] for c in CASES])
def test_two_restatements_agree_on_the_odd_paths(oracle, name, kw, post):
    """The same N-version check on the paths the golden cases do not take: other layer counts, the output
    depth variants, forced surface temperature, short / clipped coupling windows, failing and NaN inputs,
    a high sun."""
    import datetime as dt
    from roadsurf_b200 import synth
    if post == "summer":
        kw = dict(kw, start=dt.datetime(2019, 6, 21, 3, 0, 0))
    arrays, settings, params, rec = synth.make_case(**kw)
    if post == "depth":
        arrays.Depth[0, :] = 0.03
        arrays.Depth[1, ::2] = 0.5
    if post == "obs_everywhere":
        arrays.TSurfObs[:, :] = arrays.tair - 0.7
    if post == "bad":
        arrays.VZ[0, 100] = 250.0
        arrays.tair[1, 50] = float("nan")
        arrays.PrecPhase[2, :] = 9          # an unknown phase code: falls back to the interpretation
    ref = arrays.copy()
    st, executed = oracle.run_batch(ref, settings, params, nthreads=1)
    total = 0
    for p in range(arrays.npoints):
        out, nsteps, failed = py_restatement.run_point(arrays, settings, params, p)
        total += nsteps
        assert failed == bool(st[p] & 1), (name, p)
        for k, v in out.items():
            assert np.array_equal(np.isnan(v), np.isnan(ref.out[k][p])), (name, p, k)
            d = np.abs(v - ref.out[k][p])
            assert not (np.nanmax(d) >= 1e-9), (name, p, k, int(np.nanargmax(d)), float(np.nanmax(d)))
    assert total == executed, name


from extreme_cases import extreme_cases as _extreme_cases


def test_two_restatements_agree_on_extreme_inputs(oracle):
    for name, arrays, settings, params in _extreme_cases():
        ref = arrays.copy()
        st, executed = oracle.run_batch(ref, settings, params, nthreads=1)
        total = 0
        for p in range(arrays.npoints):
            out, nsteps, failed = py_restatement.run_point(arrays, settings, params, p)
            total += nsteps
            assert failed == bool(st[p] & 1), (name, p)
            for k, v in out.items():
                assert np.array_equal(np.isnan(v), np.isnan(ref.out[k][p])), (name, p, k)
                d = np.abs(v - ref.out[k][p])
                assert not (np.nanmax(d) >= 1e-9), (name, p, k, int(np.nanargmax(d)), float(np.nanmax(d)))
        assert total == executed, name
        print(name, "max snow %.2f ice %.2f water %.2f Ts [%.1f, %.1f] failed %d" % (
            ref.out["SnowOut"].max(), ref.out["IceOut"].max(), ref.out["WaterOut"].max(),
            ref.out["TsurfOut"][ref.out["TsurfOut"] > -9000].min(), ref.out["TsurfOut"].max(), int((st & 1).sum())))
