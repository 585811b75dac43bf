"""Known-answer tests that pin the CPU oracle.

The reference ships no tests or golden vectors, so the pins are: closed forms
of the model equations evaluated independently here in numpy, published astronomical values
(Meeus, Astronomical Algorithms, ch. 7) and hand-traced storage transitions.
"""
import ctypes as C
import math

import numpy as np
import pytest

from roadsurf_b200 import abi


@pytest.fixture(scope="module")
def olib(oracle):
    return oracle.load()


@pytest.fixture(scope="module")
def sp():
    return abi.default_settings(10), abi.default_parameters(30.0)


def test_layer_depths_closed_form(olib):
    """src/Initialization.f90:227-233: z(i+1) = z(i) + 0.0103*1.4**(i-1) + 0.02."""
    z = np.zeros(16)
    olib.oracle_layer_depths(15, z.ctypes.data_as(abi.c_double_p))
    expect = np.concatenate([[0.0], np.cumsum(0.0103 * 1.4 ** np.arange(15) + 0.02)])
    assert np.allclose(z, expect, atol=5e-7)  # REAL(4) literals move the depths by < 1e-6 m
    table = [0, 0.0303, 0.0647, 0.1049, 0.1532, 0.2127, 0.2881, 0.3857, 0.5143, 0.6863, 0.9191, 1.2370,
             1.6741, 2.2781, 3.1156, 4.2801]
    assert np.allclose(z, table, atol=6e-5)
    assert z[1] != 0.0103 + 0.02  # single-precision literals really are in effect


@pytest.mark.parametrize("ymd,expected", [((2019, 12, 2), 336), ((2019, 1, 1), 1), ((2020, 3, 1), 61),
                                          ((2019, 3, 1), 60), ((2020, 12, 31), 366), ((1900, 3, 1), 60),
                                          ((2000, 3, 1), 61)])
def test_julday(olib, ymd, expected):
    assert olib.oracle_julday(*ymd) == expected


@pytest.mark.parametrize("ymdhms,expected", [((2000, 1, 1, 12, 0, 0), 2451545.0),
                                             ((1999, 1, 1, 0, 0, 0), 2451179.5),
                                             ((1987, 1, 27, 0, 0, 0), 2446822.5),
                                             ((1987, 6, 19, 12, 0, 0), 2446966.0),
                                             ((1957, 10, 4, 19, 26, 24), 2436116.31),
                                             ((2019, 12, 2, 0, 0, 0), 2458819.5)])
def test_julian_ephemeris_day_meeus_examples(olib, ymdhms, expected):
    """Meeus ch. 7 worked examples; the day fraction is single precision in the reference."""
    assert abs(olib.oracle_jde(*ymdhms) - expected) < 3e-6


def _noaa_sun(y, mo, d, h, mi, s, lat, lon):
    """Independent solar elevation/azimuth (NOAA solar calculator equations, degrees)."""
    a = (14 - mo) // 12
    yy = y + 4800 - a
    mm = mo + 12 * a - 3
    jdn = d + (153 * mm + 2) // 5 + 365 * yy + yy // 4 - yy // 100 + yy // 400 - 32045
    jd = jdn + (h - 12) / 24.0 + mi / 1440.0 + s / 86400.0
    t = (jd - 2451545.0) / 36525.0
    l0 = (280.46646 + t * (36000.76983 + t * 0.0003032)) % 360
    m = 357.52911 + t * (35999.05029 - 0.0001537 * t)
    e = 0.016708634 - t * (0.000042037 + 0.0000001267 * t)
    c = (math.sin(math.radians(m)) * (1.914602 - t * (0.004817 + 0.000014 * t)) +
         math.sin(math.radians(2 * m)) * (0.019993 - 0.000101 * t) + math.sin(math.radians(3 * m)) * 0.000289)
    lam = l0 + c - 0.00569 - 0.00478 * math.sin(math.radians(125.04 - 1934.136 * t))
    eps0 = 23 + (26 + (21.448 - t * (46.815 + t * (0.00059 - t * 0.001813))) / 60) / 60
    eps = eps0 + 0.00256 * math.cos(math.radians(125.04 - 1934.136 * t))
    decl = math.degrees(math.asin(math.sin(math.radians(eps)) * math.sin(math.radians(lam))))
    vy = math.tan(math.radians(eps / 2)) ** 2
    eot = 4 * math.degrees(vy * math.sin(2 * math.radians(l0)) - 2 * e * math.sin(math.radians(m)) +
                           4 * e * vy * math.sin(math.radians(m)) * math.cos(2 * math.radians(l0)) -
                           0.5 * vy * vy * math.sin(4 * math.radians(l0)) -
                           1.25 * e * e * math.sin(2 * math.radians(m)))
    tst = (h * 60 + mi + s / 60.0 + eot + 4 * lon) % 1440
    ha = tst / 4 - 180 if tst / 4 >= 0 else tst / 4 + 180
    cz = (math.sin(math.radians(lat)) * math.sin(math.radians(decl)) +
          math.cos(math.radians(lat)) * math.cos(math.radians(decl)) * math.cos(math.radians(ha)))
    zen = math.degrees(math.acos(max(-1, min(1, cz))))
    az_c = ((math.sin(math.radians(lat)) * math.cos(math.radians(zen))) - math.sin(math.radians(decl))) / (
        math.cos(math.radians(lat)) * math.sin(math.radians(zen)))
    az = math.degrees(math.acos(max(-1, min(1, az_c))))
    az = (az + 180) % 360 if ha > 0 else (540 - az) % 360
    return 90 - zen, az


@pytest.mark.parametrize("when,lat,lon", [((2019, 6, 21, 10, 20, 0), 60.17, 24.94),
                                          ((2019, 6, 21, 4, 0, 0), 60.17, 24.94),
                                          ((2019, 6, 21, 16, 30, 30), 65.01, 25.47),
                                          ((2019, 12, 2, 10, 10, 0), 60.17, 24.94),
                                          ((2019, 3, 20, 13, 0, 0), 69.9, 27.0),
                                          ((2020, 9, 1, 7, 45, 0), 61.5, 23.8)])
def test_sun_position_against_independent_algorithm(olib, when, lat, lon):
    elev, azim = C.c_double(), C.c_double()
    stop = olib.oracle_sun_position(*when, lat, lon, C.byref(elev), C.byref(azim))
    e_ref, a_ref = _noaa_sun(*when, lat, lon)
    assert stop == 0 and e_ref > 0
    # 360.98564736629 is a REAL(4) literal in the reference: ~0.07 deg of hour angle after 20 years
    assert abs(elev.value - e_ref) < 0.15
    assert abs((azim.value - a_ref + 180) % 360 - 180) < 0.5


def test_sun_below_horizon_returns_missing(olib):
    elev, azim = C.c_double(), C.c_double()
    olib.oracle_sun_position(2019, 12, 2, 22, 0, 0, 60.17, 24.94, C.byref(elev), C.byref(azim))
    assert elev.value == pytest.approx(-9999.9, abs=1e-3) and azim.value == pytest.approx(-9999.9, abs=1e-3)
    # polar night at 69.9N on 2 December, even at local noon
    olib.oracle_sun_position(2019, 12, 2, 10, 10, 0, 69.9, 27.0, C.byref(elev), C.byref(azim))
    assert elev.value < -9999


def test_tdew_rh_identity(olib):
    """CalcRh(T, CalcTDew(T, RH)) == RH (src/InputOutput.f90:202-268)."""
    for t in (-25.0, -3.3, -0.01, 0.0, 2.5, 18.0):
        for rh in (35.0, 60.0, 88.8, 99.0):
            td = olib.oracle_calc_tdew(t, rh)
            assert td <= t + 1e-9
            # 0.01 is a REAL(4) literal in CalcTDew, 100.0 is exact in CalcRh: the round trip returns
            # rh * (0.01f * 100), which also pins the single-precision literal handling
            assert olib.oracle_calc_rh(t, td) == pytest.approx(rh * float(np.float32(0.01)) * 100.0, rel=1e-11)
    assert olib.oracle_calc_rh(5.0, 7.0) == 100.0  # clamped


def _prec(olib, sp, phase, mmh, tair, rh):
    rain, snow = C.c_double(), C.c_double()
    t = olib.oracle_prec_type(C.byref(sp[0]), C.byref(sp[1]), phase, mmh, tair, rh, C.byref(rain), C.byref(snow))
    return t, rain.value, snow.value


def test_precipitation_phase_decode(olib, sp):
    """src/Cond.f90:143-249."""
    step = 1.2 / 3600 * 30.0
    for ph in (0, 1, 4, 5):
        assert _prec(olib, sp, ph, 1.2, -5.0, 90.0) == (1, step, 0.0)
    assert _prec(olib, sp, 2, 1.2, 5.0, 90.0) == (2, step / 2, step / 2)
    for ph in (3, 6):
        assert _prec(olib, sp, ph, 1.2, 5.0, 90.0) == (3, 0.0, step)
    # below the minimum amount (0.05 mm/h): nothing, whatever the phase
    assert _prec(olib, sp, 1, 0.04, 5.0, 90.0) == (-1, 0.0, 0.0)
    assert _prec(olib, sp, -9999, 0.04, 5.0, 90.0) == (-1, 0.0, 0.0)
    # interpretation: PRain = 1/(1+exp(22 - 2.7 T - 0.2 RH)); < 0.3 snow, > 0.7 rain, else sleet
    for tair, rh in ((-3.0, 90.0), (0.5, 95.0), (0.8, 100.0), (3.0, 80.0), (1.5, 90.0)):
        prain = 1.0 / (1.0 + math.exp(22.0 - 2.7 * tair - 0.2 * rh))
        expect = 3 if prain < 0.3 else (1 if prain > 0.7 else 2)
        assert _prec(olib, sp, -9999, 1.2, tair, rh)[0] == expect
        assert _prec(olib, sp, 7, 1.2, tair, rh)[0] == expect  # unknown code -> interpretation


def test_boundary_layer_neutral_closed_form(olib, sp):
    """Ts == Tair: stability 0, psi = 0, BLCond = rho*c*k*U*/logCond, five iterations
    (src/BoundaryLayer.f90:64-96)."""
    p = sp[1]
    io = (C.c_double * 3)()
    tair, vz = 2.0, 3.0
    iters = olib.oracle_boundary_layer(C.byref(sp[0]), C.byref(p), tair, vz, 100.0, tair, 0.5, io)
    tak = tair + 273.15
    rho = 100000.0 / (287.05 * tak)
    cp = 1005.0 + (tak - 250.0) ** 2 / 3364.0
    ustar = p.VK_Const * vz / math.log((p.ZRefW + p.ZMom) / p.ZMom)
    blc = rho * cp * p.VK_Const * ustar / math.log((p.ZRefW + p.ZHeat) / p.ZHeat)
    assert iters == 5
    assert io[0] == pytest.approx(blc, rel=1e-6)
    # saturated air at surface temperature: no latent flux (0.01 is REAL(4): RH 100 -> 0.99999998)
    assert io[1] == pytest.approx(0.0, abs=1e-3)


def test_boundary_layer_stable_needs_more_iterations(olib, sp):
    io = (C.c_double * 3)()
    it_calm = olib.oracle_boundary_layer(C.byref(sp[0]), C.byref(sp[1]), 5.0, 0.4, 70.0, -5.0, 0.0, io)
    it_windy = olib.oracle_boundary_layer(C.byref(sp[0]), C.byref(sp[1]), 5.0, 8.0, 70.0, 4.0, 0.0, io)
    assert 5 <= it_windy <= it_calm <= 40
    assert io[1] <= 0.0 or io[1] == 0.0  # dry surface: positive latent flux is suppressed


def _road(olib, sp, **kw):
    st = dict(Ts=0.0, Wat=0.0, Snow=0.0, Ice=0.0, Ice2=0.0, Dep=0.0, Q2Melt=0.0, T4Melt=0.25, Evap=0.0, Alb=0.1)
    st.update(kw)
    keys = ("Ts", "Wat", "Snow", "Ice", "Ice2", "Dep", "Q2Melt", "T4Melt", "Evap", "Alb")
    arr = (C.c_double * 10)(*[st[k] for k in keys])
    olib.oracle_road_cond(C.byref(sp[0]), C.byref(sp[1]), arr)
    return dict(zip(keys, list(arr)))


TPH = 30.0 / 3600.0


def test_storage_water_freezes_below_limit(olib, sp):
    """Hand trace of src/Cond.f90:69-103 + src/Storage.f90:33-84,199-267 for Ts=-1, 0.5 mm water."""
    r = _road(olib, sp, Ts=-1.0, Wat=0.5)
    wat_after_wear = 0.5 - 0.5 * (10 * max(0.145 * 0.5, 0.06) * TPH)
    ice = wat_after_wear - 0.01 * TPH
    assert r["Wat"] == 0.0 and r["Snow"] == 0.0
    assert r["Ice"] == pytest.approx(ice, rel=1e-6) and r["Ice2"] == pytest.approx(ice, rel=1e-6)
    assert r["T4Melt"] == 0.25
    assert r["Q2Melt"] == pytest.approx(333000.0 * 999.87 * (r["Ice"] / 1000.0) / 30.0, rel=1e-12)
    assert r["Alb"] == pytest.approx(0.1 + (r["Ice"] / 1.5) * 0.5, rel=1e-9)


def test_storage_condensation_forms_deposit_and_warm_surface_melts_it(olib, sp):
    r = _road(olib, sp, Ts=-3.0, Evap=-0.002)
    assert r["Dep"] == pytest.approx(0.002 - 0.01 * TPH, rel=1e-6) and r["Wat"] == 0.0
    r = _road(olib, sp, Ts=2.0, Dep=0.5)
    assert r["Dep"] == 0.0 and r["Wat"] == pytest.approx(0.5, rel=1e-12)  # melted after the water wear
    r = _road(olib, sp, Ts=-3.0, Dep=2.5)  # overflow above MaxDepmms goes to water
    assert r["Dep"] == 2.0 and r["Wat"] > 0.0


def test_storage_snow_wears_to_ice_with_the_overridden_factor(olib, sp):
    """Snow2IceFac is overwritten with 0.25/(0.2+0.25) in single precision (src/Cond.f90:86); the
    input parameter 0.5 is never used."""
    r = _road(olib, sp, Ts=-5.0, Snow=1.0)
    tran = 0.45 * 1.0 * TPH
    fac = float(np.float32(0.25) / (np.float32(0.2) + np.float32(0.25)))
    assert r["Snow"] == pytest.approx(1.0 - tran, rel=1e-6)
    assert r["Ice"] == pytest.approx(fac * tran - 0.01 * TPH, rel=1e-5)
    assert r["Ice2"] == pytest.approx(fac * tran - 0.01 * TPH, rel=1e-5)
    assert r["Alb"] == 0.6  # snow albedo
    # thin snow wears three times faster
    r = _road(olib, sp, Ts=-5.0, Snow=0.1)
    assert r["Snow"] == pytest.approx(0.1 - 3 * 0.45 * 0.1 * TPH, rel=1e-6)


def test_storage_melt_consumes_q2melt(olib, sp):
    q = 50.0
    r = _road(olib, sp, Ts=0.3, Ice=1.0, Ice2=1.0, Q2Melt=q)
    melted = 1000.0 * q * 30.0 / (333000.0 * 999.87)
    assert r["Ice"] == pytest.approx(1.0 - melted - max(0.319 * 1.0, 0.01) * TPH, rel=1e-6)
    assert r["Wat"] == pytest.approx(melted, rel=1e-9)
    r = _road(olib, sp, Ts=0.2, Ice=1.0, Ice2=1.0, Q2Melt=q)  # below the melt limit: wear only
    assert r["Wat"] == 0.0


def test_storage_small_amounts_are_cut_and_large_clamped(olib, sp):
    r = _road(olib, sp, Ts=5.0, Wat=0.00005)
    assert r["Wat"] == 0.0
    r = _road(olib, sp, Ts=-5.0, Ice=80.0, Ice2=80.0)
    assert r["Ice"] == 50.0 and r["Ice2"] == 50.0
    r = _road(olib, sp, Ts=-5.0, Snow=150.0)
    assert r["Snow"] == pytest.approx(150.0 - 0.45 * 150.0 * TPH - 50.0, rel=1e-9)


def _couple(olib, c, flags):
    arr = (C.c_double * 13)(*c)
    fl = (C.c_int * 3)(*flags)
    olib.oracle_coupling_control(arr, fl)
    return list(arr), list(fl)


def test_coupling_control_branch_table(olib):
    """src/Coupling.f90:292-481.  c = [Ts, obs, RadCoeff, RadCoeffPrev, TsNA, TsNB, RcNA, RcNB, SwCof, LwCof,
    SWcorr, LWcorr, TsEnd1]; flags = [iterations, failed, again]."""
    base = [0.0, 0.0, 1.0, 1.0, -9999.0, -9999.0, -9999.0, -9999.0, 1.0, 1.0, 0.0, 0.0, 0.0]
    # too warm, no bracket yet: halve the coefficient and restart
    c, f = _couple(olib, [2.0, 1.0] + base[2:], [0, 0, 0])
    assert f == [1, 0, 1] and c[2] == 0.5 and c[3] == 0.5
    assert c[4] == pytest.approx(2.0 + 273.16, abs=1e-4) and c[6] == 1.0 and c[5] == -9999.0
    assert c[12] == pytest.approx(2.0 + 273.16, abs=1e-4)  # first guess remembered in Kelvin
    # too cold, no bracket: double
    c, f = _couple(olib, [0.0, 1.0] + base[2:], [0, 0, 0])
    assert f == [1, 0, 1] and c[2] == 2.0 and c[5] == pytest.approx(273.16, abs=1e-4)
    # bracketed: secant step between the nearest guesses
    k = 273.16
    c, f = _couple(olib, [1.5, 1.0, 0.5, 0.5, -9999.0, 0.2 + k, -9999.0, 0.25, 0.5, 1.0, 0.0, 0.0, 0.0], [2, 0, 0])
    above, below = 1.5 - 1.0, 1.0 - 0.2
    assert f == [3, 0, 1]
    assert c[2] == pytest.approx(0.5 - above / (above + below) * (0.5 - 0.25), rel=1e-6)
    # within 0.1 K: success, corrections kept, iteration counter back to 0
    c, f = _couple(olib, [1.05, 1.0, 0.7, 0.7, 5.0 + k, -3.0 + k, 1.0, 0.5, 0.7, 1.0, 0.0, 0.0, 0.0], [4, 0, 0])
    assert f == [0, 0, 0] and c[10] == pytest.approx(-0.3) and c[11] == 0.0 and c[2] == 1.0
    assert c[4] == -9999.0 and c[5] == -9999.0
    # coefficient underflow: coupling fails, one more pass with coefficient 1
    c, f = _couple(olib, [2.0, 1.0, 0.015, 0.015] + base[4:], [6, 0, 0])
    assert f == [7, 1, 1] and c[2] == 1.0 and c[8] == 1.0 and c[9] == 1.0
    # 25 iterations: give up; restart only if the first guess was closer
    c, f = _couple(olib, [3.0, 1.0, 0.3, 0.3] + base[4:12] + [1.5 + k], [25, 0, 0])
    assert f == [26, 1, 1]
    c, f = _couple(olib, [1.2, 1.0, 0.3, 0.3] + base[4:12] + [4.0 + k], [25, 0, 0])
    assert f == [26, 1, 0]
    # abnormal temperature
    c, f = _couple(olib, [150.0, 1.0] + base[2:], [1, 0, 0])
    assert f == [2, 1, 1]
    # the Kelvin round trip leaves Ts within rounding of its input
    assert abs(c[0] - 150.0) < 1e-10


def test_reproducibility_band_parity_vs_reference_flag_build(oracle):
    """The same restatement compiled with the reference's -Ofast flag set (reassociation,
    reciprocal math) against the strict build: measures how reproducible the reference itself is.
    Matching points agree far inside the 1e-3 tolerance; the rest are threshold flips."""
    from parity import compare
    from roadsurf_b200 import synth
    arrays, settings, params, _ = synth.make_case(48, 24, seed=5, analysis_hours=6, use_coupling=1,
                                                   use_relaxation=1)
    fast = arrays.copy()
    oracle.run_batch(arrays, settings, params, nthreads=4, fast=False)
    oracle.run_batch(fast, settings, params, nthreads=4, fast=True)
    r = compare(arrays.out, fast.out)
    assert r["max_dT_matching"] < 1e-3 and r["max_dS_matching"] < 1e-3
    assert r["mismatch_fraction"] <= 0.25, r


def test_operation_counts_are_plausible(oracle):
    """Exact per-step operation counts from the counting scalar (roofline denominator)."""
    from roadsurf_b200 import synth
    arrays, settings, params, _ = synth.make_case(1, 6, seed=3, sky_view_fraction=0.0)
    counts, steps = oracle.count_ops(arrays, settings, params, 0)
    per = {k: v / steps for k, v in counts.items()}
    flops = sum(per[k] for k in ("add", "mul", "div", "sqrt", "exp", "log", "trig", "pow"))
    assert steps == arrays.sim_len
    assert 500 < flops < 1100, per      # SURVEY.md section 8d static estimate: ~715
    # the reference also computes values it never reads (HS(2:N), GCond: 2N divides per step)
    assert 60 < per["div"] < 130 and per["exp"] >= 2 and per["sqrt"] == pytest.approx(per["log"], abs=0.05)
