"""ctypes mirrors of the five interop structs of include/roadsurf_b200.h.

Layouts follow the reference's Bind(C) types (src/InputPointers.f90.inc:4-27,
src/OutputPointers.f90.inc:4-17, src/InputSettings.f90.inc:4-18,
src/InputParameters.f90.inc:4-91, src/LocalParameters.f90.inc:4-15); default values follow the
reference's C++ example (examples/example1/src/InputParameters.h:18-110,
InputParameters.cpp:11-22, InputSettings.h:13-23, LocalParameters.h:17-27).
"""
import ctypes as C
import math

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)

INPUT_DOUBLE_FIELDS = ("c_tair", "c_tdew", "c_VZ", "c_Rhz", "c_prec", "c_SW", "c_LW", "c_SW_dir",
                       "c_LW_net", "c_TSurfObs")
INPUT_TIME_FIELDS = ("c_year", "c_month", "c_day", "c_hour", "c_minute", "c_second")
OUTPUT_FIELDS = ("c_TsurfOut", "c_SnowOut", "c_WaterOut", "c_IceOut", "c_DepositOut", "c_Ice2Out")


class InputPointers(C.Structure):
    _fields_ = [("inputLen", C.c_int),
                ("c_tair", c_double_p), ("c_tdew", c_double_p), ("c_VZ", c_double_p),
                ("c_Rhz", c_double_p), ("c_prec", c_double_p), ("c_SW", c_double_p),
                ("c_LW", c_double_p), ("c_SW_dir", c_double_p), ("c_LW_net", c_double_p),
                ("c_TSurfObs", c_double_p), ("c_PrecPhase", c_int_p),
                ("c_local_horizons", c_double_p), ("c_Depth", c_double_p),
                ("c_year", c_int_p), ("c_month", c_int_p), ("c_day", c_int_p),
                ("c_hour", c_int_p), ("c_minute", c_int_p), ("c_second", c_int_p)]


class OutputPointers(C.Structure):
    _fields_ = [("outputLen", C.c_int)] + [(n, c_double_p) for n in OUTPUT_FIELDS]


class InputSettings(C.Structure):
    _fields_ = [("SimLen", C.c_int), ("use_coupling", C.c_int), ("use_relaxation", C.c_int),
                ("force_tsurf", C.c_int), ("DTSecs", C.c_double), ("tsurfOutputDepth", C.c_double),
                ("NLayers", C.c_int), ("coupling_minutes", C.c_int),
                ("couplingEffectReduction", C.c_double), ("outputStep", C.c_int)]


PARAMETER_NAMES = (
    "NightOn", "NightOff", "CalmLimDay", "CalmLimNgt", "TrfFricNgt", "TrFfricDay",
    "Grav", "SB_Const", "VK_Const", "LVap", "LFus", "WatDens", "SnowDens", "IceDens", "DepDens",
    "WatMHeat", "PorEvaF",
    "ZRefW", "ZRefT", "ZeroDisp", "ZMom", "ZHeat", "Emiss", "Albedo", "Albedo_surroundings",
    "MaxPormms", "TClimG", "DampDpth", "Omega", "AZ", "DampWearF", "AlbDry", "AlbSnow", "vsh1",
    "vsh2", "Poro1", "Poro2", "RhoB1", "RhoB2", "Silt1", "Silt2",
    "freezing_limit_normal", "snow_melting_limit_normal", "ice_melting_limit_normal",
    "frost_melting_limit_normal", "frost_formation_limit_normal", "T4Melt_normal",
    "TLimColdH", "TLimColdL", "WetSnowFormR", "WetSnowMeltR",
    "PLimSnow", "PLimRain", "MaxSnowmms", "MaxDepmms", "MaxIcemms", "MaxExtmms",
    "MissValI", "MissValR", "Snow2IceFac",
    "MinPrecmm", "MinWatmms", "MinSnowmms", "MaxWatmms", "WDampLim", "WWetLim", "WWearLim",
    "MinDepmms", "MinIcemms")


class InputParameters(C.Structure):
    _fields_ = [(n, C.c_double) for n in PARAMETER_NAMES]


class LocalParameters(C.Structure):
    _fields_ = [("tair_relax", C.c_double), ("VZ_relax", C.c_double), ("RH_relax", C.c_double),
                ("couplingIndexI", C.c_int), ("couplingTsurf", C.c_double), ("lat", C.c_double),
                ("lon", C.c_double), ("sky_view", C.c_double), ("InitLenI", C.c_int)]


assert C.sizeof(InputPointers) == 160
assert C.sizeof(OutputPointers) == 56
assert C.sizeof(InputSettings) == 56
assert C.sizeof(InputParameters) == 552 and len(PARAMETER_NAMES) == 69
assert C.sizeof(LocalParameters) == 72
assert InputSettings.force_tsurf.offset == 12 and InputSettings.DTSecs.offset == 16
assert LocalParameters.couplingTsurf.offset == 32 and LocalParameters.InitLenI.offset == 64


def default_settings(sim_len, use_coupling=0, use_relaxation=0, dt=30.0, nlayers=15,
                     coupling_minutes=180, force_tsurf=0, tsurf_output_depth=-9999.9,
                     coupling_effect_reduction=4.0 * 3600, output_step=60):  # noqa: D401
    """examples/example1/src/InputSettings.h:13-23."""
    return InputSettings(SimLen=int(sim_len), use_coupling=int(use_coupling),
                         use_relaxation=int(use_relaxation), force_tsurf=int(force_tsurf),
                         DTSecs=float(dt), tsurfOutputDepth=float(tsurf_output_depth),
                         NLayers=int(nlayers), coupling_minutes=int(coupling_minutes),
                         couplingEffectReduction=float(coupling_effect_reduction),
                         outputStep=int(output_step))


def default_parameters(dt=30.0, **overrides):
    """examples/example1/src/InputParameters.h:18-110 + the derived values of
    InputParameters.cpp:11-22 (which depend on DTSecs)."""
    p = dict(
        NightOn=19.0, NightOff=4.0, CalmLimDay=1.5, CalmLimNgt=0.4, TrfFricNgt=5.0, TrFfricDay=10.0,
        Grav=9.81, SB_Const=5.67e-8, VK_Const=0.4, LVap=2.452e6, LFus=0.334e6, WatDens=999.87,
        SnowDens=100.0, IceDens=920.0, DepDens=920.0, WatMHeat=333000.0, PorEvaF=1.0,
        ZRefW=10.0, ZRefT=2.0, ZeroDisp=0.0, ZMom=0.4, ZHeat=0.001, Emiss=0.95, Albedo=0.10,
        Albedo_surroundings=0.15, MaxPormms=1.0, TClimG=6.4, DampDpth=2.7,
        Omega=2.0 * math.pi / 365.0, AZ=0.6, DampWearF=0.5, AlbDry=0.1, AlbSnow=0.6, vsh1=1.94e6,
        vsh2=1.28e6, Poro1=0.1, Poro2=0.4, RhoB1=2.11, RhoB2=1.6, Silt1=0.1, Silt2=0.8,
        freezing_limit_normal=-0.25, snow_melting_limit_normal=0.25, ice_melting_limit_normal=0.25,
        frost_melting_limit_normal=1.25, frost_formation_limit_normal=0.25, T4Melt_normal=0.25,
        TLimColdH=-19.0, TLimColdL=-21.0, WetSnowFormR=0.1, WetSnowMeltR=0.6,
        PLimSnow=0.3, PLimRain=0.7, MaxSnowmms=100.0, MaxDepmms=2.0, MaxIcemms=50.0, MaxExtmms=1.0,
        MissValI=-9999.0, MissValR=-99.99, Snow2IceFac=0.5)
    unknown = [k for k in overrides if k not in PARAMETER_NAMES]
    if unknown:
        raise KeyError(f"unknown InputParameters field(s): {unknown}")
    p.update({k: v for k, v in overrides.items() if k in p})
    p["MinPrecmm"] = 0.05 * dt / 3600.0
    p["MinWatmms"] = 0.01 * dt / 3600.0
    p["MinSnowmms"] = 0.1 * dt / 3600.0
    p["MaxWatmms"] = p["MaxPormms"] + p["MaxExtmms"]
    p["WDampLim"] = 0.1 * p["MaxPormms"]
    p["WWetLim"] = 0.9 * p["MaxPormms"]
    p["WWearLim"] = 0.1 * p["MaxPormms"]
    p["MinDepmms"] = 0.01 * dt / 3600.0
    p["MinIcemms"] = 0.05 * dt / 3600.0
    p.update(overrides)  # explicit overrides of derived values win
    return InputParameters(**p)


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


class PointArrays:
    """Caller-owned per-point arrays in the reference's host layout (one contiguous array per
    variable per point: examples/example1/src/InputData.cpp:5-50, OutputData.cpp:5-25), held for a
    whole batch as [npoints, sim_len] row-major numpy arrays so that row p is point p's array."""

    def __init__(self, npoints, sim_len):
        self.npoints, self.sim_len = int(npoints), int(sim_len)
        shp = (self.npoints, self.sim_len)
        for n in INPUT_DOUBLE_FIELDS:
            setattr(self, n[2:], np.full(shp, -9999.9))
        self.PrecPhase = np.full(shp, -9999, dtype=np.int32)
        self.local_horizons = np.zeros((self.npoints, 360))
        self.Depth = np.full(shp, -9999.9)
        # time axis: shared by all points unless time_per_point is filled in
        self.time = np.full((6, self.sim_len), -9999, dtype=np.int32)
        self.time_per_point = None
        self.out = {n[2:]: np.full(shp, -9999.0) for n in OUTPUT_FIELDS}
        self.local = (LocalParameters * self.npoints)()

    def input_pointers(self):
        arr = (InputPointers * self.npoints)()
        for p in range(self.npoints):
            ip = arr[p]
            ip.inputLen = self.sim_len
            for n in INPUT_DOUBLE_FIELDS:
                setattr(ip, n, _dp(getattr(self, n[2:])[p]))
            ip.c_PrecPhase = _ip(self.PrecPhase[p])
            ip.c_local_horizons = _dp(self.local_horizons[p])
            ip.c_Depth = _dp(self.Depth[p])
            t = self.time if self.time_per_point is None else self.time_per_point[p]
            for k, n in enumerate(INPUT_TIME_FIELDS):
                setattr(ip, n, _ip(t[k]))
        return arr

    def output_pointers(self):
        arr = (OutputPointers * self.npoints)()
        for p in range(self.npoints):
            op = arr[p]
            op.outputLen = self.sim_len
            for n in OUTPUT_FIELDS:
                setattr(op, n, _dp(self.out[n[2:]][p]))
        return arr

    def copy(self):
        import copy
        other = PointArrays.__new__(PointArrays)
        other.npoints, other.sim_len = self.npoints, self.sim_len
        for n in INPUT_DOUBLE_FIELDS:
            setattr(other, n[2:], getattr(self, n[2:]).copy())
        other.PrecPhase = self.PrecPhase.copy()
        other.local_horizons = self.local_horizons.copy()
        other.Depth = self.Depth.copy()
        other.time = self.time.copy()
        other.time_per_point = None if self.time_per_point is None else self.time_per_point.copy()
        other.out = {k: np.full_like(v, -9999.0) for k, v in self.out.items()}
        other.local = (LocalParameters * self.npoints)()
        C.memmove(other.local, self.local, C.sizeof(self.local))
        del copy
        return other


def pointer_arrays(structs, struct_type):
    """ctypes array of pointers to the elements of a ctypes struct array."""
    n = len(structs)
    ptrs = (C.POINTER(struct_type) * n)()
    for p in range(n):
        ptrs[p] = C.pointer(structs[p])
    return ptrs
