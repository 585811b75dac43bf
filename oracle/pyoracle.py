"""ctypes loader for the CPU oracle (oracle/roadsurf_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs; never by the roadsurf_b200 package.  PARITY UNPINNED (see roadsurf_oracle.hpp).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from roadsurf_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")


def build(force=False):
    """Compile both oracle flavours with oracle/Makefile (gcc only; a few seconds)."""
    if force:
        subprocess.run(["make", "-C", HERE, "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)


_P = C.POINTER


def _declare(lib):
    OP, IP, IS, IPa, LP = (abi.OutputPointers, abi.InputPointers, abi.InputSettings,
                           abi.InputParameters, abi.LocalParameters)
    lib.oracle_runsimulation.argtypes = [_P(OP), _P(IP), _P(IS), _P(IPa), _P(LP)]
    lib.oracle_runsimulation.restype = None
    lib.oracle_runsimulation_ex.argtypes = [_P(OP), _P(IP), _P(IS), _P(IPa), _P(LP),
                                            _P(C.c_longlong), _P(C.c_longlong), C.c_int]
    lib.oracle_runsimulation_ex.restype = C.c_int
    lib.oracle_trace_point.argtypes = [_P(OP), _P(IP), _P(IS), _P(IPa), _P(LP), abi.c_double_p]
    lib.oracle_trace_point.restype = C.c_int
    lib.oracle_run_batch.argtypes = [C.c_int, _P(_P(OP)), _P(_P(IP)), _P(IS), _P(IPa), _P(_P(LP)),
                                     C.c_int, _P(C.c_int), _P(C.c_longlong)]
    lib.oracle_run_batch.restype = None
    lib.oracle_run_batch_state.argtypes = [C.c_int, _P(_P(OP)), _P(_P(IP)), _P(IS), _P(IPa), _P(_P(LP)),
                                           C.c_int, _P(C.c_int), abi.c_double_p, abi.c_double_p]
    lib.oracle_run_batch_state.restype = None
    lib.oracle_interpolate_example2.argtypes = [abi.c_double_p, abi.c_int_p, C.c_int, C.c_double, C.c_int, C.c_int,
                                                C.c_double, abi.c_double_p]
    lib.oracle_interpolate_example2.restype = None
    lib.oracle_sun_position_batch.argtypes = [abi.c_int_p, C.c_int, abi.c_double_p, abi.c_double_p, C.c_int, C.c_int,
                                              abi.c_double_p, abi.c_double_p]
    lib.oracle_sun_position_batch.restype = None
    lib.oracle_count_ops.argtypes = [_P(OP), _P(IP), _P(IS), _P(IPa), _P(LP), _P(C.c_ulonglong)]
    lib.oracle_count_ops.restype = C.c_longlong
    lib.oracle_layer_depths.argtypes = [C.c_int, abi.c_double_p]
    lib.oracle_ground_constants.argtypes = [_P(IS), _P(IPa)] + [abi.c_double_p] * 4
    lib.oracle_julday.argtypes = [C.c_int] * 3
    lib.oracle_julday.restype = C.c_int
    lib.oracle_jde.argtypes = [C.c_int] * 6
    lib.oracle_jde.restype = C.c_double
    lib.oracle_sun_position.argtypes = [C.c_int] * 6 + [C.c_double, C.c_double, abi.c_double_p,
                                                         abi.c_double_p]
    lib.oracle_sun_position.restype = C.c_int
    lib.oracle_calc_tdew.argtypes = [C.c_double, C.c_double]
    lib.oracle_calc_tdew.restype = C.c_double
    lib.oracle_calc_rh.argtypes = [C.c_double, C.c_double]
    lib.oracle_calc_rh.restype = C.c_double
    lib.oracle_prec_type.argtypes = [_P(IS), _P(IPa), C.c_int, C.c_double, C.c_double, C.c_double,
                                     abi.c_double_p, abi.c_double_p]
    lib.oracle_prec_type.restype = C.c_int
    lib.oracle_boundary_layer.argtypes = [_P(IS), _P(IPa)] + [C.c_double] * 5 + [abi.c_double_p]
    lib.oracle_boundary_layer.restype = C.c_int
    lib.oracle_road_cond.argtypes = [_P(IS), _P(IPa), abi.c_double_p]
    lib.oracle_coupling_control.argtypes = [abi.c_double_p, abi.c_int_p]
    lib.oracle_build_flavour.restype = C.c_char_p
    return lib


_libs = {}


def load(fast=False):
    """Load (building if needed) the parity oracle, or with fast=True the flavour compiled with
    the reference's -Ofast flag set (CPU timing baseline)."""
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name not in _libs:
        path = os.path.join(BUILD, name)
        if not os.path.exists(path):
            build()
        _libs[name] = _declare(C.CDLL(path))
    return _libs[name]


REF_DIR = os.path.join(HERE, "_ref")
_ref = {}


class ReferenceUnavailable(RuntimeError):
    """oracle/_ref cannot be built here; str(e) is the reason (tests skip WITH it)."""


def load_ref(strict=False):
    """The UNMODIFIED reference built by oracle/build_ref.sh (libroadsurf + example1's Simulation.f90,
    symbol `runsimulation`).  Raises ReferenceUnavailable with the reason when no Fortran compiler /
    reference tree exists (the case in this image) and no prebuilt library lies in oracle/_ref/."""
    name = "libroadsurf_ref_strict.so" if strict else "libroadsurf_ref.so"
    if name in _ref:
        return _ref[name]
    path = os.path.join(REF_DIR, name)
    if not os.path.exists(path):
        env = dict(os.environ, REF_FLAVOUR="strict" if strict else "fast")
        r = subprocess.run(["sh", os.path.join(HERE, "build_ref.sh")], capture_output=True, text=True, env=env)
        if r.returncode != 0 or not os.path.exists(path):
            raise ReferenceUnavailable((r.stderr.strip() or r.stdout.strip() or "build_ref.sh failed")
                                       + f" [exit {r.returncode}]")
    lib = C.CDLL(path)
    OP, IP, IS, IPa, LP = (abi.OutputPointers, abi.InputPointers, abi.InputSettings,
                           abi.InputParameters, abi.LocalParameters)
    lib.runsimulation.argtypes = [_P(OP), _P(IP), _P(IS), _P(IPa), _P(LP)]
    lib.runsimulation.restype = None
    _ref[name] = lib
    return lib


def _run_batch_ref(arrays, settings, params, nthreads, strict):
    """Point by point through the reference's own runsimulation (re-entrant: all its state is local to
    the call, examples/example1/src/Simulation.f90:29-46), on `nthreads` Python threads (ctypes drops
    the GIL).  The reference has no status channel: status is derived from the -9999.0 fill."""
    import concurrent.futures as cf
    lib = load_ref(strict)
    ins, outs = arrays.input_pointers(), arrays.output_pointers()

    def one(p):
        lib.runsimulation(C.byref(outs[p]), C.byref(ins[p]), C.byref(settings), C.byref(params),
                          C.byref(arrays.local[p]))
    with cf.ThreadPoolExecutor(max(1, int(nthreads))) as ex:
        list(ex.map(one, range(arrays.npoints)))
    status = np.where(arrays.out["TsurfOut"][:, -1] == -9999.0, 1, 0).astype(np.int32)
    return status, -1


def run_batch(arrays, settings, params, nthreads=1, fast=False, backend="port"):
    """Run every point of a PointArrays through the oracle (mutates arrays' inputs exactly as the
    reference does, fills arrays.out).  Returns (status[npoints], executed_steps).
    backend "port": the C++ restatement (this directory); "ref" / "ref_strict": the unmodified
    reference from oracle/_ref (raises ReferenceUnavailable when it cannot be built)."""
    if backend in ("ref", "ref_strict"):
        return _run_batch_ref(arrays, settings, params, nthreads, backend == "ref_strict")
    lib = load(fast)
    ins = arrays.input_pointers()
    outs = arrays.output_pointers()
    in_ptrs = abi.pointer_arrays(ins, abi.InputPointers)
    out_ptrs = abi.pointer_arrays(outs, abi.OutputPointers)
    loc_ptrs = abi.pointer_arrays(arrays.local, abi.LocalParameters)
    status = np.zeros(arrays.npoints, dtype=np.int32)
    steps = C.c_longlong(0)
    lib.oracle_run_batch(arrays.npoints, out_ptrs, in_ptrs, C.byref(settings), C.byref(params),
                         loc_ptrs, int(nthreads), status.ctypes.data_as(abi.c_int_p),
                         C.byref(steps))
    return status, steps.value


def run_batch_state(arrays, settings, params, nthreads=1):
    """run_batch that also returns the final state: (status, Tmp[npoints, NLayers+2] = ground%Tmp(0:N+1),
    surf[npoints, 10] = TsurfAve, Wat, Snow, Ice, Ice2, Dep, Q2Melt, T4Melt, EvapmmTS, Albedo)."""
    lib = load(False)
    ins, outs = arrays.input_pointers(), arrays.output_pointers()
    in_ptrs = abi.pointer_arrays(ins, abi.InputPointers)
    out_ptrs = abi.pointer_arrays(outs, abi.OutputPointers)
    loc_ptrs = abi.pointer_arrays(arrays.local, abi.LocalParameters)
    status = np.zeros(arrays.npoints, dtype=np.int32)
    tmp = np.zeros((arrays.npoints, settings.NLayers + 2))
    surf = np.zeros((arrays.npoints, 10))
    lib.oracle_run_batch_state(arrays.npoints, out_ptrs, in_ptrs, C.byref(settings), C.byref(params), loc_ptrs,
                               int(nthreads), status.ctypes.data_as(abi.c_int_p), tmp.ctypes.data_as(abi.c_double_p),
                               surf.ctypes.data_as(abi.c_double_p))
    return status, tmp, surf


def interpolate_example2(raw, record_step, dt_secs, sim_len, kind=0, fill=-9999.9):
    """example2's time interpolation (AsciiSource.cpp:223-345) of one raw series [nrec] -> numpy [sim_len].
    kind 0 plain, 1 relative humidity, 2 precipitation, 3 precipitation phase."""
    raw = np.ascontiguousarray(raw, dtype=np.float64)
    rs = np.ascontiguousarray(record_step, dtype=np.int32)
    out = np.empty(int(sim_len))
    load(False).oracle_interpolate_example2(raw.ctypes.data_as(abi.c_double_p), rs.ctypes.data_as(abi.c_int_p), len(rs),
                                            float(dt_secs), int(sim_len), int(kind), float(fill),
                                            out.ctypes.data_as(abi.c_double_p))
    return out


def sun_position_batch(time_fields, lat, lon, nthreads=8):
    """Solar position (src/SunPosition.f90) for every (step, point): time_fields [6, n] int32 -> (elev, azim) [n, npoints]."""
    tf = np.ascontiguousarray(time_fields, dtype=np.int32)
    lat, lon = np.ascontiguousarray(lat, dtype=np.float64), np.ascontiguousarray(lon, dtype=np.float64)
    n, npts = tf.shape[1], lat.shape[0]
    elev, azim = np.empty((n, npts)), np.empty((n, npts))
    load(False).oracle_sun_position_batch(tf.ctypes.data_as(abi.c_int_p), n, lat.ctypes.data_as(abi.c_double_p),
                                          lon.ctypes.data_as(abi.c_double_p), npts, int(nthreads),
                                          elev.ctypes.data_as(abi.c_double_p), azim.ctypes.data_as(abi.c_double_p))
    return elev, azim


def count_ops(arrays, settings, params, point=0):
    """Exact arithmetic-operation counts of one point's run: dict + executed step count."""
    lib = load(False)
    ins = arrays.input_pointers()
    outs = arrays.output_pointers()
    counts = (C.c_ulonglong * 9)()
    steps = lib.oracle_count_ops(C.byref(outs[point]), C.byref(ins[point]), C.byref(settings),
                                 C.byref(params), C.byref(arrays.local[point]), counts)
    names = ("add", "mul", "div", "sqrt", "exp", "log", "trig", "pow", "cmp")
    return dict(zip(names, [int(c) for c in counts])), int(steps)


TRACE_NAMES = ("Ts_in", "Tair", "Prec", "Q2Melt_in", "rain", "snow", "Snow_preRC", "Wat_preRC", "Ice_preRC",
               "Q2Melt_preRC", "T1", "T2", "HStor", "BLCond", "LE", "Evap", "bl_iters", "bl_unstable")


def trace_point(arrays, settings, params, point=0):
    """Per-step internals of one point's oracle run: numpy [sim_len, 18] (columns TRACE_NAMES)."""
    lib = load(False)
    ins = arrays.input_pointers()
    outs = arrays.output_pointers()
    tr = np.zeros((arrays.sim_len, len(TRACE_NAMES)))
    lib.oracle_trace_point(C.byref(outs[point]), C.byref(ins[point]), C.byref(settings), C.byref(params),
                           C.byref(arrays.local[point]), tr.ctypes.data_as(abi.c_double_p))
    return tr
