#!/usr/bin/env python
"""Benchmark of the RoadSurf per-point simulation loop on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload): BASELINE.json config 4, "synthetic 10^7-point km-scale road grid, 24 h
forecast, sharded across 1/2/4/8 B200": every GPU runs one 1.25e6-point shard (10^7 / 8), so at
N = 8 the job is exactly the named configuration ("weak" scaling).  SimLen = 2881 steps of 30 s,
15 ground layers, hourly forcing records interpolated on the device, hourly outputs, 30 % of the
points with sky-view / shadow radiation.  One bench "step" = one full 24 h pass over the shard
(= one launch of the step kernel).

metric  road-point-timesteps/s = points * SimLen / seconds (nominal model steps).
value   kernel-resident throughput: inputs already in HBM when the timed region starts.
e2e     same metric through the C ABI with HOST buffers (roadsurf_run_host_soa): pinned host
        records -> H2D -> kernel -> D2H of the hourly outputs, all inside the timed region.
--impl reference  the reference's CPU path: the C++ restatement of the Fortran loop (oracle/,
        compiled with the reference's own -Ofast flag set; no Fortran compiler exists here, see
        DESIGN.md) on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

POINTS_PER_GPU = 1_250_000
HOURS = 24
METRIC = "road_point_timesteps_per_sec"
UNIT = "point-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--points-per-gpu", type=int, default=POINTS_PER_GPU)
    ap.add_argument("--hours", type=int, default=HOURS)
    ap.add_argument("--cpu-sample-points", type=int, default=0, help="0 = 1024 per host core, <= 8192")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload_config(args, extra=None):
    cfg = {"workload": "c4 shard: synthetic km-scale road grid, 24 h forecast, 10^7/8 points per GPU",
           "points_per_gpu": args.points_per_gpu, "sim_len": 1 + args.hours * 120, "dt_s": 30.0,
           "nlayers": 15, "forcing": "hourly records, device-side linear interpolation",
           "output": "every 120th step (hourly)", "sky_view_fraction": 0.3,
           "coupling": False, "relaxation": False,
           "point_order": "sky-view points scattered at random in the input; gathered per launch by "
                          "roadsurf_order_points (inside the timed step)",
           "l2": "inputs+outputs per step (4.4 GB) exceed the 126 MB L2; no flush needed"}
    if extra:
        cfg.update(extra)
    return cfg


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for k, n in enumerate(names) if any(len(s) > 3 + k and s[3 + k] == "Active" for s in self.samples)]
        power = [float(s[2]) for s in self.samples if len(s) > 2 and s[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "power_w_max": max(power) if power else None}


def cpu_arm(args, rec_sample, settings_hours, steps, warmup, threads):
    """Times the CPU restatement (reference flag set) on `threads` host threads over the sample."""
    from oracle import pyoracle
    from roadsurf_b200 import synth
    arrays, settings, params = synth.case_from_records(rec_sample, settings_hours)
    n = arrays.npoints * arrays.sim_len
    times = []
    for it in range(warmup + steps):
        work = arrays.copy()          # the reference mutates its inputs in place
        t0 = time.perf_counter()
        pyoracle.run_batch(work, settings, params, nthreads=threads, fast=True)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return {"value": n * len(times) / total, "seconds": total, "points": arrays.npoints,
            "sim_len": arrays.sim_len, "ms_per_step": 1e3 * total / len(times)}


def first_points(rec, npts):
    """The first `npts` points of a record sample."""
    from roadsurf_b200 import synth
    sub = synth.Records(npts, rec.nrec)
    for v in synth.RECORD_VARS:
        setattr(sub, v, getattr(rec, v)[:npts].copy())
    sub.lat, sub.lon, sub.sky_view = rec.lat[:npts], rec.lon[:npts], rec.sky_view[:npts]
    sub.horizons, sub.record_step = rec.horizons[:npts], rec.record_step
    return sub


def flops_per_point_step(rec_sample, hours, npts=8):
    """Exact arithmetic-operation count of the reference algorithm on this workload (counting
    scalar in the oracle): each add/mul/div/sqrt/exp/log/trig/pow = 1 flop."""
    from oracle import pyoracle
    from roadsurf_b200 import synth
    import numpy as np
    sub = first_points(rec_sample, npts)
    arrays, settings, params = synth.case_from_records(sub, hours)
    tot, steps = {}, 0
    for p in range(npts):
        c, s = pyoracle.count_ops(arrays, settings, params, p)
        steps += s
        for k, v in c.items():
            tot[k] = tot.get(k, 0) + v
    per = {k: v / steps for k, v in tot.items()}
    flops = sum(per[k] for k in ("add", "mul", "div", "sqrt", "exp", "log", "trig", "pow"))
    return flops, {k: round(v, 2) for k, v in per.items()}


def run_reference(args):
    """The reference's CPU implementation of the path on the box's host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from roadsurf_b200 import synth
    cores = os.cpu_count() or 1
    npts = args.cpu_sample_points or min(8192, 1024 * cores)
    rec = synth.draw_records(npts, args.hours + 2, 20191206, synth.FORECAST_START)
    rec.TSurfObs[:, :] = -9999.9
    r = cpu_arm(args, rec, args.hours, args.steps, args.warmup, cores)
    sample = f"{npts} points x {r['sim_len']} steps per step (same generator and shape as the GPU workload)"
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args),
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "C++ restatement of the Fortran path built with the reference's -Ofast "
                                     "flag set; the gfortran binary cannot be built here (no Fortran compiler)"},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from roadsurf_b200 import build as rs_build
    from roadsurf_b200 import abi, lib, sharding, synth, synth_torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        rs_build.build_library()
    sharding.barrier(dev)
    handle = lib.load()
    launches0 = lib.last_launch()["launches_total"]

    P, hours = args.points_per_gpu, args.hours
    sim_len = 1 + hours * 120
    settings = abi.default_settings(sim_len)
    params = abi.default_parameters(30.0)
    lib.set_model(settings, params)
    db = lib.DeviceBatch(P, sim_len, n_records=hours + 2, coarse=True, horizons=True, out_stride=120)
    synth_torch.fill_device_batch(db, seed=20191206 + rank)
    stream = torch.cuda.current_stream()

    def step():
        # one pass: the library gathers the sky-view points at one end of the launch (the permutation is
        # rebuilt every step here, as a caller with changing statics would; it is static per grid)
        db.build_order(stream)
        db.run(stream)

    for _ in range(max(args.warmup, 0)):
        step()
    torch.cuda.synchronize()
    sharding.barrier(dev)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    launches_before = lib.last_launch()["launches_total"] if args.warmup > 0 else launches0
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    sharding.barrier(dev)
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    kernel_ms = ms_total / args.steps                      # one step kernel launch (+ the order kernel) per step
    ms_max = sharding.reduce_over_ranks(ms_total, "max", dev)
    total_points = sharding.reduce_over_ranks(P, "sum", dev)
    value = total_points * sim_len * args.steps / (ms_max * 1e-3)
    cnt = db.counters.cpu().numpy()
    failed_points = int(cnt[lib.CNT_FAILED_POINTS])
    bl_per_step = float(cnt[lib.CNT_BL_ITERATIONS]) / max(1.0, float(cnt[lib.CNT_EXECUTED_STEPS]))
    launch = lib.last_launch()
    gpu_launches = launch["launches_total"] - launches_before   # per step: order kernel, solar table, step kernel

    # ---- end to end through the C ABI with host buffers ------------------------------------------
    e2e = None
    if not args.no_e2e:
        pin = dict(pin_memory=True)
        h_forcing = torch.empty((db.n_records, db.nvar, P), dtype=torch.float64, **pin)
        h_forcing.copy_(db.forcing[:, :, :P])
        h_local = torch.empty((lib.L_NLOCAL, P), dtype=torch.float64, **pin).copy_(db.local[:, :P])
        h_hor = torch.empty((360, P), dtype=torch.float64, **pin).copy_(db.horizons[:, :P])
        h_tf = db.time_fields.cpu()
        h_rs = db.record_step.cpu()
        h_out = torch.empty((lib.O_NVAR, db.n_out, P), dtype=torch.float64, **pin)
        h_status = torch.empty(P, dtype=torch.int32, **pin)

        def one():
            lib.run_host_soa(settings, params, h_forcing, h_tf, h_local, h_out, record_step=h_rs,
                             horizons=h_hor, status=h_status, out_stride=120, ngpus=1)
        one()                                              # warm-up: sizes the device work space
        torch.cuda.synchronize()
        sharding.barrier(dev)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one()
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        sharding.barrier(dev)
        st = lib.last_batch_stats()
        t_max = sharding.reduce_over_ranks(t_e2e, "max", dev)
        same = bool(torch.equal(h_out.to(dev), db.out[:, :, :P]))
        e2e = {"value": total_points * sim_len * args.steps / t_max, "unit": UNIT,
               "h2d_bytes_per_step": int(st["h2d_bytes"]), "d2h_bytes_per_step": int(st["d2h_bytes"]),
               "ms_per_step": 1e3 * t_max / args.steps, "api": "roadsurf_run_host_soa (pinned host buffers)",
               "breakdown_ms": {"h2d": st["h2d_ms"], "kernel": st["kernel_ms"], "d2h": st["d2h_ms"],
                                "chunks": st["groups"]},
               "outputs_equal_device_run": same}
        gpu_launches += st["kernel_launches"] * args.steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel ---------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = (db.forcing[:, :, :P].numel() + db.out[:, :, :P].numel() + db.local[:, :P].numel() +
                 db.horizons[:, :P].numel()) * 8 + P * 4
    hbm_achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    fp64_peak = lib.measure_fp64_tflops(40000)
    rec_sample = synth_torch.records_sample(db, args.cpu_sample_points or min(8192, 1024 * (os.cpu_count() or 1)))
    flops, per = flops_per_point_step(rec_sample, hours)
    fp64_achieved = flops * P * sim_len / (kernel_ms * 1e-3) / 1e12
    # DRAM traffic of this kernel from the committed ncu --set full capture (profiles/r01_h_libm_exact_ncu_summary.txt:
    # dram__bytes_read.sum + dram__bytes_write.sum = 943.6 MB for 227 328 points x 2881 steps), scaled to this launch
    traffic = 943.588352e6 / (227328 * 2881) * P * sim_len if (P >= 151552 and hours == 24) else None
    roofline = {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": hbm_achieved / hbm_peak, "traffic": traffic,
                "traffic_source": "ncu capture of the same kernel variant at 227328 points, scaled by point-steps",
                "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
                "algorithmic_bytes_per_point_step": alg_bytes / (P * sim_len),
                "note": "coarse-forcing mode moves ~1.2 B per point-step: this kernel is bound by the fp64 "
                        "pipe (see roofline_fp64; ncu: fp64 pipe 54 % busy, issue slots 59 %), not by HBM"}
    roofline_fp64 = {"bound": "fp64", "achieved": fp64_achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": fp64_achieved / fp64_peak,
                     "peak_source": "DFMA micro-benchmark run live by this script (FMA = 2 flop)",
                     "algorithmic_flops_per_point_step": flops, "ops_per_point_step": per,
                     "note": "algorithmic flops count div/sqrt/exp/log/trig as 1 each; in issued fp64 "
                             "instructions they cost ~10-40, see profiles/ for the pipe utilisation"}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        r = cpu_arm(args, rec_sample, hours, 1, 0, cores)
        r1 = cpu_arm(args, first_points(rec_sample, min(512, rec_sample.npoints)), hours, 1, 0, 1)
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port",
                        "one_core_value": r1["value"],
                        "sample": f"first {r['points']} points of the GPU workload x {r['sim_len']} steps, "
                                  f"{r['seconds']:.1f} s wall on {cores} threads",
                        "note": "C++ restatement of the Fortran path, reference -Ofast flag set (no gfortran here)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, {"kernel_regs": launch["regs_per_thread"], "grid": launch["grid"],
                                             "block": launch["block"], "bl_iterations_per_step": bl_per_step,
                                             "failed_points": failed_points}),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(gpu_launches),
            "roofline": roofline, "roofline_fp64": roofline_fp64, "cpu_baseline": cpu_baseline}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
