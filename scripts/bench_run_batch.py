"""End-to-end time of the reference-facing batched entry (roadsurf_run_batch, host array-of-pointers
in/out) on a c3-like case, next to the CPU restatement on all host cores."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from roadsurf_b200 import lib, synth
from oracle import pyoracle
from parity import compare
P = int(os.environ.get("RS_POINTS", 20000))
arrays, settings, params, rec = synth.make_case(P, 24, seed=77, analysis_hours=6, use_coupling=1, use_relaxation=1)
ref = arrays.copy()
pb = lib.PreparedBatch(arrays)
pb.run(settings, params)           # warm-up (context creation)
t0 = time.perf_counter(); st = pb.run(settings, params); t_gpu = time.perf_counter() - t0
stats = lib.last_batch_stats()
cores = os.cpu_count()
t0 = time.perf_counter(); so, steps = pyoracle.run_batch(ref, settings, params, nthreads=cores, fast=True); t_cpu = time.perf_counter() - t0
r = compare(arrays.out, ref.out)
n = P * arrays.sim_len
print(json.dumps({"points": P, "sim_len": arrays.sim_len, "gpu_e2e_s": round(t_gpu, 3), "gpu_point_steps_per_s": n / t_gpu,
                  "cpu_s": round(t_cpu, 3), "cpu_point_steps_per_s": n / t_cpu, "cores": cores, "speedup": t_cpu / t_gpu,
                  "stats": {k: (round(v, 1) if isinstance(v, float) else v) for k, v in stats.items()},
                  "mismatch_fraction_vs_fast_build": r["mismatch_fraction"], "max_dT_matching": r["max_dT_matching"]}))
