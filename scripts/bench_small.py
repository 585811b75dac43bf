"""Small batches (BASELINE configs 1-2): example1's 401 stations x 8881 steps (48 h analysis + 26 h forecast,
coupling + relaxation) through the reference-facing entries -- roadsurf_run_batch (one call) and runsimulation
(one point; and 401 points from a pool of 16 host threads, combined into batches by the library) -- with the
16-core CPU restatement beside it.  A run this small cannot fill a B200: it sits at the latency of one warp per
32 points; the numbers say what a user of an unchanged main gets."""
import json, os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roadsurf_b200 import lib, synth
from oracle import pyoracle

arrays, settings, params, rec = synth.make_case(401, 26, seed=20191203, analysis_hours=48, use_coupling=1, use_relaxation=1)
res = {"points": 401, "sim_len": arrays.sim_len}
work = arrays.copy(); lib.run_batch(work, settings, params)          # warm-up (allocations)
ts = []
for _ in range(3):
    # the ctypes argument arrays are built outside the timed call (a C++ main has them anyway; building 401 structs
    # in Python costs ~35 ms, more than everything but the kernel)
    work = arrays.copy(); pb = lib.PreparedBatch(work)
    t0 = time.perf_counter(); st = pb.run(settings, params); ts.append(time.perf_counter() - t0)
stats = lib.last_batch_stats()
res["run_batch_breakdown_ms"] = {k: round(stats[k], 2) for k in ("pack_ms", "h2d_ms", "kernel_ms", "d2h_ms", "unpack_ms", "wall_ms")}
res["run_batch_wall_ms"] = round(min(ts) * 1e3, 1)
res["run_batch_kernel_ms"] = round(stats["kernel_ms"], 1)
res["run_batch_point_steps_per_s"] = 401 * arrays.sim_len / min(ts)
ref = arrays.copy(); t0 = time.perf_counter(); st_cpu, _ = pyoracle.run_batch(ref, settings, params, nthreads=os.cpu_count(), fast=True)
res["cpu_port_wall_ms"] = round((time.perf_counter() - t0) * 1e3, 1); res["cpu_threads"] = os.cpu_count()
ref2 = arrays.copy(); pyoracle.run_batch(ref2, settings, params, nthreads=os.cpu_count())
res["bit_identical_to_oracle"] = bool(all(np.array_equal(work.out[k], ref2.out[k], equal_nan=True) for k in ref2.out))
import ctypes as C
one = arrays.copy(); i1, o1, L = one.input_pointers(), one.output_pointers(), lib.load()
t0 = time.perf_counter()
L.runsimulation(C.byref(o1[0]), C.byref(i1[0]), C.byref(settings), C.byref(params), C.byref(one.local[0]))
res["runsimulation_one_point_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
pool = arrays.copy(); c0 = lib.runsimulation_counters()
ins, outs = pool.input_pointers(), pool.output_pointers()   # built once: the pool threads only call
def worker(ids):
    for p in ids: L.runsimulation(C.byref(outs[p]), C.byref(ins[p]), C.byref(settings), C.byref(params), C.byref(pool.local[p]))
t0 = time.perf_counter()
th = [threading.Thread(target=worker, args=(range(k, 401, 16),)) for k in range(16)]
[t.start() for t in th]; [t.join() for t in th]
res["runsimulation_401_points_16_threads_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
c1 = lib.runsimulation_counters(); res["runsimulation_batches"] = c1[1] - c0[1]
res["pool_equal_to_batch"] = bool(all(np.array_equal(pool.out[k], work.out[k], equal_nan=True) for k in work.out))
res["launch"] = lib.last_launch()
print(json.dumps(res))
