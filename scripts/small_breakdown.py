"""Where the wall time of small reference-facing calls goes (warmed): roadsurf_last_batch_stats of roadsurf_run_batch on
401 x 8881 coupled points and of a single-point runsimulation."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roadsurf_b200 import lib, synth
arrays, settings, params, rec = synth.make_case(401, 26, seed=20191203, analysis_hours=48, use_coupling=1, use_relaxation=1)
out = {}
for rep in range(3):
    w = arrays.copy(); t0 = time.perf_counter(); lib.run_batch(w, settings, params); wall = (time.perf_counter() - t0) * 1e3
    out["run_batch_401"] = dict(lib.last_batch_stats(), measured_wall_ms=round(wall, 2))
for rep in range(3):
    w = arrays.copy(); t0 = time.perf_counter(); lib.runsimulation(w, settings, params, point=7); wall = (time.perf_counter() - t0) * 1e3
    out["runsimulation_1"] = dict(lib.last_batch_stats(), measured_wall_ms=round(wall, 2))
print(json.dumps(out))
