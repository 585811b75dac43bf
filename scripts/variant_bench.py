"""Times one library variant (ROADSURF_B200_LIBNAME) on the coarse c4-like workload and checks it
against the oracle on a small coupled case."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from roadsurf_b200 import abi, lib, synth, synth_torch
from oracle import pyoracle
from parity import compare

P = int(os.environ.get("RS_POINTS", 227328)); hours = 24; sim_len = 1 + hours * 120
lib.set_model(abi.default_settings(sim_len), abi.default_parameters(30.0))
db = lib.DeviceBatch(P, sim_len, n_records=hours + 2, coarse=True, horizons=True, out_stride=120)
synth_torch.fill_device_batch(db, seed=7)
db.run(); torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); db.run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
res = {"lib": os.environ.get("ROADSURF_B200_LIBNAME", "default"), "regs": lib.last_launch()["regs_per_thread"],
       "ms": round(min(ts), 2), "rate": P * sim_len / min(ts) * 1e3}
# full-res path too (c3-like memory behaviour)
npts = int(os.environ.get("RS_PARITY_POINTS", 2048))
arrays, settings, params, _ = synth.make_case(npts, 24, seed=99, analysis_hours=6, use_coupling=1, use_relaxation=1)
ref = arrays.copy()
sg = lib.run_batch(arrays, settings, params)
so, steps = pyoracle.run_batch(ref, settings, params, nthreads=16)
r = compare(arrays.out, ref.out)
res.update(parity_points=npts, mismatch_fraction=r["mismatch_fraction"], max_dT_matching=r["max_dT_matching"],
           max_dS_matching=r["max_dS_matching"], max_dT_all=r["max_dT_all"], status_equal=bool((sg == so).all()),
           kernel_ms_c3like=round(lib.last_batch_stats()["kernel_ms"], 2))
print(json.dumps(res))
