"""Step latency of warps that are alone on their scheduler (small batches, c3's stragglers): kernel time per model
step for (a) one warp, plain forecast, full-resolution forcing; (b) 401 coupled points x 8881 steps (example1's
shape).  Run once plain, then under  ncu --set full -k regex:rs_run_kernel  for the per-warp stall breakdown."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roadsurf_b200 import lib, synth

res = {}
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, npts, fh, kw in (("one_warp_plain", 32, 24, dict()),
                           ("four_warps_plain", 128, 24, dict()),
                           ("c1c2_401_coupled", 401, 26, dict(analysis_hours=48, use_coupling=1, use_relaxation=1))):
    if only and name != only: continue
    arrays, settings, params, rec = synth.make_case(npts, fh, seed=20191203, **kw)
    for rep in range(2):
        work = arrays.copy()
        lib.run_batch(work, settings, params)
    st = lib.last_batch_stats()
    res[name] = {"points": npts, "sim_len": arrays.sim_len, "kernel_ms": round(st["kernel_ms"], 3),
                 "executed_steps": st["executed_steps"], "launches": st["kernel_launches"],
                 "us_per_nominal_step": round(st["kernel_ms"] * 1e3 / arrays.sim_len, 3), "launch": lib.last_launch()}
print(json.dumps(res))
