"""Same-box A/B of the step kernel's two bodies (option "latency_body": 0 = throughput body, 1 = latency body) on
latency-bound launches: kernel time of one warp, four warps, and the 401-point coupled case; modes alternate
within one process, minimum of three runs each."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roadsurf_b200 import lib, synth

res = {"lib": os.environ.get("ROADSURF_B200_LIBNAME", "default")}
for name, npts, fh, kw in (("one_warp_plain", 32, 24, dict()),
                           ("four_warps_plain", 128, 24, dict()),
                           ("c1c2_401_coupled", 401, 26, dict(analysis_hours=48, use_coupling=1, use_relaxation=1))):
    arrays, settings, params, rec = synth.make_case(npts, fh, seed=20191203, **kw)
    best = {0: 1e30, 1: 1e30, 2: 1e30}
    regs = {}
    for rep in range(4):
        for mode in (0, 1, 2):     # 2 = latency body + one point per warp (the shipped default for small batches)
            lib.set_option("latency_body", min(mode, 1))
            lib.set_option("spread_small", int(mode == 2))
            work = arrays.copy()
            lib.run_batch(work, settings, params)
            if rep:
                best[mode] = min(best[mode], lib.last_batch_stats()["kernel_ms"])
            regs[mode] = lib.last_launch()["regs_per_thread"]
    res[name] = {"throughput_body_ms": round(best[0], 3), "latency_body_ms": round(best[1], 3),
                 "latency_body_one_point_per_warp_ms": round(best[2], 3),
                 "ratio": round(best[1] / best[0], 4), "ratio_spread": round(best[2] / best[0], 4), "regs": [regs[0], regs[1]]}
lib.set_option("latency_body", -1)
lib.set_option("spread_small", 1)
print(json.dumps(res))
