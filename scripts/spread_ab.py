"""Same-box A/B of option "spread_small" (few points per warp for batches of at most 4 x 16 points per SM... see
rs_launch_run) on plain forecasts of growing size: kernel time with 32 points per warp against the spread mapping."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roadsurf_b200 import lib, synth
res = {}
for npts in (32, 401, 592, 1000, 2000, 4000, 9000, 12000):
    arrays, settings, params, rec = synth.make_case(npts, 6, seed=20191203)
    pb = None
    best = {0: 1e30, 1: 1e30}
    grid = {}
    for rep in range(3):
        for mode in (0, 1):
            lib.set_option("spread_small", mode)
            work = arrays.copy()
            lib.run_batch(work, settings, params)
            if rep:
                best[mode] = min(best[mode], lib.last_batch_stats()["kernel_ms"])
            grid[mode] = lib.last_launch()["grid"]
    res[str(npts)] = {"dense_ms": round(best[0], 3), "spread_ms": round(best[1], 3), "ratio": round(best[1] / best[0], 3),
                      "grids": [grid[0], grid[1]], "us_per_step": round(best[1] * 1e3 / arrays.sim_len, 2)}
lib.set_option("spread_small", 1)
print(json.dumps(res))
