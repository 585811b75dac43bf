"""Parity report: the CUDA path (through roadsurf_run_batch) against the CPU restatement on the
synthetic configurations, with the distribution of the differences (not only the maxima).
usage: python scripts/parity_report.py > profiles/rNN_parity_report.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from roadsurf_b200 import lib, synth
from oracle import pyoracle
from parity import compare, T_TOL, S_TOL, STORAGES

CASES = [
    ("c4-like: 24 h forecast, 30 % sky view, no coupling", dict(npoints=4096, hours=24, seed=20191206)),
    ("c3-like: 6 h analysis + 48 h forecast, coupling + relaxation", dict(npoints=4096, hours=48, seed=20191205,
                                                                       analysis_hours=6, use_coupling=1, use_relaxation=1)),
    ("c1/c2-like: 48 h analysis + 26 h forecast (example1's span), coupling + relaxation",
     dict(npoints=401, hours=26, seed=20191203, analysis_hours=48, use_coupling=1, use_relaxation=1)),
]
print("GPU (roadsurf_run_batch, sm_100a) vs CPU restatement (oracle/, strict build); tolerances %.0e K / %.0e mm" % (T_TOL, S_TOL))
print("flip = a point with any value beyond tolerance (threshold-induced state flip, DESIGN.md section 4)\n")
for title, kw in CASES:
    arrays, settings, params, _ = synth.make_case(**kw)
    ref = arrays.copy()
    st_gpu = lib.run_batch(arrays, settings, params)
    st_cpu, _ = pyoracle.run_batch(ref, settings, params, nthreads=os.cpu_count())
    r = compare(arrays.out, ref.out)
    dT = np.abs(arrays.out["TsurfOut"] - ref.out["TsurfOut"])
    dS = np.max([np.abs(arrays.out[n] - ref.out[n]) for n in STORAGES], axis=0)
    bad = (dT.max(axis=1) > T_TOL) | (dS.max(axis=1) > S_TOL)
    good = ~bad
    print(title)
    print("  points %d x steps %d; status words equal: %s; coupled points: %d; coupling failed: %d"
          % (arrays.npoints, arrays.sim_len, bool((st_gpu == st_cpu).all()), int(((st_gpu & 8) > 0).sum()),
             int(((st_gpu & 16) > 0).sum())))
    print("  flips: %d of %d points (%.2f %%)" % (bad.sum(), bad.size, 100 * bad.mean()))
    q = [50, 90, 99, 99.9, 100]
    pT = np.percentile(dT[good].ravel(), q); pS = np.percentile(dS[good].ravel(), q)
    print("  matching points, |dT| [K]  percentiles %s: %s" % (q, " ".join("%.2e" % v for v in pT)))
    print("  matching points, |dS| [mm] percentiles %s: %s" % (q, " ".join("%.2e" % v for v in pS)))
    print("  bit-identical values among matching points: %.2f %% of Tsurf, %.2f %% of storages"
          % (100 * (dT[good] == 0).mean(), 100 * (dS[good] == 0).mean()))
    if bad.any():
        print("  flipped points: max |dT| %.2e K, max |dS| %.2e mm; first diverging step: median %d, min %d"
              % (dT[bad].max(), dS[bad].max(),
                 int(np.median([np.argmax(np.maximum(dT[p] / T_TOL, dS[p] / S_TOL) > 1) for p in np.where(bad)[0]])),
                 int(np.min([np.argmax(np.maximum(dT[p] / T_TOL, dS[p] / S_TOL) > 1) for p in np.where(bad)[0]]))))
    print()
