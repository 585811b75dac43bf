import sys, time, json
import numpy as np
sys.path.insert(0, '.')
from roadsurf_b200 import synth, abi, lib
from oracle import pyoracle
sys.path.insert(0, 'tests')
from parity import compare

def run(npts, hours, ana, cpl, rel, seed, sky=0.3):
    pa, st, prm, rec = synth.make_case(npts, hours, seed=seed, analysis_hours=ana, use_coupling=cpl, use_relaxation=rel, sky_view_fraction=sky)
    pb = pa.copy()
    t = time.time(); so, steps = pyoracle.run_batch(pa, st, prm, nthreads=8); to = time.time() - t
    t = time.time(); sg = lib.run_batch(pb, st, prm, ngpus=1); tg = time.time() - t
    r = compare(pb.out, pa.out)
    r.update(oracle_s=round(to, 3), gpu_s=round(tg, 3), status_equal=bool((so == sg).all()), steps=steps, stats=lib.last_batch_stats(), launch=lib.last_launch())
    print(json.dumps(r))
    if not (so == sg).all():
        w = np.where(so != sg)[0][:10]; print("status diff", w, so[w], sg[w])
    return pa, pb, r

print(lib.load().roadsurf_version())
run(64, 6, 0, 0, 0, 1, sky=0.0)
run(64, 24, 0, 0, 0, 2)
run(256, 24, 6, 1, 1, 3)
run(1000, 48, 6, 1, 1, 4)
print("fp64 TF", lib.measure_fp64_tflops(20000))
