import sys, time, json
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from roadsurf_b200 import synth, lib, abi
from oracle import pyoracle
from parity import compare

def timed(fn, reps=3):
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), ts

# ---- coarse parity: device interpolation vs host-interpolated oracle
npts, hours = 512, 24
pa, st, prm, rec = synth.make_case(npts, hours, seed=11)
so, steps = pyoracle.run_batch(pa, st, prm, nthreads=8)
lib.set_model(st, prm)
db = lib.DeviceBatch(npts, pa.sim_len, n_records=rec.nrec, coarse=True, horizons=True, out_stride=1)
db.load_records(rec); db.time_fields.copy_(torch.from_numpy(pa.time)); db.load_local(pa.local, pa.local_horizons)
db.run(); torch.cuda.synchronize()
r = compare(db.outputs(), pa.out); r["status_equal"] = bool((db.status.cpu().numpy()[:npts] == so).all())
print("coarse parity", json.dumps(r))
# strided
db2 = lib.DeviceBatch(npts, pa.sim_len, n_records=rec.nrec, coarse=True, horizons=True, out_stride=120)
db2.load_records(rec); db2.time_fields.copy_(torch.from_numpy(pa.time)); db2.load_local(pa.local, pa.local_horizons)
db2.run(); torch.cuda.synchronize()
o2 = db2.outputs(); o1 = db.outputs()
print("strided equal", all(np.array_equal(o2[k], o1[k][:, ::120]) for k in o1), o2["TsurfOut"].shape)

# ---- perf, coarse mode, c4-like
for P in (56832, 113664, 1250016):
    reps = (P + npts - 1) // npts
    dbp = lib.DeviceBatch(P, pa.sim_len, n_records=rec.nrec, coarse=True, horizons=True, out_stride=120)
    dbp.forcing.copy_(db.forcing[:, :, :npts].repeat(1, 1, reps)[:, :, :dbp.ld])
    dbp.record_step.copy_(db.record_step); dbp.time_fields.copy_(db.time_fields)
    dbp.local.copy_(db.local[:, :npts].repeat(1, reps)[:, :dbp.ld]); dbp.local[lib.L_ACTIVE, P:] = 0
    dbp.horizons.copy_(db.horizons[:, :npts].repeat(1, reps)[:, :dbp.ld])
    best, ts = timed(dbp.run)
    cnt = dbp.counters.cpu().numpy()
    print("coarse P", P, "ms", round(best, 2), [round(t, 1) for t in ts], "pt-steps/s %.3e" % (P * pa.sim_len / best * 1e3), "bl iters/step %.2f" % (cnt[1] / cnt[0]), lib.last_launch())
    del dbp
# ---- perf, full-res mode
P = 113664
reps = (P + npts - 1) // npts
dbf = lib.DeviceBatch(P, pa.sim_len, horizons=True)
small = lib.DeviceBatch(npts, pa.sim_len, horizons=True); small.load_point_arrays(pa)
dbf.forcing.copy_(small.forcing[:, :, :npts].repeat(1, 1, reps)[:, :, :dbf.ld])
dbf.time_fields.copy_(small.time_fields)
dbf.local.copy_(small.local[:, :npts].repeat(1, reps)[:, :dbf.ld]); dbf.local[lib.L_ACTIVE, P:] = 0
dbf.horizons.copy_(small.horizons[:, :npts].repeat(1, reps)[:, :dbf.ld])
best, ts = timed(dbf.run)
print("full P", P, "ms", round(best, 2), [round(t, 1) for t in ts], "pt-steps/s %.3e" % (P * pa.sim_len / best * 1e3), "GB/s %.1f" % (P * pa.sim_len * 136 / best / 1e6))
small.run(); torch.cuda.synchronize()
print("full-res device parity", json.dumps(compare(small.outputs(), pa.out)))
print("fp64 TF", lib.measure_fp64_tflops(20000))
