"""Small coupled runs through every launch mode (split at the window end, compacted passes, coarse
records, extended outputs, batched entry): a quick multi-mode check (compute-sanitizer is closed on this pool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from roadsurf_b200 import lib, synth
npts = 203
arrays, settings, params, rec = synth.make_case(npts, 2, seed=51, analysis_hours=2, use_coupling=1, use_relaxation=1,
                                                 obs_bias=False, settings_kw=dict(coupling_minutes=30))
arrays.local[3].couplingTsurf = -9999.9; arrays.local[3].couplingIndexI = -9999
lib.set_model(settings, params)
db = lib.DeviceBatch(npts, arrays.sim_len, horizons=True, coupling=True, state=True)
db.load_point_arrays(arrays); db.run(); torch.cuda.synchronize()
ref = lib.DeviceBatch(npts, arrays.sim_len, horizons=True, coupling=True)
ref.load_point_arrays(arrays); ref.run(); torch.cuda.synchronize()
print("full-res equal", torch.equal(db.out, ref.out), "window", db.coupling_window_end)
c = lib.DeviceBatch(npts, arrays.sim_len, n_records=rec.nrec, coarse=True, horizons=True, coupling=True, state=True, out_stride=60,
                    out_start=7, extended_outputs=True)
c.load_records(rec); c.time_fields.copy_(torch.from_numpy(arrays.time)); c.load_local(arrays.local, arrays.local_horizons)
c.run(); torch.cuda.synchronize()
print("coarse ok", int(c.counters[0]))
a2 = arrays.copy(); lib.run_batch(a2, settings, params)
print("batch equal", np.array_equal(a2.out["TsurfOut"], db.outputs()["TsurfOut"]))
