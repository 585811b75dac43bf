"""Effect of point ORDER on the c4 shard: points with sky-view radiation (30 %) scattered at random vs\ngathered at the end, so that 70 % of the warps skip the solar geometry altogether."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roadsurf_b200 import abi, lib, synth_torch
P = 1250304; hours = 24; sim_len = 1 + hours * 120
lib.set_model(abi.default_settings(sim_len), abi.default_parameters(30.0))
db = lib.DeviceBatch(P, sim_len, n_records=hours + 2, coarse=True, horizons=True, out_stride=120)
synth_torch.fill_device_batch(db, seed=7)
def t():
    db.run(); torch.cuda.synchronize(); ts=[]
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); db.run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
a = t(); ref = db.out.clone()
sky = (db.local[lib.L_SKY_VIEW] < 1.0)
perm = torch.argsort(sky.to(torch.int8), stable=True)      # open-sky points first, sky-view points last
for name in ("forcing", "local", "horizons"):
    x = getattr(db, name); x.copy_(x.index_select(-1, perm))
b = t()
same = bool(torch.equal(db.out, ref.index_select(-1, perm)))
print(json.dumps({"random_order_ms": round(a, 2), "sorted_by_sky_view_ms": round(b, 2), "gain": a / b, "outputs_equal_after_permutation": same,
                  "sky_fraction": float(sky.float().mean())}))
