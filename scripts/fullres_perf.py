"""Throughput of the full-resolution (staged forcing) kernel on a large grid."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roadsurf_b200 import abi, lib, synth
P = int(os.environ.get("RS_POINTS", 151552)); hours = 24
arrays, settings, params, rec = synth.make_case(1024, hours, seed=5)
lib.set_model(settings, params)
small = lib.DeviceBatch(1024, arrays.sim_len, horizons=True); small.load_point_arrays(arrays)
db = lib.DeviceBatch(P, arrays.sim_len, horizons=True)
idx = torch.arange(db.ld, device="cuda") % 1024
db.forcing.copy_(small.forcing[:, :, idx]); db.time_fields.copy_(small.time_fields)
db.local.copy_(small.local[:, idx]); db.local[lib.L_ACTIVE, P:] = 0; db.horizons.copy_(small.horizons[:, idx])
db.run(); torch.cuda.synchronize()
ts = []
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); db.run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
small.run(); torch.cuda.synchronize()
same = bool(torch.equal(db.out[:, :, :P], small.out[:, :, idx[:P]]))
print(json.dumps({"P": P, "ms": round(min(ts), 2), "rate": P * arrays.sim_len / min(ts) * 1e3, "GBps": P * arrays.sim_len * 136 / min(ts) / 1e6,
                  "replicas_equal": same, "launch": lib.last_launch()}))
