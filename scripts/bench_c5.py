"""One GPU's share of BASELINE config 5: 51 members x 1e6 points / 8 GPUs = 6.375e6 point-runs, 48 h
(SimLen 5761), hourly forcing records interpolated on the device, hourly output.  The ensemble is the
flattened (member, point) index: every member is an independent point-run with its own forcing."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roadsurf_b200 import abi, lib, synth_torch
P = int(os.environ.get("RS_POINTS", 6_375_000)); hours = 48; sim_len = 1 + hours * 120
lib.set_model(abi.default_settings(sim_len), abi.default_parameters(30.0))
db = lib.DeviceBatch(P, sim_len, n_records=hours + 2, coarse=True, horizons=True, out_stride=120)
synth_torch.fill_device_batch(db, seed=51)
db.run(); torch.cuda.synchronize()
ts = []
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); db.run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
cnt = db.counters.cpu().numpy()
print(json.dumps({"config": "c5 shard", "point_runs": P, "sim_len": sim_len, "ms": round(min(ts), 1),
                  "point_steps_per_s": P * sim_len / min(ts) * 1e3, "failed": int(cnt[3]),
                  "hbm_gb": round(torch.cuda.max_memory_allocated() / 1e9, 1), "launch": lib.last_launch()}))
