"""Short single-GPU run of the step kernel for ncu (one 24 h pass over 113 664 points, coarse forcing)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from roadsurf_b200 import abi, lib, synth_torch  # noqa: E402

P = int(os.environ.get("RS_PROFILE_POINTS", 113664))
hours = int(os.environ.get("RS_PROFILE_HOURS", 24))
sim_len = 1 + hours * 120
lib.set_model(abi.default_settings(sim_len), abi.default_parameters(30.0))
db = lib.DeviceBatch(P, sim_len, n_records=hours + 2, coarse=True, horizons=True, out_stride=120)
synth_torch.fill_device_batch(db, seed=7)
for _ in range(int(os.environ.get("RS_PROFILE_REPS", 2))):
    db.run()
torch.cuda.synchronize()
cnt = db.counters.cpu().numpy()
print("ok", P, sim_len, "failed", int(cnt[3]), "bl/step", cnt[1] / cnt[0], lib.last_launch())
