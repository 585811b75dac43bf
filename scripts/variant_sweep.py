"""Times library variants (roadsurf_b200/libvar_*.so, built with ROADSURF_B200_LIBNAME / ROADSURF_B200_NVCC_EXTRA)
on the bench workload (config-4 shard, coarse forcing, point order rebuilt per pass) in one process per
variant, and prints one JSON line each with a checksum of the outputs (variants must agree bit for bit).

usage: python scripts/variant_sweep.py [name ...]        (no names: every libvar_*.so in the package)
env:   RS_POINTS (default 1250000), RS_REPS (default 2)
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, os, sys
sys.path.insert(0, %(root)r)
import torch
from roadsurf_b200 import abi, lib, synth_torch
P = int(os.environ.get("RS_POINTS", 1250000)); hours = 24; sim_len = 1 + hours * 120
lib.set_model(abi.default_settings(sim_len), abi.default_parameters(30.0))
db = lib.DeviceBatch(P, sim_len, n_records=hours + 2, coarse=True, horizons=True, out_stride=120)
synth_torch.fill_device_batch(db, seed=20191206)
st = torch.cuda.current_stream()
def step():
    db.build_order(st); db.run(st)
step(); torch.cuda.synchronize()
ts = []
for _ in range(int(os.environ.get("RS_REPS", 2))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st); step(); e1.record(st); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
out = db.out[:, :, :P].contiguous().view(torch.int64)
chk = int(out.sum().item()) & 0xffffffffffff
cnt = db.counters.cpu().numpy()
li = lib.last_launch()
print(json.dumps({"lib": os.environ.get("ROADSURF_B200_LIBNAME"), "ms": round(min(ts), 2), "ms_all": [round(t, 2) for t in ts],
                  "rate": P * sim_len / min(ts) * 1e3, "regs": li["regs_per_thread"], "block": li["block"],
                  "grid": li["grid"], "checksum": chk, "bl_per_step": float(cnt[1]) / max(1.0, float(cnt[0]))}))
"""


def main():
    names = sys.argv[1:]
    if not names:
        names = sorted(os.path.basename(p)[len("libvar_"):-3] for p in glob.glob(os.path.join(ROOT, "roadsurf_b200", "libvar_*.so")))
    for n in names:
        env = dict(os.environ, ROADSURF_B200_LIBNAME=f"libvar_{n}.so")
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env, capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else json.dumps({"lib": n, "error": r.stderr[-400:]})
        print(line, flush=True)


if __name__ == "__main__":
    main()
