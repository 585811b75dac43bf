import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from roadsurf_b200 import lib, synth
P = int(os.environ.get("RS_POINTS", 60000))
t0 = time.perf_counter()
arrays, settings, params, rec = synth.make_case(P, 24, seed=77, analysis_hours=6, use_coupling=1, use_relaxation=1)
gen = time.perf_counter() - t0
pb = lib.PreparedBatch(arrays)
pb.run(settings, params)
ts = []
for _ in range(2):
    t0 = time.perf_counter(); pb.run(settings, params); ts.append(time.perf_counter() - t0)
st = lib.last_batch_stats()
print(json.dumps({"lib": os.environ.get("ROADSURF_B200_LIBNAME", "default"), "points": P, "sim_len": arrays.sim_len, "gen_s": round(gen, 1),
                  "e2e_s": round(min(ts), 3), "stats": {k: (round(v, 1) if isinstance(v, float) else v) for k, v in st.items()},
                  "tsum": float(np.nansum(arrays.out["TsurfOut"][:, ::97]))}))
