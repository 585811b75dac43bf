"""Loads tests/golden/oracle_v1.npz back into runnable cases (see tests/golden/make_golden.py)."""
import os

import numpy as np

from roadsurf_b200 import synth

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_v1.npz")
CASE_NAMES = ("plain", "coupled")


def load_case(name):
    z = np.load(PATH)
    npts, hours, ana, cpl, rel, stride, steps = [int(v) for v in z[f"{name}/meta"]]
    nrec = z[f"{name}/rec/tair"].shape[1]
    rec = synth.Records(npts, nrec)
    for v in synth.RECORD_VARS:
        setattr(rec, v, z[f"{name}/rec/{v}"].copy())
    rec.lat, rec.lon, rec.sky_view = z[f"{name}/rec/lat"], z[f"{name}/rec/lon"], z[f"{name}/rec/sky_view"]
    rec.horizons = z[f"{name}/rec/horizons"]
    rec.record_step = z[f"{name}/rec/record_step"]
    key = f"{name}/rec/obs_bias"
    rec.obs_bias = z[key] if key in z.files else None
    arrays, settings, params = synth.case_from_records(rec, hours, ana, cpl, rel)
    golden = {k.split("/")[-1]: z[k] for k in z.files if k.startswith(f"{name}/out/")}
    return arrays, settings, params, golden, z[f"{name}/status"], stride, steps
