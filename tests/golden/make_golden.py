"""Generates tests/golden/oracle_v1.npz.

The reference has no golden vectors and cannot be built here, so these are
REGRESSION pins of the CPU restatement (oracle/), not outputs of the Fortran library: they freeze
the oracle's behaviour at the commit that introduced them so that later edits to the oracle or to
the CUDA path cannot drift unnoticed.  Inputs are stored as coarse records + statics; the test
rebuilds the per-step arrays deterministically (no random numbers).

Run from the repository root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import pyoracle  # noqa: E402
from roadsurf_b200 import synth  # noqa: E402

CASES = {
    # name: (npoints, hours, seed, analysis_hours, use_coupling, use_relaxation)
    "plain": (6, 12, 101, 0, 0, 0),
    "coupled": (6, 6, 202, 6, 1, 1),
}
STRIDE = 20


def pack_records(rec):
    d = {v: getattr(rec, v) for v in synth.RECORD_VARS}
    d.update(lat=rec.lat, lon=rec.lon, sky_view=rec.sky_view, horizons=rec.horizons,
             record_step=rec.record_step)
    if rec.obs_bias is not None:
        d["obs_bias"] = rec.obs_bias
    return d


def main():
    out = {}
    for name, (npts, hours, seed, ana, cpl, rel) in CASES.items():
        arrays, settings, params, rec = synth.make_case(npts, hours, seed, analysis_hours=ana, use_coupling=cpl,
                                                        use_relaxation=rel, sky_view_fraction=0.5)
        status, steps = pyoracle.run_batch(arrays, settings, params, nthreads=1)
        for k, v in pack_records(rec).items():
            out[f"{name}/rec/{k}"] = v
        out[f"{name}/meta"] = np.array([npts, hours, ana, cpl, rel, STRIDE, steps], dtype=np.int64)
        out[f"{name}/status"] = status
        for k, v in arrays.out.items():
            out[f"{name}/out/{k}"] = v[:, ::STRIDE].copy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
