"""Device-side generator of the synthetic coarse-forcing workload (BASELINE config 4 / 5 shape).

Same distributions as synth.draw_records but drawn with torch directly in
the device structure-of-arrays layout [record, variable, point], so that a 10^6-point shard is
generated in a fraction of a second.  The random stream differs from the numpy generator's; the
CPU-baseline sample is cut from these very tensors, so both arms see identical inputs."""
import datetime as _dt
import math

import numpy as np

from . import lib as _lib
from . import synth

SIGMA = 5.67e-8


def fill_device_batch(db, seed, start=None, dt_secs=30.0, record_secs=3600, sky_view_fraction=0.3,
                      phase_missing_fraction=0.5):
    """Fill a coarse DeviceBatch (forcing, record_step, time_fields, local, horizons) in place."""
    import torch
    start = start or synth.FORECAST_START
    dev = db.forcing.device
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    P, nrec = db.ld, db.n_records
    f64 = dict(dtype=torch.float64, device=dev)

    def randn(*shape):
        return torch.randn(*shape, generator=g, **f64)

    def rand(*shape):
        return torch.rand(*shape, generator=g, **f64)

    per = int(round(record_secs / dt_secs))
    rs = per * torch.arange(nrec, dtype=torch.int32, device=dev)
    db.record_step.copy_(rs)
    db.time_fields.copy_(torch.from_numpy(synth.time_axis(start, db.sim_len, dt_secs)))

    lat = 59.8 + (69.9 - 59.8) * rand(P)
    lon = 20.0 + (31.0 - 20.0) * rand(P)
    obstructed = rand(P) < sky_view_fraction
    svf = torch.where(obstructed, 0.4 + 0.6 * rand(P), torch.ones(P, **f64))
    L = db.local
    L.zero_()
    L[_lib.L_TAIR_RELAX:_lib.L_RH_RELAX + 1] = -9999.0
    L[_lib.L_COUPLING_TSURF] = -9999.0
    L[_lib.L_COUPLING_INDEX] = -9999.0
    L[_lib.L_LAT], L[_lib.L_LON], L[_lib.L_SKY_VIEW] = lat, lon, svf
    L[_lib.L_ACTIVE, :db.npoints] = 1.0
    if db.horizons is not None:
        ang = torch.deg2rad(torch.arange(360, **f64))[:, None]
        prof = torch.zeros(360, P, **f64)
        for j in range(6):
            prof += randn(P)[None, :] * torch.cos((j + 1) * ang + j)
        prof /= math.sqrt(6.0)
        db.horizons.copy_(torch.where(obstructed[None, :], 8.0 * prof.abs(), torch.zeros_like(prof)))
        del prof

    t0 = -1.0 + 3.0 * randn(P)
    amp = 1.0 + 4.0 * rand(P)
    phase_missing = rand(P) < phase_missing_fraction
    x = randn(P)
    F = db.forcing
    for j in range(nrec):
        when = start + _dt.timedelta(seconds=float(per * j) * dt_secs)
        hour = when.hour + when.minute / 60.0
        x = 0.8 * x + 0.6 * randn(P)
        tair = t0 + amp * math.sin(2.0 * math.pi * (hour - 15.0) / 24.0) + x
        rh = torch.clamp(88.0 + 8.0 * randn(P), 35.0, 100.0)
        doy = when.timetuple().tm_yday
        decl = math.radians(-23.44) * math.cos(2.0 * math.pi * (doy + 10) / 365.0)
        ha = torch.deg2rad(15.0 * (hour - 12.0) + lon)
        latr = torch.deg2rad(lat)
        sin_el = torch.sin(latr) * math.sin(decl) + torch.cos(latr) * math.cos(decl) * torch.cos(ha)
        cloud = 0.2 + 0.8 * rand(P)
        sw = torch.clamp(1361.0 * 0.7 * sin_el, min=0.0) * cloud
        lw = torch.clamp(0.80 * SIGMA * (tair + 273.15) ** 4 + 15.0 * randn(P), 100.0, 450.0)
        alpha = torch.where(tair >= 0.0, 17.269, 21.875)
        beta = torch.where(tair >= 0.0, 237.3, 265.5)
        esat = 0.61078 * torch.exp(alpha * tair / (tair + beta))
        xx = torch.log(0.01 * rh * esat / 0.61078)
        wet = rand(P) < 0.15
        amount = -0.8 * torch.log1p(-rand(P))
        given = torch.where(tair < -0.5, 3.0, torch.where(tair < 1.0, 2.0, 1.0))
        F[j, _lib.F_NAMES.index("tair")] = tair
        F[j, _lib.F_NAMES.index("tdew")] = beta * xx / (alpha - xx)
        F[j, _lib.F_NAMES.index("VZ")] = torch.clamp(torch.exp(math.log(3.0) + 0.6 * randn(P)), 0.0, 25.0)
        F[j, _lib.F_NAMES.index("Rhz")] = rh
        F[j, _lib.F_NAMES.index("prec")] = torch.where(wet, amount, torch.zeros_like(amount))
        F[j, _lib.F_NAMES.index("SW")] = sw
        F[j, _lib.F_NAMES.index("LW")] = lw
        F[j, _lib.F_NAMES.index("SW_dir")] = 0.6 * sw
        F[j, _lib.F_NAMES.index("LW_net")] = lw - 0.95 * SIGMA * (tair + 272.15) ** 4
        F[j, _lib.F_NAMES.index("TSurfObs")] = -9999.9
        F[j, _lib.F_NAMES.index("PrecPhase")] = torch.where(phase_missing, torch.full_like(given, -9999.0), given)
        if db.nvar > _lib.F_NVAR:
            F[j, _lib.F_NVAR] = -9999.9
    return db


def fill_device_batch_chunked(db, seed, start=None, chunk=1_250_000, **kw):
    """fill_device_batch for batches of any size with bounded temporaries: the batch is drawn in pieces of
    `chunk` points (piece k with seed + 1000 k, so the first piece of a large batch equals the batch of
    `chunk` points drawn with `seed`)."""
    if db.ld <= chunk:
        return fill_device_batch(db, seed, start, **kw)
    for k, p0 in enumerate(range(0, db.npoints, chunk)):
        n = min(chunk, db.npoints - p0)
        tmp = _lib.DeviceBatch(n, db.sim_len, nlayers=db.nlayers, n_records=db.n_records, nvar=db.nvar, coarse=True,
                               horizons=db.horizons is not None, out_stride=db.sim_len)
        fill_device_batch(tmp, seed + 1000 * k, start, **kw)
        db.forcing[:, :, p0:p0 + n] = tmp.forcing[:, :, :n]
        db.local[:, p0:p0 + n] = tmp.local[:, :n]
        if db.horizons is not None:
            db.horizons[:, p0:p0 + n] = tmp.horizons[:, :n]
        db.record_step.copy_(tmp.record_step)
        db.time_fields.copy_(tmp.time_fields)
        del tmp
    db.local[_lib.L_ACTIVE, db.npoints:] = 0.0
    db.local[_lib.L_SKY_VIEW, db.npoints:] = 1.0
    return db


def records_sample(db, count):
    """The first `count` points of a coarse DeviceBatch as a synth.Records (numpy), for the CPU arm."""
    count = min(int(count), db.npoints)
    F = db.forcing[:, :, :count].cpu().numpy()
    rec = synth.Records(count, db.n_records)
    for v, name in enumerate(synth.RECORD_VARS):
        setattr(rec, name, np.ascontiguousarray(F[:, v, :].T))
    L = db.local[:, :count].cpu().numpy()
    rec.lat, rec.lon, rec.sky_view = L[_lib.L_LAT].copy(), L[_lib.L_LON].copy(), L[_lib.L_SKY_VIEW].copy()
    if db.horizons is not None:
        rec.horizons = np.ascontiguousarray(db.horizons[:, :count].cpu().numpy().T)
    rec.record_step = db.record_step.cpu().numpy().astype(np.int32)
    return rec
