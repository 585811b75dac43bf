"""Example-1-style driver on top of the C ABI: config JSON + forecast / observation JSON + sky-view
files in, forecast JSON out.

This mirrors the flow of the reference's examples/example1 (roadrunner.cpp, DataHandler, JsonSource,
SkyView, InputSettings) so that BASELINE configs c1 / c2 can be run from files in the reference's own
schema.  It is host-side glue for the examples, not the product: the model itself runs only through
libroadsurf_b200.so (`runner` defaults to lib.run_batch; tests inject a CPU checker).

Formats (all from the reference):
  config   JSON with // comments: time.{now,analysis,forecast,coupling_minutes},
           model.{use_coupling,use_relaxation,DTSecs,tsurfOutputDepth,NLayers,couplingEffectReduction},
           output.{step,filename}, parameters.{sky_view_file,local_horizon_file,<InputParameters>},
           input[] = {name,path,type:"json",source:"forecast"|"observations"}
           (examples/example1/example_config.json, InputSettings.cpp:69-104, InputParameters.cpp:30-109)
  weather  array of stations {statId, lat, lon, time:["%Y-%m-%d %H:%M"], "Temperature 2m", "Humidity",
           "DewPoint", "WindSpeed", "PrecipitationForm", "Precipitation", "RadiationNetSurfaceLW",
           "RadiationLW", "RadiationGlobal", "RadiationDirectSW", "RoadTemperature"}
           (JsonSource.cpp:183-316)
  sky view "id name lat lon sky_view" per line; horizons "id name lat lon h0 ... h359" (SkyView.cpp:14-122)
  output   array of {statId, lat, lon, time:["%Y-%m-%dT%H:%M"], RoadTemperature, Water, Ice, Snow,
           Deposit} every output.step minutes (roadrunner.cpp:285-347)
Times are UTC (the reference uses mktime/localtime: run it with TZ=UTC).
"""
import calendar
import datetime as _dt
import json
import re

import numpy as np

from . import abi

VARIABLES = (("Temperature 2m", "tair"), ("Humidity", "Rhz"), ("DewPoint", "tdew"), ("WindSpeed", "VZ"),
             ("PrecipitationForm", "PrecPhase"), ("Precipitation", "prec"), ("RadiationNetSurfaceLW", "LW_net"),
             ("RadiationLW", "LW"), ("RadiationGlobal", "SW"), ("RadiationDirectSW", "SW_dir"),
             ("RoadTemperature", "TSurfObs"))
FIELDS = tuple(f for _, f in VARIABLES)


def load_json(path):
    """JSON with // and /* */ comments (jsoncpp accepts them, Python's json does not)."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"(?m)(^|[\s,{}\[\]])//.*$", r"\1", text)
    return json.loads(text)


def _epoch(s, fmt):
    return calendar.timegm(_dt.datetime.strptime(s, fmt).timetuple())


class Times:
    """InputSettings.cpp:9-104: forecast (wall clock) time, model start / end, SimLen."""

    def __init__(self, cfg, now=None):
        t = cfg.get("time", {})
        if now is None:
            now = t.get("now")
        if now is None:
            raise ValueError("forecast start time needed: time.now in the config or the `now` argument")
        self.forecast = _epoch(now, "%Y%m%dT%H%M") if isinstance(now, str) else int(now)
        self.start = self.forecast - int(t.get("analysis", 24)) * 3600
        self.end = self.forecast + int(t.get("forecast", 48)) * 3600
        self.coupling_minutes = int(t.get("coupling_minutes", 0))


def parse_config(cfg, now=None):
    """-> (Times, abi.InputSettings, abi.InputParameters, output step minutes)."""
    times = Times(cfg, now)
    model = cfg.get("model", {})
    dt = float(model.get("DTSecs", 30.0))
    sim_len = 1 + int((times.end - times.start) / dt)
    out_step = int(cfg.get("output", {}).get("step", 60))
    settings = abi.default_settings(sim_len, use_coupling=int(model.get("use_coupling", 0)),
                                    use_relaxation=int(model.get("use_relaxation", 0)), dt=dt,
                                    nlayers=int(model.get("NLayers", 15)),
                                    tsurf_output_depth=float(model.get("tsurfOutputDepth", -9999.9)),
                                    coupling_effect_reduction=float(model.get("couplingEffectReduction", 4.0 * 3600)),
                                    output_step=out_step)
    if times.coupling_minutes > 0:
        settings.coupling_minutes = times.coupling_minutes
    overrides = {k: float(v) for k, v in cfg.get("parameters", {}).items() if k in abi.PARAMETER_NAMES}
    params = abi.default_parameters(dt, **overrides)
    return times, settings, params, out_step


def read_sky_view(sky_view_file=None, local_horizon_file=None):
    """SkyView.cpp:14-122 -> {id: (sky_view, horizons[360])}."""
    data = {}
    if sky_view_file:
        for line in open(sky_view_file):
            parts = line.split()
            if len(parts) < 5:
                continue
            sv = float(parts[4])
            data[int(parts[0])] = [sv if 0.0 <= sv <= 1.0 else 1.0, np.zeros(360)]
    if local_horizon_file:
        for line in open(local_horizon_file):
            parts = line.split()
            if len(parts) < 5:
                continue
            hor = np.array([float(v) for v in parts[4:]])
            entry = data.setdefault(int(parts[0]), [1.0, np.zeros(360)])
            entry[1] = hor
    return {k: (v[0], v[1]) for k, v in data.items()}


def tdew_or_rh(t, tdew=None, rh=None):
    """MeteorologyTools.cpp:12-51 (vectorised): RH from dew point, or dew point from RH."""
    alpha = np.where(t >= 0.0, 17.269, 21.875)
    beta = np.where(t >= 0.0, 237.3, 265.5)
    afact = 0.61078
    esat = afact * np.exp(alpha * t / (t + beta))
    if tdew is not None:
        esat_td = afact * np.exp(alpha * tdew / (tdew + beta))
        return np.minimum((esat_td / esat) * 100.0, 100.0)
    xx = np.log(0.01 * rh * esat / afact)
    return beta * xx / (alpha - xx)


def interpolate(rawtime, raw, simtime, miss=-100.0, next_record=False):
    """JsonSource.cpp:49-176 for one variable (vectorised).  rawtime / simtime: increasing integer
    seconds.  Simulation times before the first or at/after the last raw time are left missing
    (the reference's while loop stops at the last record).  next_record: precipitation phase rule."""
    rawtime = np.asarray(rawtime, dtype=np.int64)
    simtime = np.asarray(simtime, dtype=np.int64)
    raw = np.asarray(raw, dtype=np.float64)
    out = np.full(len(simtime), -9999.0 if next_record else -9999.9)
    if len(rawtime) < 2:
        return out
    k = np.searchsorted(rawtime, simtime, side="right") - 1
    ok = (k >= 0) & (k + 1 < len(rawtime))
    k = np.clip(k, 0, len(rawtime) - 2)
    a, b = raw[k], raw[k + 1]
    exact = simtime == rawtime[k]
    if next_record:
        val = np.where(exact, a, b)
        return np.where(ok & (val > -100.0), val, out)
    dt = (simtime - rawtime[k]).astype(np.float64)
    span = (rawtime[k + 1] - rawtime[k]).astype(np.float64)
    between = a + dt * (b - a) / span
    val = np.where(exact, np.where(a > miss, a, out), np.where((a > miss) & (b > miss), between, out))
    return np.where(ok, val, out)


class JsonSource:
    """One input source: per station the raw series interpolated to the model steps."""

    def __init__(self, path, simtime, is_observation=False):
        self.is_observation = is_observation
        self.stations = {}
        for st in load_json(path):
            sid = int(st["statId"])
            times = st.get("time", [])
            n = len(times)
            rec = {"lat": float(st["lat"]), "lon": float(st["lon"]), "fields": None}
            if n:
                rawtime = [_epoch(t, "%Y-%m-%d %H:%M") for t in times]
                raw = {f: np.full(n, -9999.9) for f in FIELDS}
                raw["PrecPhase"] = np.full(n, -9999.0)
                for name, f in VARIABLES:
                    if st.get(name) is not None:
                        raw[f] = np.array([float(v) for v in st[name]], dtype=np.float64)
                t, td, rh = raw["tair"], raw["tdew"], raw["Rhz"]
                need_td = (td < -100) & (rh > -100) & (t > -100)
                need_rh = (rh < -100) & (td > -100) & (t > -100)
                with np.errstate(all="ignore"):
                    raw["tdew"] = np.where(need_td, tdew_or_rh(t, rh=rh), td)
                    raw["Rhz"] = np.where(need_rh, tdew_or_rh(t, tdew=td), rh)
                rec["fields"] = {f: interpolate(rawtime, raw[f], simtime, -1000.0 if f == "LW_net" else -100.0,
                                                next_record=(f == "PrecPhase")) for f in FIELDS}
            self.stations[sid] = rec

    def latest_obs_index(self, sid):
        """JsonSource.cpp:397-414: 1-based count up to the last step with an air temperature."""
        rec = self.stations.get(sid)
        if rec is None or rec["fields"] is None:
            return -9999
        idx = np.nonzero(rec["fields"]["tair"] > -100)[0]
        return int(idx[-1]) + 1 if len(idx) else -9999


def build_inputs(cfg, now=None, base_dir="."):
    """DataHandler + read_input: -> (arrays, settings, params, station ids, times, output step).
    Later sources overwrite earlier ones where they have data (DataHandler.cpp:75-82)."""
    import os
    times, settings, params, out_step = parse_config(cfg, now)
    sim_len = settings.SimLen
    simtime = times.start + (np.arange(sim_len) * settings.DTSecs).astype(np.int64)
    sources = [JsonSource(os.path.join(base_dir, s["path"]), simtime, s.get("source") == "observations")
               for s in cfg.get("input", []) if s.get("type", "json") == "json"]
    ids = []
    for s in sources:
        for sid in s.stations:
            if sid not in ids:
                ids.append(sid)
    p = cfg.get("parameters", {})
    sky = read_sky_view(os.path.join(base_dir, p["sky_view_file"]) if p.get("sky_view_file") else None,
                        os.path.join(base_dir, p["local_horizon_file"]) if p.get("local_horizon_file") else None)
    arrays = abi.PointArrays(len(ids), sim_len)
    when = [_dt.datetime.fromtimestamp(int(t), _dt.timezone.utc) for t in simtime]
    arrays.time[:] = np.array([[w.year, w.month, w.day, w.hour, w.minute, w.second] for w in when], dtype=np.int32).T
    latest = np.full(len(ids), -9999, dtype=np.int32)
    for q, sid in enumerate(ids):
        lp = arrays.local[q]
        lp.tair_relax = lp.VZ_relax = lp.RH_relax = lp.couplingTsurf = -9999.0
        lp.couplingIndexI, lp.sky_view, lp.InitLenI = -9999, 1.0, 0
        for s in sources:
            rec = s.stations.get(sid)
            if rec is None:
                continue
            lp.lat, lp.lon = rec["lat"], rec["lon"]
            if rec["fields"] is None:
                continue
            for f in FIELDS:
                src = rec["fields"][f]
                dst = getattr(arrays, f)[q]
                m = src > (-1000.0 if f == "LW_net" else -100.0)
                dst[m] = src[m] if f != "PrecPhase" else src[m].astype(np.int32)
            if s.is_observation:
                latest[q] = max(latest[q], s.latest_obs_index(sid))
        if sid in sky:
            lp.sky_view = sky[sid][0]
            arrays.local_horizons[q, :len(sky[sid][1])] = sky[sid][1][:360]
    forecast_step = int((times.forecast - times.start) / settings.DTSecs)
    return arrays, settings, params, ids, times, out_step, forecast_step, latest


def save_output(arrays, settings, ids, times, out_step):
    """roadrunner.cpp:285-327: every output step, keys as the reference writes them."""
    step = int(out_step * 60 / settings.DTSecs)
    idx = np.arange(0, arrays.sim_len, step)
    stamps = [_dt.datetime.fromtimestamp(int(times.start + i * settings.DTSecs), _dt.timezone.utc).strftime("%Y-%m-%dT%H:%M")
              for i in idx]
    out = []
    for q, sid in enumerate(ids):
        lp = arrays.local[q]
        o = arrays.out
        out.append({"statId": int(sid), "lat": lp.lat, "lon": lp.lon, "time": stamps,
                    "RoadTemperature": [float(v) for v in o["TsurfOut"][q, idx]],
                    "Water": [float(v) for v in o["WaterOut"][q, idx]],
                    "Ice": [float(v) for v in o["IceOut"][q, idx]],
                    "Snow": [float(v) for v in o["SnowOut"][q, idx]],
                    "Deposit": [float(v) for v in o["DepositOut"][q, idx]]})
    return out


def run(config_path, now=None, runner=None, ngpus=1, write=True):
    """main() of the example: returns (forecast list, arrays, status).  `runner(arrays, settings,
    params)` runs the model; the default is the CUDA library through roadsurf_run_batch."""
    import os
    from . import lib
    cfg = load_json(config_path)
    base = os.path.dirname(os.path.abspath(config_path))
    arrays, settings, params, ids, times, out_step, forecast_step, latest = build_inputs(cfg, now, base)
    ok = lib.read_input_derive(arrays, settings, forecast_step, latest_obs_index=latest)
    if runner is None:
        def runner(a, s, p):
            return lib.run_batch(a, s, p, ngpus=ngpus)
    good = np.nonzero(ok)[0]
    status = np.full(len(ids), lib.ST_NOT_RUN, dtype=np.int32)
    if len(good) == len(ids):
        status[:] = runner(arrays, settings, params)
    elif len(good):
        sub = abi.PointArrays(len(good), arrays.sim_len)
        for f in abi.INPUT_DOUBLE_FIELDS:
            getattr(sub, f[2:])[:] = getattr(arrays, f[2:])[good]
        sub.PrecPhase[:] = arrays.PrecPhase[good]
        sub.local_horizons[:] = arrays.local_horizons[good]
        sub.time[:] = arrays.time
        for q, p_ in enumerate(good):
            sub.local[q] = arrays.local[int(p_)]
        status[good] = runner(sub, settings, params)
        for k in arrays.out:
            arrays.out[k][good] = sub.out[k]
    forecast = save_output(arrays, settings, ids, times, out_step)
    name = cfg.get("output", {}).get("filename")
    if write and name:
        with open(os.path.join(base, name), "w") as f:
            json.dump(forecast, f, indent=3)
    return forecast, arrays, status


def write_synthetic_inputs(directory, nstations=5, seed=1, analysis=6, forecast=6, now="20191202T0000",
                           use_coupling=1, use_relaxation=1, sky_view_fraction=0.4):
    """Stand-ins for the reference's missing example_forecast.json / example_observations.json
    (listed in its .MISSING_LARGE_BLOBS): same schema, synthetic values (synth.draw_records)."""
    import os
    from . import synth
    t_now = _dt.datetime.strptime(now, "%Y%m%dT%H%M")
    start = t_now - _dt.timedelta(hours=analysis)
    nrec = analysis + forecast + 2
    rec = synth.draw_records(nstations, nrec, seed, start, sky_view_fraction=sky_view_fraction)
    ids = [100100 + 2 * k for k in range(nstations)]
    stamps = [(start + _dt.timedelta(hours=j)).strftime("%Y-%m-%d %H:%M") for j in range(nrec)]
    fc, ob = [], []
    rng = np.random.default_rng(seed + 5)
    for q, sid in enumerate(ids):
        base = {"statId": sid, "lat": float(rec.lat[q]), "lon": float(rec.lon[q])}
        fc.append(dict(base, time=stamps, **{
            "Temperature 2m": rec.tair[q].tolist(), "Humidity": rec.Rhz[q].tolist(), "WindSpeed": rec.VZ[q].tolist(),
            "PrecipitationForm": rec.PrecPhase[q].tolist(), "Precipitation": rec.prec[q].tolist(),
            "RadiationNetSurfaceLW": rec.LW_net[q].tolist(), "RadiationLW": rec.LW[q].tolist(),
            "RadiationGlobal": rec.SW[q].tolist(), "RadiationDirectSW": rec.SW_dir[q].tolist()}))
        n_obs = analysis + 1   # observations up to the forecast start, every hour
        bias = rng.normal(0.0, 1.0)
        ob.append(dict(base, time=stamps[:n_obs], **{
            "Temperature 2m": (rec.tair[q, :n_obs] + bias).tolist(),
            "Humidity": np.clip(rec.Rhz[q, :n_obs] + 3.0, 35, 100).tolist(),
            "WindSpeed": rec.VZ[q, :n_obs].tolist(), "Precipitation": rec.prec[q, :n_obs].tolist(),
            "RoadTemperature": rec.TSurfObs[q, :n_obs].tolist()}))
    os.makedirs(directory, exist_ok=True)
    json.dump(fc, open(os.path.join(directory, "forecast.json"), "w"))
    json.dump(ob, open(os.path.join(directory, "observations.json"), "w"))
    with open(os.path.join(directory, "skyview.txt"), "w") as f:
        for q, sid in enumerate(ids):
            f.write(f"{sid} point{sid}  {rec.lat[q]:.7f}   {rec.lon[q]:.7f} {rec.sky_view[q]:.3f}\n")
    with open(os.path.join(directory, "horizons.txt"), "w") as f:
        for q, sid in enumerate(ids):
            f.write(f"{sid} point{sid}   {rec.lat[q]:.7f}   {rec.lon[q]:.7f} " +
                    " ".join(f"{h:.1f}" for h in rec.horizons[q]) + "\n")
    cfg = f"""{{
    // synthetic stand-in for examples/example1/example_config.json
    "time": {{ "now": "{now}", "analysis": {analysis}, "forecast": {forecast} }},
    "model": {{ "use_coupling": {use_coupling}, "use_relaxation": {use_relaxation}, "DTSecs": 30.0 }},
    "parameters": {{ "sky_view_file": "skyview.txt", "local_horizon_file": "horizons.txt" }},
    "output": {{ "step": 60, "filename": "output.json" }},
    "input": [
        {{ "name": "forecast", "path": "forecast.json", "type": "json", "source": "forecast" }},
        {{ "name": "obs", "path": "observations.json", "type": "json", "source": "observations" }}
    ]
}}
"""
    path = os.path.join(directory, "config.json")
    open(path, "w").write(cfg)
    return path
