"""Inputs for the parity tests that reach code the synthetic weather rarely does (shared by the CPU N-version test
and the GPU parity test)."""
import numpy as np


def extreme_cases():
    """Inputs built to reach code the synthetic weather rarely does: storage caps and the snow / ice / deposit
    transitions under heavy precipitation, a hot summer surface, other hemispheres and longitudes, a leap day
    and a turn of the year."""
    import datetime as dt
    from roadsurf_b200 import synth
    out = []
    a, s, p, _ = synth.make_case(4, 8, seed=61)
    a.prec *= 25.0; a.tair -= 6.0; a.tdew -= 6.0                                   # heavy snowfall
    out.append(("heavy snow", a, s, p))
    a, s, p, _ = synth.make_case(4, 8, seed=62)
    a.prec *= 25.0; a.tair[:] = np.where(np.arange(a.sim_len) < a.sim_len // 2, a.tair - 5.0, a.tair + 6.0)   # snow, then thaw + rain
    a.tdew[:] = a.tair - 0.5
    out.append(("snow then thaw and rain", a, s, p))
    a, s, p, _ = synth.make_case(4, 8, seed=63, start=dt.datetime(2019, 7, 1, 6, 0, 0), sky_view_fraction=1.0)
    a.tair += 28.0; a.tdew += 20.0; a.SW *= 1.0; a.LW += 80.0                       # hot summer day
    out.append(("hot summer day", a, s, p))
    a, s, p, _ = synth.make_case(6, 6, seed=64, start=dt.datetime(2020, 2, 28, 21, 0, 0), sky_view_fraction=1.0)
    for k, (lat, lon) in enumerate(((-33.9, 151.2), (-54.8, -68.3), (0.0, 0.0), (64.1, -21.9), (35.7, 139.7), (78.2, 15.6))):
        a.local[k].lat, a.local[k].lon = lat, lon                                  # leap day, six places on the globe
    out.append(("leap day around the globe", a, s, p))
    a, s, p, _ = synth.make_case(3, 6, seed=65, start=dt.datetime(2019, 12, 31, 21, 0, 0), sky_view_fraction=1.0)
    out.append(("turn of the year", a, s, p))
    return out
