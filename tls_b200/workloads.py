"""Deterministic synthetic light curves for the BASELINE.json configurations.

Shared by ``bench.py``, the tests and ``oracle/make_golden.py`` so that every
side reads identical bytes.  Definitions follow SURVEY.md §8(d): the planet of
``/root/reference/tutorials/01 Quick start with synthetic data.ipynb`` cell 1
(quadratic limb darkening u=[0.4,0.4], rp=6371/696342, a=19, inc=90) injected
with :mod:`tls_b200.limbdark`, plus seeded Gaussian noise from the legacy numpy
generator (``numpy.random.seed`` / ``normal``), ``dy=None``.
"""
from __future__ import annotations

import numpy as np

from . import limbdark

WORKLOADS = {
    # name: (t0, t1, N, planet period [d], noise [ppm], seed, power() kwargs)
    "cfg1": (3.14, 93.14, 4320, 10.123, 50, 0, {}),          # 90 d @ 30 min (tutorial 01 shape)
    "tutorial01": (3.14, 103.14, 4800, 10.123, 50, 0, {}),   # the notebook's exact 100 d
    "cfg1_500ppm": (3.14, 93.14, 4320, 10.123, 500, 0, {}),
    "cfg2": (3.14, 1464.14, 70128, 10.123, 50, 1, {}),       # Kepler-long, 4 yr @ 30 min
    "cfg3": (3.14, 30.14, 19440, 3.3, 500, 2, {"duration_grid_step": 1.02}),  # TESS sector @ 2 min
    "small": (3.14, 33.14, 720, 4.3, 200, 3, {}),             # quick CPU-sized case
}


def inject(t, period, t0, rp=6371.0 / 696342.0, a=19.0, inc=90.0, u=(0.4, 0.4)):
    p = limbdark.TransitParams()
    p.t0, p.per, p.rp, p.a, p.inc, p.ecc, p.w = t0, period, rp, a, inc, 0.0, 90.0
    p.u, p.limb_dark = list(u), "quadratic"
    return limbdark.TransitModel(p, t, n_nodes=192).light_curve(p)


def lightcurve(name, hetero=False, planets=None):
    """Return ``(t, y, dy, power_kwargs)`` for a named workload.

    ``hetero=True`` adds per-point uncertainties ``dy = sigma * U(0.5, 2)`` (the
    regime of the reference's ``tests/test_uncertainties.py:43-47``)."""
    t0, t1, n, planet_period, ppm, seed, kw = WORKLOADS[name]
    t = np.linspace(t0, t1, n)
    flux = np.ones(n)
    for per in planets or [planet_period]:
        flux = flux * inject(t, per, t0)
    np.random.seed(seed)
    sigma = ppm * 1e-6
    y = flux + np.random.normal(0, sigma, n)
    dy = None
    if hetero:
        dy = sigma * np.random.uniform(0.5, 2.0, n)
    return t, y, dy, dict(kw)


_PROFILE = {}


def _phase_profile(rp=6371.0 / 696342.0, a=19.0, u=(0.4, 0.4), half_width=0.03, n=12001):
    """The transit of ``inject`` as a function of orbital PHASE only (circular orbit, inc = 90: the projected
    separation is a*sin(2 pi phase), whatever the period), tabulated once: batches of synthetic curves are then O(n)
    interpolations instead of one limb-darkening integration per curve."""
    key = (rp, a, tuple(u), half_width, n)
    if key not in _PROFILE:
        ph = np.linspace(-half_width, half_width, n)
        _PROFILE[key] = (ph, inject(ph, 1.0, 0.0, rp=rp, a=a, u=u))
    return _PROFILE[key]


def batch_lightcurves(n_curves, first_seed=1000, only=None):
    """cfg-4: ``n_curves`` light curves shaped as cfg-1 (90 d @ 30 min, 4320 points, shared time
    stamps), curve c seeded with ``first_seed + c``: planet period ~ U(1, 40) d, epoch ~ U(0, period), white noise
    ~ logU(50, 500) ppm.  ``only``: the curve indices to generate (a rank's shard); the others stay 1.0.
    Returns ``(t, ys)``."""
    t = np.linspace(3.14, 93.14, 4320)
    ph_tab, f_tab = _phase_profile()
    ys = np.ones((n_curves, len(t)))
    for c in (range(n_curves) if only is None else only):
        rng = np.random.RandomState(first_seed + int(c))
        per = rng.uniform(1.0, 40.0)
        t0 = 3.14 + rng.uniform(0.0, per)
        ppm = 10 ** rng.uniform(np.log10(50.0), np.log10(500.0))
        x = (t - t0) / per
        x = x - np.rint(x)  # phase in [-0.5, 0.5)
        ys[c] = np.interp(x, ph_tab, f_tab, left=1.0, right=1.0) + rng.normal(0.0, ppm * 1e-6, len(t))
    return t, ys
