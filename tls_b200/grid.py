"""Period and duration grids (host side, O(P) scalar math once per search).

Behaviour-matched restatement of ``/root/reference/transitleastsquares/grid.py``:
``T14`` (:9-32), ``duration_grid`` (:35-56), ``period_grid`` (:59-156).  These
are the *inputs* of the GPU hot path, so their values are pinned by
``tests/test_host.py`` against the reference's own golden numbers
(its ``tests/test_period_grid.py``, ``tests/test_duration_grid.py``) and, in the
build container, bit for bit against the reference itself (``tests/test_reference_diff.py``).
"""
from __future__ import annotations

import math
import warnings

import numpy as np

from . import constants as C

ONE_THIRD = 1 / 3


def T14(R_s, M_s, P, upper_limit=C.FRACTIONAL_TRANSIT_DURATION_MAX, small=False):
    """Longest plausible transit duration as a fraction of the period P [days]
    for a star of radius R_s [R_sun] and mass M_s [M_sun]  (grid.py:9-32).

    ``small`` selects the small-planet limit; otherwise a 2 R_jup planet is assumed.
    The operation order is the reference's so the value is bit-identical to the
    numba-compiled original (checked on full grids).
    """
    P = P * C.SECONDS_PER_DAY
    R_s = C.R_sun * R_s
    M_s = C.M_sun * M_s
    cube = ((4 * P) / (math.pi * C.G * M_s)) ** ONE_THIRD
    if small:
        t14 = R_s * cube
    else:
        t14 = (R_s + 2 * C.R_jup) * cube
    frac = t14 / P
    return upper_limit if frac > upper_limit else frac


def duration_grid(periods, shortest=None, log_step=C.DURATION_GRID_STEP):
    """Geometric grid of fractional durations (grid.py:35-56).

    The end points come from the package-level stellar limits, *not* from the
    user's ``R_star_min`` etc., and ``shortest`` is accepted but unused — both as
    in the reference."""
    longest = T14(C.R_STAR_MAX, C.M_STAR_MAX, np.min(periods), small=False)
    d = T14(C.R_STAR_MIN, C.M_STAR_MIN, np.max(periods), small=True)
    grid = [d]
    while d * log_step < longest:
        d = d * log_step
        grid.append(d)
    grid.append(longest)
    return grid


def _clamp_star(value, lo, hi, name, lo_reset=None):
    """Range clamps with the reference's warning texts (grid.py:71-105)."""
    if value < lo:
        warnings.warn(
            "Warning: %s was set to %s for period_grid (was unphysical: %s)" % (name, lo, value)
        )
        return lo if lo_reset is None else lo_reset
    if value > hi:
        warnings.warn(
            "Warning: %s was set to %s for period_grid (was unphysical: %s)" % (name, hi, value)
        )
        return hi
    return value


def period_grid(
    R_star,
    M_star,
    time_span,
    period_min=0,
    period_max=float("inf"),
    oversampling_factor=C.OVERSAMPLING_FACTOR,
    n_transits_min=C.N_TRANSITS_MIN,
):
    """Trial periods [days], descending, with the cubic-in-frequency spacing of
    Ofir (2014, A&A 561, A138) (grid.py:59-156)."""
    # grid.py:71-78 resets a too-small radius to 0.1 although the text says 0.01
    R_star = _clamp_star(R_star, 0.01, 10000, "R_star", lo_reset=0.1)
    M_star = _clamp_star(M_star, 0.01, 1000, "M_star")

    R = R_star * C.R_sun
    M = M_star * C.M_sun
    span = time_span * C.SECONDS_PER_DAY

    f_min = n_transits_min / span
    f_max = 1.0 / (2 * math.pi) * math.sqrt(C.G * M / (3 * R) ** 3)

    # Ofir 2014 eqs. (5)-(7)
    A = (
        (2 * math.pi) ** (2.0 / 3)
        / math.pi
        * R
        / (C.G * M) ** (1.0 / 3)
        / (span * oversampling_factor)
    )
    offset = f_min ** (1.0 / 3) - A / 3.0
    n_opt = (f_max ** (1.0 / 3) - f_min ** (1.0 / 3) + A / 3) * 3 / A

    k = np.arange(n_opt) + 1
    periods = 1 / (A / 3 * k + offset) ** 3 / C.SECONDS_PER_DAY
    keep = np.where(np.logical_and(periods > period_min, periods <= period_max))
    count = np.size(periods[keep])

    if count > 10 ** 6:
        warnings.warn(
            "period_grid generates a very large grid ("
            + str(count)
            + "). Recommend to check physical plausibility for stellar mass, radius, and time series duration."
        )
    if count < C.MINIMUM_PERIOD_GRID_SIZE:
        if span < 5 * C.SECONDS_PER_DAY:
            span = 5 * C.SECONDS_PER_DAY
        warnings.warn(
            "period_grid defaults to R_star=1 and M_star=1 as given density yielded grid with too few values"
        )
        return period_grid(R_star=1, M_star=1, time_span=span / C.SECONDS_PER_DAY)
    return periods[keep]
