"""Transit template bank (host side, once per search).

Behaviour-matched restatement of ``/root/reference/transitleastsquares/transit.py``
(``reference_transit`` :8-42, ``fractional_transit`` :45-95, ``get_cache`` :98-160)
and of the linear interpolation it uses (``interpolation.py:7-58``), with the
limb-darkened curve coming from :mod:`tls_b200.limbdark` instead of the absent
``batman`` package.  ``pack_templates`` then flattens the ragged bank into the
arrays the C ABI takes (``include/tlsb200.h``: ``tlsb_templates``).
"""
from __future__ import annotations

import collections

import numpy as np

from . import constants as C
from . import limbdark


def _lerp_resample(x_new, x, y):
    """Piecewise-linear y(x) sampled at x_new, clamping the bracket to the end
    segments exactly like ``interpolation.py:7-26`` + ``lerp`` (:29-31)."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    x_new = np.asarray(x_new, dtype=float)
    if x_new.size == 0:
        return np.zeros(0)
    idx = np.clip(np.searchsorted(x, x_new, side="right") - 1, 0, len(x) - 2)
    theta = (x_new - x[idx]) / (x[idx + 1] - x[idx])
    return (1 - theta) * y[idx] + theta * y[idx + 1]


_REFERENCE_CACHE = {}


def reference_transit(samples, per, rp, a, inc, ecc, w, u, limb_dark):
    """In-transit part of a model transit, resampled to ``samples`` points and
    rescaled to 0 at mid-transit and 1 at the edges (transit.py:8-42).  The reference rebuilds
    it on every call (three times per ``power()``); here the few most recent shapes are kept."""
    key = (int(samples), float(per), float(rp), float(a), float(inc), float(ecc), float(w),
           tuple(float(x) for x in np.atleast_1d(u)), str(limb_dark))
    hit = _REFERENCE_CACHE.get(key)
    if hit is not None:
        return hit.copy()
    out = _reference_transit(samples, per, rp, a, inc, ecc, w, u, limb_dark)
    if len(_REFERENCE_CACHE) >= 32:
        _REFERENCE_CACHE.pop(next(iter(_REFERENCE_CACHE)))
    _REFERENCE_CACHE[key] = out
    return out.copy()


_SUPERSAMPLED = {}


def _supersampled_flux(per, rp, a, inc, ecc, w, u, limb_dark):
    """The model light curve on the reference's fixed supersampled time axis (transit.py:12-25).  It does not depend
    on ``samples``, so the limb-darkening integration runs once per transit shape and process; every
    ``reference_transit`` of that shape (the bank, and two more per ``power()`` for the model curves, main.py:316-330)
    is then a resampling."""
    key = (float(per), float(rp), float(a), float(inc), float(ecc), float(w),
           tuple(float(x) for x in np.atleast_1d(u)), str(limb_dark))
    hit = _SUPERSAMPLED.get(key)
    if hit is None:
        t = np.linspace(-0.5, 0.5, C.SUPERSAMPLE_SIZE)
        p = limbdark.TransitParams()
        p.t0, p.per, p.rp, p.a, p.inc, p.ecc, p.w = 0, per, rp, a, inc, ecc, w
        p.u, p.limb_dark = u, limb_dark
        hit = (t, limbdark.TransitModel(p, t).light_curve(p))
        if len(_SUPERSAMPLED) >= 16:
            _SUPERSAMPLED.pop(next(iter(_SUPERSAMPLED)))
        _SUPERSAMPLED[key] = hit
    return hit


def _reference_transit(samples, per, rp, a, inc, ecc, w, u, limb_dark):
    t, flux = _supersampled_flux(per, rp, a, inc, ecc, w, u, limb_dark)

    first = int(np.argmax(flux < 1))
    in_flux = flux[first : -first + 1]
    in_time = t[first : -first + 1]
    grid = np.linspace(t[first], t[-first - 1], samples)
    binned = _lerp_resample(grid, in_time, in_flux)
    lowest = np.min(binned)
    return (lowest - binned) / (lowest - 1)


def fractional_transit(
    duration,
    maxwidth,
    depth,
    samples,
    per,
    rp,
    a,
    inc,
    ecc,
    w,
    u,
    limb_dark,
    cached_reference_transit=None,
):
    """Reference transit squeezed to ``duration/maxwidth`` of ``samples`` points,
    padded with ones and scaled to ``depth`` (transit.py:45-95)."""
    if cached_reference_transit is None:
        shape = reference_transit(samples, per, rp, a, inc, ecc, w, u, limb_dark)
    else:
        shape = cached_reference_transit

    base = np.linspace(-0.5, 0.5, samples)
    occupied = int((duration / maxwidth) * samples)
    squeezed = _lerp_resample(np.linspace(-0.5, 0.5, occupied), base, shape)

    pad = np.ones(int((samples - occupied) * 0.5))
    out = np.concatenate([pad, squeezed, pad])
    if out.size < samples:
        out = np.append(out, 1.0)
    return 1 - ((1 - out) * depth)


def get_cache(durations, maxwidth_in_samples, per, rp, a, inc, ecc, w, u, limb_dark, verbose=True):
    """One trimmed template per trial duration plus its metadata (transit.py:98-160).

    Returns ``(lc_cache_overview, lc_arr)`` with the reference's dtypes:
    a structured array ``{duration f8, width_in_samples i8, overshoot f8}`` and a
    ragged object array of float64 templates."""
    if verbose:
        print("Creating model cache for", str(len(durations)), "durations")
    # The bank depends on these arguments only, and curves of one campaign (batches, the reruns of an
    # iterative search, repeated calls) ask for the same one again and again: keep the last few.
    key = (np.asarray(durations, dtype=np.float64).tobytes(), int(maxwidth_in_samples), float(per), float(rp),
           float(a), float(inc), float(ecc), float(w), tuple(float(x) for x in np.ravel(u)), str(limb_dark))
    hit = _BANKS.get(key)
    if hit is not None:
        _BANKS.move_to_end(key)
        lc_arr = np.empty(len(hit[1]), dtype=object)
        for row, signal in enumerate(hit[1]):
            lc_arr[row] = signal  # read-only arrays, shared
        return hit[0].copy(), lc_arr
    overview, lc_arr = _build_bank(durations, maxwidth_in_samples, per, rp, a, inc, ecc, w, u, limb_dark)
    for signal in lc_arr:
        signal.setflags(write=False)
    _BANKS[key] = (overview.copy(), list(lc_arr))
    while len(_BANKS) > _BANKS_MAX:
        _BANKS.popitem(last=False)
    return overview, lc_arr


_BANKS = collections.OrderedDict()
_BANKS_MAX = 8


def clear_caches():
    """Forget the cached template banks and model curves (bench.py: the cold figure of ``.power()``)."""
    _BANKS.clear()
    _REFERENCE_CACHE.clear()
    _SUPERSAMPLED.clear()


def _build_bank(durations, maxwidth_in_samples, per, rp, a, inc, ecc, w, u, limb_dark):
    rows = np.size(durations)
    overview = np.zeros(
        rows, dtype=[("duration", "f8"), ("width_in_samples", "i8"), ("overshoot", "f8")]
    )
    shape = reference_transit(maxwidth_in_samples, per, rp, a, inc, ecc, w, u, limb_dark)
    longest = np.max(durations)
    bank = []
    for row, duration in enumerate(durations):
        full = fractional_transit(
            duration, longest, C.SIGNAL_DEPTH, maxwidth_in_samples,
            per, rp, a, inc, ecc, w, u, limb_dark, cached_reference_transit=shape,
        )
        overview["duration"][row] = duration
        overview["width_in_samples"][row] = int((duration / longest) * maxwidth_in_samples)
        used = np.where(full < (1 - C.NUMERICAL_STABILITY_CUTOFF))
        signal = np.array(full[np.min(used) : np.max(used) + 1])  # own memory: the bank outlives `full`
        bank.append(signal)
        ratio = np.mean(signal) / np.min(signal)
        overview["overshoot"][row] = 1 / (2 - ratio)

    lc_arr = np.empty(rows, dtype=object)
    for row, signal in enumerate(bank):
        lc_arr[row] = signal
    return overview, lc_arr


def pack_templates(lc_arr, lc_cache_overview):
    """Flatten the ragged bank for the C ABI.

    Returns a dict of contiguous arrays: ``signal`` f8[sum L], ``offset`` i8[R],
    ``length`` i8[R], ``width`` i8[R], ``overshoot`` f8[R] (row order preserved)."""
    rows = len(lc_arr)
    length = np.array([len(s) for s in lc_arr], dtype=np.int64)
    offset = np.zeros(rows, dtype=np.int64)
    if rows > 1:
        offset[1:] = np.cumsum(length)[:-1]
    signal = (
        np.ascontiguousarray(np.concatenate([np.asarray(s, dtype=np.float64) for s in lc_arr]))
        if rows
        else np.zeros(0)
    )
    return dict(
        signal=signal,
        offset=offset,
        length=length,
        width=np.ascontiguousarray(lc_cache_overview["width_in_samples"], dtype=np.int64),
        overshoot=np.ascontiguousarray(lc_cache_overview["overshoot"], dtype=np.float64),
    )
