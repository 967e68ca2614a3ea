"""Light-curve hygiene helpers kept for API compatibility with
``/root/reference/transitleastsquares/helpers.py`` (host side, O(N) once)."""
from __future__ import annotations

import numpy as np


def _valid(values):
    """helpers.py:22-29 — not None, not NaN, strictly positive and finite."""
    arr = np.asarray(values)
    if arr.dtype == object:
        arr = np.array([np.nan if v is None else v for v in arr], dtype=float)
    else:
        arr = arr.astype(float, copy=False)
    if np.ma.isMaskedArray(values):
        arr = np.ma.filled(np.ma.asarray(values, dtype=float), np.nan)
    with np.errstate(invalid="ignore"):
        return arr, (~np.isnan(arr)) & (arr > 0) & (arr < np.inf)


def cleaned_array(t, y, dy=None):
    """Drop every sample where t, y (or dy) is None/NaN/masked/non-positive/inf
    (helpers.py:18-61; vectorised instead of the reference's Python loop)."""
    n = len(y)
    tt, ok_t = _valid(t)
    yy, ok_y = _valid(y)
    keep = ok_t[:n] & ok_y[:n]
    if dy is None:
        return tt[:n][keep].astype(float), yy[:n][keep].astype(float)
    dd, ok_d = _valid(dy)
    keep &= ok_d[:n]
    return tt[:n][keep].astype(float), yy[:n][keep].astype(float), dd[:n][keep].astype(float)


def transit_mask(t, period, duration, T0):
    """True for samples within duration/2 of a transit centre (helpers.py:64-67)."""
    return np.abs((t - T0 + 0.5 * period) % period - 0.5 * period) < 0.5 * duration


def resample(time, flux, factor):
    """Linear-interpolation rebinning by ``factor`` (helpers.py:8-15)."""
    from .transit import _lerp_resample

    n_new = int(len(flux) / factor)
    grid = np.linspace(np.min(time), np.max(time), n_new)
    return grid, _lerp_resample(grid, time, flux)


def running_mean(data, width_signal):
    """Window mean through cumulative sums (helpers.py:70-73)."""
    cs = np.cumsum(np.insert(data, 0, 0))
    return (cs[width_signal:] - cs[:-width_signal]) / float(width_signal)


def _pad_to(values, n):
    """Repeat the first/last value so the result has length n (helpers.py:100-108)."""
    missing = n - len(values)
    front = int(missing * 0.5)
    return np.concatenate(
        [np.full(front, values[0]), values, np.full(missing - front, values[-1])]
    )


def running_median(data, kernel):
    """Sliding median of width ``kernel``, edge-padded to len(data) (helpers.py:93-108)."""
    kernel = int(kernel)
    windows = np.lib.stride_tricks.sliding_window_view(np.asarray(data), kernel)
    return _pad_to(np.median(windows, axis=1), len(data))


def impact_to_inclination(b, semimajor_axis):
    """Impact parameter -> inclination [deg] (helpers.py:111-113)."""
    return np.degrees(np.arccos(b / semimajor_axis))
