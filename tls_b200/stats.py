"""Post-search statistics (host side, O(P) or O(N) once per search).

Everything here runs AFTER the GPU period search and turns its per-period
(chi2, row, depth) arrays into the SDE spectrum and the descriptive numbers of
the results object.  Each function restates, in behaviour including quirks, the
function of the same name in ``/root/reference/transitleastsquares/stats.py``
(line ranges in the docstrings).  Out of scope as GPU work in this round
(SURVEY.md §8f); ``final_T0_fit`` already runs on the GPU.
"""
from __future__ import annotations

from os import path

import numpy as np

from . import constants as C
from .helpers import running_median, transit_mask

_FAP_TABLE = None


def fold(time, period, T0):
    """Phase fold with epoch T0 (core.py:9-12).  Written in the
    reciprocal-multiply form that numba's fastmath build of the reference
    actually executes, so phases are bit-identical."""
    x = (np.asarray(time, dtype=float) - T0) * (1.0 / period)
    return x - np.floor(x)


def FAP(SDE):
    """False-alarm probability looked up from the reference's white-noise table
    (stats.py:10-18; table = the reference's fap.csv stored as arrays)."""
    global _FAP_TABLE
    if _FAP_TABLE is None:
        with np.load(path.join(C.resources_dir, "fap_table.npz")) as z:
            _FAP_TABLE = (z["fap"], z["sde"])
    fap, sde = _FAP_TABLE
    return fap[np.argmax(sde > SDE)]


_LD_WEIGHTS = {
    # law: (number of parameters, function giving the flux-weighted correction)
    "linear": (1, lambda p: 1 - p[0] / 3),
    "quadratic": (2, lambda p: 1 - p[0] / 3 - p[1] / 6),
    "squareroot": (2, lambda p: 1 - p[0] / 3 - p[1] / 5),
    "logarithmic": (2, lambda p: 1 + 2 * p[1] / 9 - p[0] / 3),
    "nonlinear": (4, lambda p: 1 - p[0] / 5 - p[1] / 3 - 3 * p[2] / 7 - p[3] / 2),
}


def rp_rs_from_depth(depth, law, params):
    """Planet/star radius ratio from the maximum depth, Heller (2019)
    (stats.py:21-69; same validation messages)."""
    laws = "linear, quadratic, squareroot, logarithmic, nonlinear"
    values = [params] if isinstance(params, (int, float)) else list(params)
    if not all(isinstance(v, (float, int, np.floating, np.integer)) and not isinstance(v, bool) for v in values):
        raise ValueError("All limb-darkening parameters must be numbers")
    if law not in laws:
        raise ValueError("Please provide a supported limb-darkening law:", laws)
    count, weight = _LD_WEIGHTS[law]
    if len(values) != count:
        if count == 1:
            raise ValueError("Please provide exactly one parameter")
        if count == 2:
            raise ValueError("Please provide exactly two limb-darkening parameters")
        raise ValueError("Please provide exactly four limb-darkening parameters")
    return (depth * weight([float(v) for v in values])) ** (1 / 2)


def pink_noise(data, width):
    """Mean over all windows of std(window)/sqrt(width) (stats.py:72-77)."""
    data = np.asarray(data, dtype=float)
    windows = len(data) - width + 1
    if windows <= 0:
        return 0 / windows  # as the reference: an empty loop, then the division by the window count
    # every window's population variance from two cumulative sums of the mean-shifted data (O(N)
    # instead of one numpy.std call per window; agrees with the per-window loop to ~1e-12 relative)
    x = data - np.mean(data)
    c1 = np.concatenate([[0.0], np.cumsum(x)])
    c2 = np.concatenate([[0.0], np.cumsum(x * x)])
    s1 = c1[width:] - c1[:-width]
    s2 = c2[width:] - c2[:-width]
    var = np.maximum(s2 / width - (s1 / width) ** 2, 0.0)
    return float(np.sum(np.sqrt(var) / width ** 0.5) / windows)


def period_uncertainty(periods, power):
    """Half of the full width at half maximum of the highest peak; ``inf`` when
    the peak touches the grid edge (stats.py:80-102)."""
    try:
        top = int(np.argmax(power))
        half = 0.5 * power[top]
        hi = top + 1
        while not power[hi] <= half:
            hi += 1
        lo = top - 1
        while not power[lo] <= half:  # negative indices wrap exactly like the reference
            lo -= 1
        return 0.5 * (periods[hi] - periods[lo])
    except Exception:
        return float("inf")


def median_window(oversampling_factor):
    """The reference's running-median kernel (stats.py:115-117)."""
    kernel = oversampling_factor * C.SDE_MEDIAN_KERNEL_SIZE
    if kernel % 2 == 0:
        kernel = kernel + 1
    return int(kernel)


def spectra(chi2, oversampling_factor, device=None):
    """chi2[P] -> (SR, power_raw, power, SDE_raw, SDE)  (stats.py:105-132).

    Runs on the B200 through ``tlsb_spectra`` (``include/tlsb200.h``): SR, its mean and
    population std, the running median of width ``median_window`` (helpers.py:93-108) and the
    re-normalisation are three small kernels; no CPU fallback."""
    from . import native

    SR, power_raw, power, SDE_raw, SDE, _ = native.spectra(chi2, median_window(oversampling_factor), device=device)
    return SR, power_raw, power, SDE_raw, SDE


def t0_fit_inputs(signal, depth, t, y, period, T0_fit_margin):
    """Host-side inputs of the T0 scan (stats.py:141-156): the in-transit model scaled to the
    fitted depth and the grid of trial epochs."""
    dur = len(signal)
    scale = C.SIGNAL_DEPTH / (1 - depth)
    model_in = 1 - ((1 - np.asarray(signal, dtype=float)) / scale)
    n = np.size(y)
    points = n if T0_fit_margin == 0 else int(n / (T0_fit_margin * dur))
    points = min(points, n)
    trials = np.linspace(start=np.min(t), stop=np.min(t) + period, num=points)
    return model_in, trials


def final_T0_fit(signal, depth, t, y, dy, period, T0_fit_margin, show_progress_bar, verbose, device=None, dist=None):
    """Scan mid-transit epochs at the best period and return the best T0 (stats.py:135-204).

    The loop over trial epochs (stats.py:165-202: fold, stable argsort, roll, weighted
    residuals) runs on the B200 in ONE launch through ``tlsb_final_t0_fit_lc``
    (``include/tlsb200.h``); there is no CPU fallback.  Quirk kept on purpose (SURVEY.md
    §3.3): the reference overwrites its weights with a second roll of the already rolled
    flux (stats.py:191), so the residuals are weighted by 1/flux^2 and ``dy`` has no effect.
    ``show_progress_bar`` is accepted and ignored (the scan takes a millisecond)."""
    from . import native

    model_in, trials = t0_fit_inputs(signal, depth, t, y, period, T0_fit_margin)
    if verbose:
        print("Searching for best T0 for period", format(period, ".5f"), "days")
    if len(trials) == 0:
        return 0
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1 and len(trials) >= 4 * dist.get_world_size():
        # one process per GPU: trial epoch k -> rank k mod world (the epochs are independent), ONE all-gather of the
        # residuals, then the reference's strict '<' scan from +inf over ALL of them (stats.py:200-202)
        from .distributed import all_gather_interleaved

        rank, world = dist.get_rank(), dist.get_world_size()
        _, mine = native.final_t0_fit(t, y, dy, model_in, period, trials[rank::world], device=device)
        resid = all_gather_interleaved(mine, len(trials), dist, device=device)
        best, lowest = -1, np.inf
        below = np.flatnonzero(resid < np.inf)
        if len(below):
            k = int(np.argmin(resid[below]))  # first minimum among the finite ones = the strict-'<' winner
            best, lowest = int(below[k]), resid[below][k]
        return trials[best] if best >= 0 else 0
    best, _ = native.final_t0_fit(t, y, dy, model_in, period, trials, device=device)
    return trials[best] if best >= 0 else 0  # stats.py:163: T0 = 0 when nothing is below +inf


def all_transit_times(T0, t, period):
    """Mid-transit times inside the time series (stats.py:245-262)."""
    tmin, tmax = np.min(t), np.max(t)
    first = T0 + period if T0 < tmin else T0
    times = [first]
    end = tmin + (tmax - tmin)
    while times[-1] + period < end:
        times.append(times[-1] + period)
    return times


def calculate_stretch(t, period, transit_times):
    """(time span / period) / number of epochs (stats.py:279-291)."""
    return ((np.max(t) - np.min(t)) / period) / len(transit_times)


def calculate_fill_factor(t):
    """Fraction of cadences present assuming a constant cadence (stats.py:294-301)."""
    cadence = np.median(np.diff(t))
    return (len(t) - 1) / ((np.max(t) - np.min(t)) / cadence)


def calculate_transit_duration_in_days(t, period, transit_times, duration):
    """Fractional duration -> days, corrected for gaps (stats.py:265-276)."""
    raw = duration * calculate_stretch(t, period, transit_times) * period
    return raw * calculate_fill_factor(t)


def model_lightcurve(transit_times, period, t, model_transit_single):
    """Tile the single-transit model over all epochs (one extra on either side)
    and crop to the data span (stats.py:207-242)."""
    epochs = np.concatenate([[transit_times[0] - period], transit_times, [transit_times[-1] + period]])
    samples = (int(len(t) / len(transit_times))) * C.OVERSAMPLE_MODEL_LIGHT_CURVE
    xs = np.concatenate(
        [np.linspace(e - period / 2, e + period / 2, samples) for e in epochs]
    ) if len(epochs) else np.array([])
    ys = np.concatenate([model_transit_single for _ in epochs]) if len(epochs) else np.array([])
    if np.all(np.isnan(xs)):
        return None, None
    start = np.nanargmax(xs > np.min(t))
    stop = np.nanargmax(xs > np.max(t))
    return ys[start:stop], xs[start:stop]


def count_stats(t, y, transit_times, transit_duration_in_days):
    """Points in transit and in the two neighbouring bins of one duration, summed
    over epochs that lie fully inside the data (stats.py:304-341)."""
    inside = after = before = 0
    d = transit_duration_in_days
    tmin, tmax = np.min(t), np.max(t)
    ascending = _is_ascending(t)

    def between(lo, hi):  # number of samples with lo < t < hi
        if ascending:
            return max(0, int(np.searchsorted(t, hi, side="left") - np.searchsorted(t, lo, side="right")))
        return int(np.count_nonzero((t > lo) & (t < hi)))

    for mid in transit_times:
        a, b, c, e = mid - 1.5 * d, mid - 0.5 * d, mid + 0.5 * d, mid + 1.5 * d
        if a > tmin and e < tmax:
            inside += between(b, c)
            before += between(a, b)
            after += between(c, e)
    return inside, after, before


def _is_ascending(t):
    """Time stamps in non-decreasing order (the usual case): windows become binary searches."""
    return bool(np.all(t[1:] >= t[:-1]))


def _epoch_window(t, mid, d, ascending=False):
    """Index of the samples with mid - d/2 < t < mid + d/2 (ascending order), None for a NaN epoch."""
    lo, hi = mid - 0.5 * d, mid + 0.5 * d
    if np.isnan(lo) or np.isnan(hi):
        return None
    if ascending:  # the same samples in the same order as the mask below, found in O(log n)
        start = int(np.searchsorted(t, lo, side="right"))
        return slice(start, max(start, int(np.searchsorted(t, hi, side="left"))))
    return np.where(np.logical_and(t > lo, t < hi))


def intransit_stats(t, y, transit_times, transit_duration_in_days):
    """Per-epoch depths/counts and the odd/even in-transit flux sets
    (stats.py:344-420)."""
    odd_parts, even_parts = [np.array([])], [np.array([])]
    n_epochs = len(transit_times)
    counts = np.zeros([n_epochs])
    depths = np.zeros([n_epochs])
    errors = np.zeros([n_epochs])
    m_odd = m_even = s_odd = s_even = np.nan
    ascending = _is_ascending(t)
    for i, mid in enumerate(transit_times):
        sel = _epoch_window(t, mid, transit_duration_in_days, ascending)
        flux = y[sel] if sel is not None else np.array([])
        n_in = np.size(flux)
        depths[i] = np.mean(flux) if n_in > 0 else np.nan
        errors[i] = np.std(flux) / np.sqrt(n_in) if n_in > 0 else np.nan
        counts[i] = n_in
        (even_parts if i % 2 == 0 else odd_parts).append(flux)
    # the reference re-evaluates these after every epoch (stats.py:399-408); only the last values survive
    odd, even = np.concatenate(odd_parts), np.concatenate(even_parts)
    if len(odd) > 0:
        m_odd = np.mean(odd)
        s_odd = np.std(odd) / len(odd) ** 0.5
    if len(even) > 0:
        m_even = np.mean(even)
        s_even = np.std(even) / len(even) ** 0.5
    return m_odd, m_even, s_odd, s_even, odd, even, counts, depths, errors


def snr_stats(t, y, period, duration, T0, transit_times, transit_duration_in_days, per_transit_count):
    """Per-epoch white and pink signal-to-noise (stats.py:423-469)."""
    n_epochs = len(transit_times)
    snr = np.zeros([n_epochs])
    snr_pink = np.zeros([n_epochs])
    outside = y[~transit_mask(t, period, 2 * duration, T0)]
    try:
        pink = pink_noise(outside, int(np.mean(per_transit_count)))
    except Exception:
        pink = np.nan
    std = np.std(outside) if len(outside) > 0 else np.nan
    ascending = _is_ascending(t)
    for i, mid in enumerate(transit_times):
        sel = _epoch_window(t, mid, transit_duration_in_days, ascending)
        flux = y[sel] if sel is not None else np.array([])
        n_in = np.size(flux)
        mean_flux = np.mean(flux) if n_in > 0 else np.nan
        try:
            snr_pink[i] = (1 - mean_flux) / pink
            if n_in > 0 and not np.isnan(std):
                snr[i] = (1 - mean_flux) / (std / n_in ** 0.5)
            else:
                snr[i] = 0
                snr_pink[i] = 0
        except Exception:
            snr[i] = 0
            snr_pink[i] = 0
    return snr, snr_pink
