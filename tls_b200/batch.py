"""Batches and iterative searches on one resident handle.

* :func:`batch_power` — many light curves of one campaign (same length and time span, hence the
  same period grid and template bank, main.py:53-88) through ONE ``tlsb_search_batch`` call:
  per curve the plan + search kernels, then ``spectra`` for all curves at once on the device,
  then one ``final_T0_fit`` launch per curve on the still-resident light curve.  Returns the
  vetting summary of every curve (SDE, period, T0, depth, duration, ...); the full
  ``transitleastsquaresresults`` of an interesting curve is one ``.power()`` call away.
* :func:`search_planets` — the reference's multi-planet recipe
  (``transitleastsquares/tests/test_multi_planet.py:33-40``, tutorial 03): ``power()``, mask the
  transits with ``transit_mask(t, period, 2*duration, T0)``, ``cleaned_array``, search again.

With ``torch.distributed`` initialised (one process per GPU) the curves of a batch are dealt to
the ranks round-robin — the flattened (curve, period) space cut along curves, so no chi2 row is
ever split — and the per-curve summaries are all-gathered at the end (one collective).
"""
from __future__ import annotations

import numpy as np

from . import constants as C
from . import stats
from .helpers import cleaned_array, transit_mask
from .main import transitleastsquares

SUMMARY_FIELDS = ("SDE", "SDE_raw", "period", "T0", "depth", "duration", "chi2_min", "chi2red_min", "rp_rs",
                  "transit_count", "best_row")


class BatchResults(dict):
    """dict with attribute access (like ``transitleastsquaresresults``); arrays have one entry per curve."""

    __getattr__ = dict.__getitem__


def _validate_batch(t, ys, dys):
    ys = np.ascontiguousarray(ys, dtype=np.float64)
    if ys.ndim != 2:
        raise ValueError("ys must be a [curves, n] array")
    t = np.ascontiguousarray(t, dtype=np.float64)
    if t.shape[-1] != ys.shape[1] or t.ndim not in (1, 2) or (t.ndim == 2 and t.shape != ys.shape):
        raise ValueError("t must be [n] (shared) or [curves, n]")
    if not (np.all(np.isfinite(ys)) and np.all(np.isfinite(t))):
        raise ValueError("batch inputs must be finite; run cleaned_array per curve first")
    if np.any(ys <= 0):  # validate_inputs drops such samples (helpers.cleaned_array): the curves would lose their common length
        raise ValueError("batch fluxes must be positive; run cleaned_array per curve first")
    if dys is None:  # validate.py:39-40: dy = std(y) everywhere, per curve
        dys = np.repeat(np.std(ys, axis=1)[:, None], ys.shape[1], axis=1)
        given = False
    else:
        given = True
    dys = np.ascontiguousarray(dys, dtype=np.float64)
    if dys.shape != ys.shape:
        raise ValueError("dys must have the shape of ys")
    if np.any(dys <= 0) or not np.all(np.isfinite(dys)):
        raise ValueError("dy must be positive and finite")
    if given:
        dys = dys / np.mean(dys, axis=1)[:, None]  # weights, not absolute errors (validate.py:18)
    if t.ndim == 2:
        spans = t.max(axis=1) - t.min(axis=1)
        if np.max(np.abs(spans - spans[0])) > 1e-9 * spans[0]:
            raise ValueError("all curves of a batch must cover the same time span (they share one period grid)")
    return t, ys, dys


def shard_curves(n_curves, rank, world):
    """Curves rank ``rank`` searches (round-robin)."""
    return np.arange(rank, n_curves, world)


def argmin_in_ascending_period_order(chi2_by_input, periods):
    """Index (in INPUT order) of the chi2 minimum the reference picks: it sorts by ascending period first
    (main.py:190-199), so among tied minima the SHORTEST period wins, not the first in input order."""
    asc = np.argsort(periods, kind="stable")
    return int(asc[int(np.argmin(np.asarray(chi2_by_input)[asc]))])


def summarize_curve(model, t, y, chi2_by_input, rows_by_input, depths_by_input, SDE, SDE_raw, best_index,
                    inputs, T0):
    """The scalar part of main.py:198-455 for one curve (host, O(1) + one pass over t)."""
    periods = inputs.periods
    period = periods[best_index]
    depth = depths_by_input[best_index]
    k_min = argmin_in_ascending_period_order(chi2_by_input, periods)
    best_row = int(rows_by_input[k_min])
    duration = inputs.overview["duration"][best_row]
    transit_times = stats.all_transit_times(T0, t, period)
    days = stats.calculate_transit_duration_in_days(t, period, transit_times, duration)
    chi2_min = float(chi2_by_input[k_min])
    return dict(SDE=SDE, SDE_raw=SDE_raw, period=period, T0=T0, depth=depth, duration=days, chi2_min=chi2_min,
                chi2red_min=chi2_min / (len(t) - 4), transit_count=len(transit_times), best_row=best_row,
                rp_rs=stats.rp_rs_from_depth(depth=1 - depth, law=model.limb_dark, params=model.u))


def batch_power(t, ys, dys=None, device=None, dist=None, return_power=False, **kwargs):
    """Search every row of ``ys`` with the reference's ``power(**kwargs)`` semantics.

    Returns a :class:`BatchResults` with one entry per curve for each of ``SUMMARY_FIELDS``, the
    common ``periods`` (ascending) and, when ``return_power`` is set, ``power`` ``[curves, P]``."""
    import time as _time

    from . import native

    tm = dict(prepare=0.0, upload=0.0, search=0.0, t0_fit=0.0, summaries=0.0, gather=0.0)
    tick = _time.perf_counter()
    t, ys, dys = _validate_batch(t, ys, dys)
    B, n = ys.shape
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
    mine = shard_curves(B, rank, world)
    t0_axis = t if t.ndim == 1 else t[0]
    model = transitleastsquares(t0_axis, ys[0], dys[0], verbose=False)
    if len(model.t) != n:
        raise ValueError("input validation removed samples from the first curve; clean the batch first")
    kwargs.setdefault("show_progress_bar", False)
    inputs = model.prepare(**kwargs)  # grids and bank depend on (span, n) only: shared by the batch
    no_detection = dict(SDE=0.0, SDE_raw=0.0, period=np.nan, T0=0.0, depth=1.0, duration=np.nan, rp_rs=np.nan,
                        transit_count=0)
    summary = np.full((B, len(SUMMARY_FIELDS)), np.nan)
    power = np.zeros((B, len(inputs.periods))) if return_power else None
    tm["prepare"] = _time.perf_counter() - tick
    if len(mine):
        s = native.Searcher.acquire(device=-1 if device is None else device)
        try:
            tick = _time.perf_counter()
            s.set_templates(inputs.templates, inputs.params)
            s.set_periods(inputs.periods)
            s.set_lightcurves(t if t.ndim == 1 else t[mine], ys[mine], dys[mine])
            tm["upload"] = _time.perf_counter() - tick
            tick = _time.perf_counter()
            out = s.search_batch(stats.median_window(model.oversampling_factor), want_power=return_power)
            tm["search"] = _time.perf_counter() - tick
            tick_loop = _time.perf_counter()
            for k, c in enumerate(mine):
                tc = t if t.ndim == 1 else t[c]
                chi2 = out["chi2"][k]
                if np.max(chi2) == np.min(chi2):  # main.py:209-267: nothing was fitted
                    row = dict(no_detection, chi2_min=float(chi2[0]), chi2red_min=float(chi2[0]) / (n - 4),
                               best_row=int(out["row"][k][0]))
                else:
                    best = int(out["best_index"][k])
                    k_min = argmin_in_ascending_period_order(chi2, inputs.periods)
                    signal = inputs.lc_arr[int(out["row"][k][k_min])]
                    model_in, trials = stats.t0_fit_inputs(signal, out["depth"][k][best], tc, ys[c],
                                                           inputs.periods[best], model.T0_fit_margin)
                    s.select(k)
                    tick = _time.perf_counter()
                    idx, _ = s.final_t0_fit(model_in, inputs.periods[best], trials)
                    tm["t0_fit"] += _time.perf_counter() - tick
                    T0 = trials[idx] if idx >= 0 else 0
                    row = summarize_curve(model, tc, ys[c], chi2, out["row"][k], out["depth"][k], out["SDE"][k],
                                          out["SDE_raw"][k], best, inputs, T0)
                    if return_power:
                        power[c] = out["power"][k]
                summary[c] = [row[f] for f in SUMMARY_FIELDS]
            tm["summaries"] = (_time.perf_counter() - tick_loop) - tm["t0_fit"]
        finally:
            s.release()
    if world > 1:
        tick = _time.perf_counter()
        summary = _all_gather_rows(summary, mine, B, dist, device)
        tm["gather"] = _time.perf_counter() - tick
        if return_power:
            power = _all_gather_rows(power, mine, B, dist, device)
    res = BatchResults({f: summary[:, i] for i, f in enumerate(SUMMARY_FIELDS)})
    res["periods"] = np.sort(inputs.periods)
    res["n_curves"] = B
    res["timings"] = tm  # this rank's wall-clock seconds per section (search = plan + search + spectra kernels and their copies)
    if return_power:
        res["power"] = power
    return res


def _all_gather_rows(rows, mine, n_rows, dist, device):
    """ONE all-gather of the ranks' rows (padded to the largest shard), un-dealt to curve order."""
    import torch

    world = dist.get_world_size()
    cap = (n_rows + world - 1) // world
    local = np.zeros((cap, rows.shape[1]))
    local[: len(mine)] = rows[mine]
    backend = dist.get_backend()
    dev = "cpu" if backend == "gloo" else "cuda:%d" % (torch.cuda.current_device() if device is None else device)
    src = torch.from_numpy(local).reshape(-1).to(dev)
    out = torch.empty(world * src.numel(), dtype=src.dtype, device=dev)
    dist.all_gather_into_tensor(out, src)
    out = out.cpu().numpy().reshape(world, cap, rows.shape[1])
    full = np.array(rows, copy=True)
    for r in range(world):
        idx = shard_curves(n_rows, r, world)
        full[idx] = out[r, : len(idx)]
    return full


def search_planets(t, y, dy=None, n_planets=3, SDE_min=0.0, verbose=False, timings=None, **kwargs):
    """Iterative multi-planet search: ``power()``, mask ``transit_mask(t, period, 2*duration, T0)``,
    ``cleaned_array``, repeat (tests/test_multi_planet.py:33-40).  Returns the list of results
    objects, strongest signal first; stops early when a run's SDE is below ``SDE_min`` or nothing
    was fitted.  ``timings``: a list that receives one dict of section seconds per run (``power().timings`` plus
    ``validate`` = input cleaning in the constructor and ``mask`` = masking + ``cleaned_array``)."""
    import time as _time

    t, y = np.asarray(t, dtype=float), np.asarray(y, dtype=float)
    dy = None if dy is None else np.asarray(dy, dtype=float)
    found = []
    for _ in range(n_planets):
        tick = _time.perf_counter()
        model = transitleastsquares(t, y, dy, verbose=verbose)
        t_validate = _time.perf_counter() - tick
        res = model.power(**dict(kwargs, show_progress_bar=False))
        if timings is not None:
            timings.append(dict(model.timings, validate=t_validate))
        if not np.isfinite(res.period) or res.SDE < SDE_min:
            break
        found.append(res)
        tick = _time.perf_counter()
        intransit = transit_mask(t, res.period, 2 * res.duration, res.T0)
        if dy is None:
            t, y = cleaned_array(t[~intransit], y[~intransit])
        else:
            t, y, dy = cleaned_array(t[~intransit], y[~intransit], dy[~intransit])
        if timings is not None:
            timings[-1]["mask"] = _time.perf_counter() - tick
        if len(t) < 10:
            break
    return found
