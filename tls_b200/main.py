"""``transitleastsquares(t, y, dy).power(**kwargs)`` — the drop-in entry point.

Host orchestration with the same call sequence as
``/root/reference/transitleastsquares/main.py:44-455``; the one thing that
changes is the period loop (main.py:121-196): instead of mapping
``core.search_period`` over a process pool, ALL trial periods go to the B200 in
one batched call through the C ABI (``include/tlsb200.h: tlsb_search_periods``).
There is no CPU fallback: if the CUDA library or a GPU is missing the call
raises.
"""
from __future__ import annotations

import multiprocessing
import warnings

import numpy as np

from . import constants as C
from . import stats
from .grid import duration_grid, period_grid
from .helpers import transit_mask
from .results import transitleastsquaresresults
from .transit import fractional_transit, get_cache, pack_templates
from .validate import validate_args, validate_inputs


class SearchInputs(object):
    """Everything the period search needs, in the layout of the C ABI."""

    __slots__ = ("t", "y", "dy", "periods", "templates", "params", "lc_arr", "overview", "durations")

    def __init__(self, t, y, dy, periods, lc_arr, overview, durations, params):
        self.t = np.ascontiguousarray(t, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        self.dy = np.ascontiguousarray(dy, dtype=np.float64)
        self.periods = np.ascontiguousarray(periods, dtype=np.float64)
        self.lc_arr = lc_arr
        self.overview = overview
        self.durations = durations
        self.templates = pack_templates(lc_arr, overview)
        self.params = params  # dict: transit_depth_min, R/M_star_min/max, T0_fit_margin


class transitleastsquares(object):
    """Compute the transit least squares of limb-darkened transit models (main.py:44-49)."""

    def __init__(self, t, y, dy=None, verbose=True):
        self.t, self.y, self.dy = validate_inputs(t, y, dy)
        self.verbose = verbose
        self.timings = {}

    # ------------------------------------------------------------------ host prep
    def prepare(self, **kwargs):
        """validate kwargs -> period grid -> duration grid -> template bank
        (main.py:53-88).  Returns a :class:`SearchInputs`."""
        self, kwargs = validate_args(self, kwargs)
        periods = period_grid(
            R_star=self.R_star,
            M_star=self.M_star,
            time_span=np.max(self.t) - np.min(self.t),
            period_min=self.period_min,
            period_max=self.period_max,
            oversampling_factor=self.oversampling_factor,
            n_transits_min=self.n_transits_min,
        )
        durations = duration_grid(periods, shortest=1 / len(self.t), log_step=self.duration_grid_step)
        maxwidth = int(np.max(durations) * np.size(self.y))
        if maxwidth % 2 != 0:
            maxwidth += 1
        overview, lc_arr = get_cache(
            durations=durations, maxwidth_in_samples=maxwidth, per=self.per, rp=self.rp,
            a=self.a, inc=self.inc, ecc=self.ecc, w=self.w, u=self.u,
            limb_dark=self.limb_dark, verbose=self.verbose,
        )
        params = dict(
            transit_depth_min=float(self.transit_depth_min),
            R_star_min=float(self.R_star_min), R_star_max=float(self.R_star_max),
            M_star_min=float(self.M_star_min), M_star_max=float(self.M_star_max),
            T0_fit_margin=float(self.T0_fit_margin),
        )
        return SearchInputs(self.t, self.y, self.dy, periods, lc_arr, overview, durations, params)

    # ------------------------------------------------------------------ the hot path
    def _search(self, inputs, devices, dist=None):
        """main.py:121-196 replaced by one call into the CUDA library.  With ``dist`` (an initialised
        ``torch.distributed`` module, one process per GPU) the periods are dealt to the ranks and
        the per-period records all-gathered once (tls_b200/distributed.py); every rank gets the
        whole result and carries on with identical post-processing."""
        from . import native

        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            from .distributed import search_periods_distributed

            device = None if devices is None else int(np.atleast_1d(devices)[0])
            return search_periods_distributed(inputs.t, inputs.y, inputs.dy, inputs.periods, inputs.templates,
                                              inputs.params, dist, device=device)
        return native.search_periods(
            inputs.t, inputs.y, inputs.dy, inputs.periods, inputs.templates, inputs.params,
            devices=devices,
        )

    def _spectra(self, chi2):
        """stats.py:105-132 on the GPU (``tlsb_spectra``)."""
        return stats.spectra(chi2, self.oversampling_factor, device=getattr(self, "_t0_device", None))

    def _final_T0_fit(self, signal, depth, period):
        """stats.py:135-204 on the GPU (``tlsb_final_t0_fit_lc``)."""
        return stats.final_T0_fit(
            signal=signal, depth=depth, t=self.t, y=self.y, dy=self.dy, period=period,
            T0_fit_margin=self.T0_fit_margin, show_progress_bar=self.show_progress_bar,
            verbose=self.verbose, device=getattr(self, "_t0_device", None), dist=getattr(self, "_dist", None),
        )

    def power(self, **kwargs):
        """Compute the periodogram for a set of user-defined parameters (main.py:51).

        ``self.timings`` holds the wall-clock seconds of the call's sections afterwards (prepare = grids + template
        bank, search, spectra, t0_fit, statistics): what bench.py reports as the shares of ``.power()``."""
        import time as _time

        tick = _time.perf_counter()
        self.timings = {}
        inputs = self.prepare(**kwargs)
        self.timings["prepare"] = _time.perf_counter() - tick
        if self.verbose:
            print(C.VERSION)
        periods = inputs.periods
        durations = inputs.durations
        lc_arr, overview = inputs.lc_arr, inputs.overview
        devices = kwargs.get("devices", kwargs.get("device", None))

        if self.verbose:
            print(
                "Searching " + str(len(self.y)) + " data points, " + str(len(periods))
                + " periods from " + str(round(min(periods), 3)) + " to "
                + str(round(max(periods), 3)) + " days"
            )
            print("Using the B200 search kernels (use_threads=%d is accepted and ignored)" % self.use_threads)

        dist = kwargs.get("dist", None)
        self._dist = dist
        if dist is not None and devices is None:
            import torch

            devices = torch.cuda.current_device()  # one process per GPU: this rank's device
        tick = _time.perf_counter()
        chi2_by_input, rows_by_input, depths_by_input = self._search(inputs, devices, dist)
        self.timings["search"] = _time.perf_counter() - tick
        self._t0_device = None if devices is None else int(np.atleast_1d(devices)[0])

        # main.py:190-196: ascending period order (the grid arrives descending: a stable sort sees one run)
        order = np.argsort(periods, kind="stable")
        test_statistic_periods = periods[order]
        chi2 = np.asarray(chi2_by_input)[order]
        rows = np.asarray(rows_by_input)[order]
        depths = np.asarray(depths_by_input)[order]
        tick = _time.perf_counter()
        res = self._postprocess(test_statistic_periods, chi2, rows, depths, lc_arr, overview, durations)
        self.timings["statistics"] = (_time.perf_counter() - tick) - self.timings.get("spectra", 0.0) - self.timings.get("t0_fit", 0.0)
        return res

    # ------------------------------------------------------------------ host post
    def _postprocess(self, test_statistic_periods, chi2, rows, depths, lc_arr, overview, durations):
        """main.py:198-455: spectra, T0 fit, statistics, results object."""
        t, y, dy = self.t, self.y, self.dy
        idx_best = np.argmin(chi2)
        best_row = rows[idx_best]
        duration = overview["duration"][best_row]
        maxwidth_in_samples = int(np.max(durations) * np.size(t))

        no_fit = np.max(chi2) == np.min(chi2)
        if no_fit:
            warnings.warn('No transit were fit. Try smaller "transit_depth_min"')

        chi2red = chi2 / (len(t) - 4)
        chi2_min = np.min(chi2)
        chi2red_min = np.min(chi2red)
        nan = np.nan

        if no_fit:  # main.py:216-267
            power_raw = np.zeros(len(chi2))
            power = np.zeros(len(chi2))
            period, depth, SR, SDE, SDE_raw, T0 = nan, 1, 0, 0, 0, 0
            transit_times = transit_duration = nan
            folded_phase = folded_y = folded_dy = nan
            model_folded_phase = model_folded_model = nan
            model_lightcurve_model = model_lightcurve_time = nan
            m_odd = m_even = s_odd = s_even = nan
            per_transit_count = transit_depths = transit_depths_unc = nan
            snr_per_transit = snr_pink_per_transit = nan
            depth_mean = depth_mean_std = snr = rp_rs = nan
            odd_even_mismatch = nan
            transit_count = empty_transit_count = distinct_transit_count = nan
            duration = nan
            in_count = after_count = before_count = nan
        else:
            import time as _time

            tick = _time.perf_counter()
            SR, power_raw, power, SDE_raw, SDE = self._spectra(chi2)
            self.timings["spectra"] = _time.perf_counter() - tick
            top = np.argmax(power)
            period = test_statistic_periods[top]
            depth = depths[top]
            tick = _time.perf_counter()
            T0 = self._final_T0_fit(lc_arr[best_row], depth, period)
            self.timings["t0_fit"] = _time.perf_counter() - tick
            transit_times = stats.all_transit_times(T0, t, period)
            transit_duration = stats.calculate_transit_duration_in_days(t, period, transit_times, duration)

            phases = stats.fold(t, period, T0=T0 + period / 2)
            order = np.argsort(phases)
            folded_phase, folded_y, folded_dy = phases[order], y[order], dy[order]
            half_cadence = 1 / np.size(t) / 2
            model_folded_phase = np.linspace(0 + half_cadence, 1 + half_cadence, np.size(t))

            fill_half = 1 - ((1 - stats.calculate_fill_factor(t)) * 0.5)
            stretch = stats.calculate_stretch(t, period, transit_times)
            internal_samples = (int(len(y) / len(transit_times))) * C.OVERSAMPLE_MODEL_LIGHT_CURVE
            shape = dict(per=self.per, rp=self.rp, a=self.a, inc=self.inc, ecc=self.ecc,
                         w=self.w, u=self.u, limb_dark=self.limb_dark)
            model_folded_model = fractional_transit(
                duration=duration * maxwidth_in_samples * fill_half,
                maxwidth=maxwidth_in_samples / stretch, depth=1 - depth,
                samples=int(len(t / len(transit_times))), **shape)
            model_transit_single = fractional_transit(
                duration=(duration * maxwidth_in_samples),
                maxwidth=maxwidth_in_samples / stretch, depth=1 - depth,
                samples=internal_samples, **shape)
            model_lightcurve_model, model_lightcurve_time = stats.model_lightcurve(
                transit_times, period, t, model_transit_single)

            (m_odd, m_even, s_odd, s_even, flux_odd, flux_even, per_transit_count,
             transit_depths, transit_depths_unc) = stats.intransit_stats(
                t, y, transit_times, transit_duration)
            flux_in = np.concatenate([flux_odd, flux_even])
            snr_per_transit, snr_pink_per_transit = stats.snr_stats(
                t=t, y=y, period=period, duration=duration, T0=T0, transit_times=transit_times,
                transit_duration_in_days=transit_duration, per_transit_count=per_transit_count)
            flux_out = y[~transit_mask(t, period, 2 * duration, T0)]
            depth_mean = np.mean(flux_in)
            depth_mean_std = np.std(flux_in) / np.sum(per_transit_count) ** 0.5
            snr = ((1 - depth_mean) / np.std(flux_out)) * len(flux_in) ** 0.5
            rp_rs = stats.rp_rs_from_depth(depth=1 - depth, law=self.limb_dark, params=self.u)

            m_odd, s_odd = _mean_and_error(flux_odd)
            m_even, s_even = _mean_and_error(flux_even)
            in_count, after_count, before_count = stats.count_stats(t, y, transit_times, transit_duration)
            odd_even_mismatch = abs(m_odd - m_even) / (s_odd + s_even)

            transit_count = len(transit_times)
            empty_transit_count = np.count_nonzero(per_transit_count == 0)
            distinct_transit_count = transit_count - empty_transit_count
            duration = transit_duration
            if empty_transit_count / transit_count >= 0.33:
                warnings.warn(
                    str(empty_transit_count) + " of " + str(transit_count)
                    + " transits without data. The true period may be twice the given period."
                )

        return transitleastsquaresresults(
            SDE, SDE_raw, chi2_min, chi2red_min, period,
            stats.period_uncertainty(test_statistic_periods, power), T0, duration, depth,
            (depth_mean, depth_mean_std), (m_even, s_even), (m_odd, s_odd),
            transit_depths, transit_depths_unc, rp_rs, snr, snr_per_transit,
            snr_pink_per_transit, odd_even_mismatch, transit_times, per_transit_count,
            transit_count, distinct_transit_count, empty_transit_count, stats.FAP(SDE),
            in_count, after_count, before_count, test_statistic_periods, power, power_raw,
            SR, chi2, chi2red, model_lightcurve_time, model_lightcurve_model,
            model_folded_phase, folded_y, folded_dy, folded_phase, model_folded_model,
        )


def _mean_and_error(values):
    """mean and std/sqrt(n), NaNs for an empty set (main.py:372-390)."""
    if len(values) > 0:
        return np.mean(values), np.std(values) / len(values) ** 0.5
    return np.nan, np.nan
