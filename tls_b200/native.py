"""ctypes binding of libtlsb200.so (``include/tlsb200.h``) — the only way the
package reaches the GPU.  There is deliberately no CPU fallback: every entry
point raises ``RuntimeError`` if the library is missing or CUDA is unavailable.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build as _build

_LIB = None
ABI_VERSION = "tlsb200 0.3"  # tlsb_version() must start with this (struct layouts and argtypes below)

EXPORTS = (
    "tlsb_search_periods", "tlsb_create", "tlsb_destroy", "tlsb_set_lightcurve",
    "tlsb_set_templates", "tlsb_set_periods", "tlsb_search_async", "tlsb_get_results",
    "tlsb_last_launch_count", "tlsb_last_search_kernel_ms", "tlsb_last_path_resident",
    "tlsb_last_error", "tlsb_version", "tlsb_device_count", "tlsb_set_plan_mode",
    "tlsb_plan_fallback_count", "tlsb_last_layout",
    "tlsb_final_t0_fit", "tlsb_final_t0_fit_lc", "tlsb_last_t0_fit_ms",
    "tlsb_last_path", "tlsb_last_chunk", "tlsb_set_path", "tlsb_spectra", "tlsb_last_sort_info", "tlsb_last_block", "tlsb_resolve_plan", "tlsb_plan_repair_count",
    "tlsb_set_lightcurves", "tlsb_select_lightcurve", "tlsb_lightcurve_count", "tlsb_search_batch",
    "tlsb_current_device", "tlsb_last_tiled_widths", "tlsb_set_filter", "tlsb_last_filter_stats",
    "tlsb_set_inputs_async", "tlsb_unshard_records",
)

_c_i64 = ctypes.c_int64
_c_vp = ctypes.c_void_p


class LightCurve(ctypes.Structure):
    _fields_ = [("t", _c_vp), ("y", _c_vp), ("dy", _c_vp), ("n", _c_i64)]


class Templates(ctypes.Structure):
    _fields_ = [("signal", _c_vp), ("offset", _c_vp), ("length", _c_vp), ("width", _c_vp),
                ("overshoot", _c_vp), ("rows", _c_i64)]


class Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in (
        "transit_depth_min", "R_star_min", "R_star_max", "M_star_min", "M_star_max", "T0_fit_margin")]


class Exec(ctypes.Structure):
    _fields_ = [("devices", _c_vp), ("n_devices", ctypes.c_int32)]


def library_path():
    return _build.LIB


def lib():
    """Load (building first if the sources are newer and nvcc is present)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("TLSB200_LIB") or _build.LIB  # TLSB200_LIB: an experimental build (scripts/gpu_variants.sh)
    if path == _build.LIB and _build.is_stale():
        try:
            _build.nvcc_path()
            have_nvcc = True
        except RuntimeError:
            have_nvcc = False
        if have_nvcc:
            _build.build()  # a compile error in edited sources must surface, never a silently stale binary
        elif not os.path.exists(path):
            raise RuntimeError("libtlsb200.so is missing and there is no nvcc to build it")
        else:
            import warnings

            warnings.warn("libtlsb200.so is older than its sources and there is no nvcc here to rebuild it: using the shipped binary")
    L = ctypes.CDLL(path)
    L.tlsb_last_error.restype = ctypes.c_char_p
    L.tlsb_version.restype = ctypes.c_char_p
    L.tlsb_device_count.restype = ctypes.c_int32
    L.tlsb_current_device.restype = ctypes.c_int32
    L.tlsb_last_launch_count.restype = _c_i64
    L.tlsb_last_launch_count.argtypes = [_c_vp]
    L.tlsb_last_search_kernel_ms.restype = ctypes.c_double
    L.tlsb_last_search_kernel_ms.argtypes = [_c_vp]
    L.tlsb_last_path_resident.restype = ctypes.c_int32
    L.tlsb_last_path_resident.argtypes = [_c_vp]
    L.tlsb_plan_repair_count.restype = _c_i64
    L.tlsb_plan_repair_count.argtypes = [_c_vp]
    L.tlsb_resolve_plan.argtypes = [_c_vp, _c_vp, _c_vp]
    L.tlsb_plan_fallback_count.restype = _c_i64
    L.tlsb_plan_fallback_count.argtypes = [_c_vp]
    L.tlsb_set_plan_mode.argtypes = [_c_vp, ctypes.c_int32]
    L.tlsb_last_layout.argtypes = [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp]
    L.tlsb_create.argtypes = [ctypes.POINTER(_c_vp), ctypes.c_int32]
    L.tlsb_destroy.argtypes = [_c_vp]
    L.tlsb_set_lightcurve.argtypes = [_c_vp, ctypes.POINTER(LightCurve)]
    L.tlsb_set_templates.argtypes = [_c_vp, ctypes.POINTER(Templates), ctypes.POINTER(Params)]
    L.tlsb_set_periods.argtypes = [_c_vp, _c_vp, _c_i64]
    L.tlsb_search_async.argtypes = [_c_vp, _c_vp, _c_vp]
    L.tlsb_get_results.argtypes = [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]
    L.tlsb_search_periods.argtypes = [
        ctypes.POINTER(LightCurve), _c_vp, _c_i64, ctypes.POINTER(Templates), ctypes.POINTER(Params),
        ctypes.POINTER(Exec), _c_vp, _c_vp, _c_vp, _c_vp]
    L.tlsb_final_t0_fit.argtypes = [_c_vp, _c_vp, _c_vp, _c_i64, ctypes.c_double, _c_vp, _c_i64, _c_vp, _c_vp]
    L.tlsb_final_t0_fit_lc.argtypes = [ctypes.POINTER(LightCurve), ctypes.c_int32, _c_vp, _c_i64, ctypes.c_double,
                                       _c_vp, _c_i64, _c_vp, _c_vp]
    L.tlsb_last_path.restype = ctypes.c_int32
    L.tlsb_last_path.argtypes = [_c_vp]
    L.tlsb_last_block.restype = ctypes.c_int32
    L.tlsb_last_block.argtypes = [_c_vp]
    L.tlsb_last_tiled_widths.restype = ctypes.c_int32
    L.tlsb_last_tiled_widths.argtypes = [_c_vp]
    L.tlsb_last_chunk.restype = ctypes.c_int32
    L.tlsb_last_chunk.argtypes = [_c_vp]
    L.tlsb_set_path.argtypes = [_c_vp, ctypes.c_int32, ctypes.c_int32]
    L.tlsb_last_sort_info.argtypes = [_c_vp, _c_vp, _c_vp, _c_vp]
    L.tlsb_set_filter.argtypes = [_c_vp, ctypes.c_int32, ctypes.c_int32]
    L.tlsb_set_inputs_async.argtypes = [_c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64]
    L.tlsb_unshard_records.argtypes = [_c_vp, _c_i64, ctypes.c_int32, _c_vp, _c_vp]
    L.tlsb_last_filter_stats.argtypes = [_c_vp, _c_vp, _c_vp, _c_vp]
    L.tlsb_spectra.argtypes = [ctypes.c_int32, _c_vp, _c_i64, _c_i64, _c_i64, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp]
    L.tlsb_set_lightcurves.argtypes = [_c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_i64, ctypes.c_int32]
    L.tlsb_select_lightcurve.argtypes = [_c_vp, _c_i64]
    L.tlsb_lightcurve_count.restype = _c_i64
    L.tlsb_lightcurve_count.argtypes = [_c_vp]
    L.tlsb_search_batch.argtypes = [_c_vp, _c_vp, _c_i64] + [_c_vp] * 8
    L.tlsb_last_t0_fit_ms.restype = ctypes.c_double
    L.tlsb_last_t0_fit_ms.argtypes = [_c_vp]
    version = L.tlsb_version().decode()
    if not version.startswith(ABI_VERSION):
        raise RuntimeError("libtlsb200.so reports %r but this binding was written for %r: rebuild the library" % (version, ABI_VERSION))
    _LIB = L
    return L


def _check(rc, what):
    if rc != 0:
        msg = lib().tlsb_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _ptr(a):
    return a.ctypes.data  # an int: ctypes converts it for c_void_p arguments and fields (data_as costs 3x as much)


class _Packed(object):
    """Keeps the numpy buffers alive next to the ctypes structs that point into them."""

    def __init__(self, t, y, dy, templates, params):
        self.t, self.y, self.dy = _f64(t), _f64(y), _f64(dy)
        if not (len(self.t) == len(self.y) == len(self.dy)):
            raise ValueError("Arrays (t, y, dy) must be of the same dimensions")
        self.signal = _f64(templates["signal"])
        self.offset, self.length = _i64(templates["offset"]), _i64(templates["length"])
        self.width, self.overshoot = _i64(templates["width"]), _f64(templates["overshoot"])
        self.lc = LightCurve(_ptr(self.t), _ptr(self.y), _ptr(self.dy), len(self.t))
        self.tp = Templates(_ptr(self.signal), _ptr(self.offset), _ptr(self.length),
                            _ptr(self.width), _ptr(self.overshoot), len(self.width))
        self.prm = Params(*[float(params[n]) for n, _ in Params._fields_])


def device_count():
    return int(lib().tlsb_device_count())


def search_periods(t, y, dy, periods, templates, params, devices=None, return_t0_index=False):
    """One batched search through ``tlsb_search_periods`` with HOST buffers.

    Replaces the period loop of main.py:140-185.  Returns ``(chi2, row, depth)`` (and the
    window-start index of the best model when asked) in the order of ``periods``."""
    L = lib()
    pk = _Packed(t, y, dy, templates, params)
    periods = _f64(periods)
    P = len(periods)
    chi2 = np.empty(P, np.float64)
    row = np.empty(P, np.int64)
    depth = np.empty(P, np.float64)
    t0 = np.empty(P, np.int64)
    if devices is None:
        ex = None
    else:
        devs = np.ascontiguousarray(np.atleast_1d(devices), dtype=np.int32)
        ex = Exec(_ptr(devs), len(devs))
    rc = L.tlsb_search_periods(ctypes.byref(pk.lc), _ptr(periods), P, ctypes.byref(pk.tp),
                               ctypes.byref(pk.prm), ctypes.byref(ex) if ex is not None else None,
                               _ptr(chi2), _ptr(row), _ptr(depth), _ptr(t0))
    _check(rc, "tlsb_search_periods")
    return (chi2, row, depth, t0) if return_t0_index else (chi2, row, depth)


def spectra(chi2, median_window, device=None):
    """``tlsb_spectra``: chi2 ``[P]`` or ``[curves, P]`` -> ``(SR, power_raw, power, SDE_raw, SDE, argmax)``
    (stats.py:105-132 on the device).  Scalars come back as arrays of length ``curves`` for 2-D input."""
    chi2 = _f64(chi2)
    single = chi2.ndim == 1
    rows = np.atleast_2d(chi2)
    C, P = rows.shape
    SR, pr, pw = np.empty_like(rows), np.empty_like(rows), np.empty_like(rows)
    sde_raw, sde = np.empty(C), np.empty(C)
    amax = np.empty(C, np.int64)
    rc = lib().tlsb_spectra(-1 if device is None else int(device), _ptr(rows), P, C, int(median_window),
                            _ptr(SR), _ptr(pr), _ptr(pw), _ptr(sde_raw), _ptr(sde), _ptr(amax))
    _check(rc, "tlsb_spectra")
    if single:
        return SR[0], pr[0], pw[0], float(sde_raw[0]), float(sde[0]), int(amax[0])
    return SR, pr, pw, sde_raw, sde, amax


def final_t0_fit(t, y, dy, model_in, period, trials, device=None):
    """All trial epochs of stats.py:165-202 in one launch (``tlsb_final_t0_fit_lc``, HOST buffers).

    Returns ``(best_index, residuals)``; ``T0 = trials[best_index]``."""
    t, y, dy = _f64(t), _f64(y), _f64(dy)
    model_in, trials = _f64(model_in), _f64(trials)
    lc = LightCurve(_ptr(t), _ptr(y), _ptr(dy), len(t))
    resid = np.empty(len(trials), np.float64)
    best = ctypes.c_int64(-1)
    rc = lib().tlsb_final_t0_fit_lc(ctypes.byref(lc), -1 if device is None else int(device), _ptr(model_in),
                                    len(model_in), float(period), _ptr(trials), len(trials), _ptr(resid),
                                    ctypes.byref(best))
    _check(rc, "tlsb_final_t0_fit_lc")
    return int(best.value), resid


class Searcher(object):
    """Handle API: inputs stay resident in HBM between searches."""

    def __init__(self, device=-1):
        self._h = _c_vp()
        _check(lib().tlsb_create(ctypes.byref(self._h), int(device)), "tlsb_create")
        self._keep = None
        self.n_periods = 0

    def close(self):
        if self._h:
            lib().tlsb_destroy(self._h)
            self._h = _c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # Creating and (above all) destroying a handle costs tens to hundreds of milliseconds of
    # cudaMalloc / cudaFree - more than a whole cfg-1 search - so the drivers that open a handle
    # per call (batch_power, ShardedSearch) borrow one from a per-device pool instead.  Every
    # setter replaces the state it owns, so a recycled handle carries nothing over.
    _POOL = {}

    @classmethod
    def acquire(cls, device=-1):
        device = int(device)
        if device < 0:
            device = int(lib().tlsb_current_device())
        idle = cls._POOL.get(device)
        if idle:
            s = idle.pop()
            s.set_path("auto")
            s.set_plan_mode(0)
            s.set_filter(os.environ.get("TLSB_FILTER", "1") != "0", False)
            return s
        s = cls(device=device)
        s._pool_key = device
        return s

    def release(self):
        """Back to the pool (or destroyed when it was not acquired from it)."""
        key = getattr(self, "_pool_key", None)
        if key is None or not self._h:
            self.close()
            return
        self._keep = None
        Searcher._POOL.setdefault(key, []).append(self)

    @classmethod
    def drain_pool(cls):
        for idle in cls._POOL.values():
            for s in idle:
                s.close()
        cls._POOL.clear()

    def set_inputs(self, t, y, dy, templates, params):
        pk = _Packed(t, y, dy, templates, params)
        _check(lib().tlsb_set_lightcurve(self._h, ctypes.byref(pk.lc)), "tlsb_set_lightcurve")
        _check(lib().tlsb_set_templates(self._h, ctypes.byref(pk.tp), ctypes.byref(pk.prm)), "tlsb_set_templates")
        self._keep = pk

    def set_inputs_async(self, t, y, dy, templates, params, periods, stream=None):
        """``tlsb_set_inputs_async``: light curve, bank and periods in one call, uploads asynchronous on ``stream``, no
        synchronisation.  The arrays are kept alive by this object; they must not be MODIFIED until the stream has
        passed the uploads (e.g. until the next results have been read)."""
        pk = _Packed(t, y, dy, templates, params)
        periods = _f64(periods)
        _check(lib().tlsb_set_inputs_async(self._h, _c_vp(stream or 0), ctypes.byref(pk.lc), ctypes.byref(pk.tp),
                                           ctypes.byref(pk.prm), _ptr(periods), len(periods)), "tlsb_set_inputs_async")
        self._keep = (pk, periods)
        self.n_periods = len(periods)

    def set_lightcurve(self, t, y, dy):
        t, y, dy = _f64(t), _f64(y), _f64(dy)
        lc = LightCurve(_ptr(t), _ptr(y), _ptr(dy), len(t))
        _check(lib().tlsb_set_lightcurve(self._h, ctypes.byref(lc)), "tlsb_set_lightcurve")

    def set_templates(self, templates, params):
        pk = _Packed(np.zeros(1), np.zeros(1), np.ones(1), templates, params)
        _check(lib().tlsb_set_templates(self._h, ctypes.byref(pk.tp), ctypes.byref(pk.prm)), "tlsb_set_templates")

    def set_lightcurves(self, t, ys, dys):
        """Several curves of the same length: ``t`` is ``[n]`` (shared) or ``[curves, n]``."""
        ys, dys, t = _f64(ys), _f64(dys), _f64(t)
        if ys.ndim != 2 or dys.shape != ys.shape:
            raise ValueError("ys and dys must be [curves, n] arrays of the same shape")
        shared = t.ndim == 1
        if (t.shape[-1] != ys.shape[1]) or (not shared and t.shape != ys.shape):
            raise ValueError("t must be [n] or [curves, n]")
        _check(lib().tlsb_set_lightcurves(self._h, _ptr(t), _ptr(ys), _ptr(dys), ys.shape[1], ys.shape[0],
                                          1 if shared else 0), "tlsb_set_lightcurves")

    def select(self, index):
        _check(lib().tlsb_select_lightcurve(self._h, int(index)), "tlsb_select_lightcurve")

    @property
    def n_curves(self):
        return int(lib().tlsb_lightcurve_count(self._h))

    def search_batch(self, median_window, stream=None, want_power=True, want_records=True):
        """``tlsb_search_batch`` -> dict(chi2, row, depth, t0_index, power, SDE_raw, SDE, best_index)."""
        B, P = self.n_curves, self.n_periods
        out = dict(SDE_raw=np.empty(B), SDE=np.empty(B), best_index=np.empty(B, np.int64))
        if want_records:
            out.update(chi2=np.empty((B, P)), row=np.empty((B, P), np.int64), depth=np.empty((B, P)),
                       t0_index=np.empty((B, P), np.int64))
        if want_power:
            out["power"] = np.empty((B, P))
        opt = lambda k: _ptr(out[k]) if k in out else None  # noqa: E731
        _check(lib().tlsb_search_batch(self._h, _c_vp(stream or 0), int(median_window), opt("chi2"), opt("row"),
                                       opt("depth"), opt("t0_index"), opt("power"), _ptr(out["SDE_raw"]),
                                       _ptr(out["SDE"]), _ptr(out["best_index"])), "tlsb_search_batch")
        return out

    def set_periods(self, periods):
        periods = _f64(periods)
        _check(lib().tlsb_set_periods(self._h, _ptr(periods), len(periods)), "tlsb_set_periods")
        self.n_periods = len(periods)

    def search_async(self, stream=None, records_ptr=None):
        _check(lib().tlsb_search_async(self._h, _c_vp(stream or 0), _c_vp(records_ptr or 0)), "tlsb_search_async")

    def results(self, stream=None):
        P = self.n_periods
        chi2 = np.empty(P, np.float64)
        row = np.empty(P, np.int64)
        depth = np.empty(P, np.float64)
        t0 = np.empty(P, np.int64)
        _check(lib().tlsb_get_results(self._h, _c_vp(stream or 0), _ptr(chi2), _ptr(row), _ptr(depth), _ptr(t0)),
               "tlsb_get_results")
        return chi2, row, depth, t0

    def final_t0_fit(self, model_in, period, trials, stream=None):
        """``tlsb_final_t0_fit`` on the resident light curve -> ``(best_index, residuals)``."""
        model_in, trials = _f64(model_in), _f64(trials)
        resid = np.empty(len(trials), np.float64)
        best = ctypes.c_int64(-1)
        _check(lib().tlsb_final_t0_fit(self._h, _c_vp(stream or 0), _ptr(model_in), len(model_in), float(period),
                                       _ptr(trials), len(trials), _ptr(resid), ctypes.byref(best)),
               "tlsb_final_t0_fit")
        return int(best.value), resid

    @property
    def t0_fit_ms(self):
        return float(lib().tlsb_last_t0_fit_ms(self._h))

    def resolve_plan(self, stream=None, records_ptr=None):
        """``tlsb_resolve_plan``: settle the periods the device plan flagged, re-search those that changed."""
        _check(lib().tlsb_resolve_plan(self._h, _c_vp(stream or 0), _c_vp(records_ptr or 0)), "tlsb_resolve_plan")

    @property
    def plan_repairs(self):
        return int(lib().tlsb_plan_repair_count(self._h))

    def set_plan_mode(self, mode):
        """0 device plan, 1 exact host plan, 2 device plan flagging every period, 3 = 2 + wrong ranges (tests)."""
        _check(lib().tlsb_set_plan_mode(self._h, int(mode)), "tlsb_set_plan_mode")

    PATHS = {0: "auto", 1: "resident", 2: "tiled", 3: "streaming"}

    def set_path(self, path, chunk=0):
        """Force a kernel layout: 'auto', 'resident', 'tiled' or 'streaming'; ``chunk`` caps the tiled
        path's chunk capacity in doubles (tests)."""
        code = {v: k for k, v in self.PATHS.items()}[path] if isinstance(path, str) else int(path)
        _check(lib().tlsb_set_path(self._h, code, int(chunk)), "tlsb_set_path")

    def set_filter(self, on=True, count_stats=False):
        """Equal weights: fp32 filter pass on (default) or off (every gate survivor through the exact fp64
        evaluation; same results, slower); ``count_stats`` counts survivors / finalists on the device."""
        _check(lib().tlsb_set_filter(self._h, 1 if on else 0, 1 if count_stats else 0), "tlsb_set_filter")

    @property
    def filter_stats(self):
        """dict(candidates, finalists, overflows) of the last search (needs ``set_filter(count_stats=True)``)."""
        c, f, o = _c_i64(0), _c_i64(0), _c_i64(0)
        _check(lib().tlsb_last_filter_stats(self._h, ctypes.byref(c), ctypes.byref(f), ctypes.byref(o)),
               "tlsb_last_filter_stats")
        return dict(candidates=c.value, finalists=f.value, overflows=o.value)

    @property
    def path(self):
        return self.PATHS.get(int(lib().tlsb_last_path(self._h)), "?")

    @property
    def chunk(self):
        return int(lib().tlsb_last_chunk(self._h))

    @property
    def sort_info(self):
        """Tiled path: dict(segment_capacity, n_segments, global_sort_periods) of the last search."""
        cap, ns, fb = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64()
        _check(lib().tlsb_last_sort_info(self._h, ctypes.byref(cap), ctypes.byref(ns), ctypes.byref(fb)),
               "tlsb_last_sort_info")
        return dict(segment_capacity=cap.value, n_segments=ns.value, global_sort_periods=fb.value)

    @property
    def plan_fallbacks(self):
        return int(lib().tlsb_plan_fallback_count(self._h))

    @property
    def layout(self):
        th, cp, qc = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        sm = ctypes.c_int64()
        _check(lib().tlsb_last_layout(self._h, ctypes.byref(th), ctypes.byref(cp), ctypes.byref(qc), ctypes.byref(sm)),
               "tlsb_last_layout")
        return dict(threads=th.value, ctas_per_sm=cp.value, queue_capacity=qc.value, smem_bytes=sm.value,
                    resident=self.resident, path=self.path, chunk=self.chunk,
                    block=int(lib().tlsb_last_block(self._h)),
                    tiled_widths=int(lib().tlsb_last_tiled_widths(self._h)))

    @property
    def launch_count(self):
        return int(lib().tlsb_last_launch_count(self._h))

    @property
    def kernel_ms(self):
        return float(lib().tlsb_last_search_kernel_ms(self._h))

    @property
    def resident(self):
        return bool(lib().tlsb_last_path_resident(self._h))


def unshard_records(gathered_ptr, n_periods, world, out_ptr, stream=None):
    """``tlsb_unshard_records``: rank-major all-gathered records -> 3 * n_periods + 1 words in the job's period order
    (device pointers; asynchronous on ``stream``)."""
    _check(lib().tlsb_unshard_records(_c_vp(gathered_ptr), int(n_periods), int(world), _c_vp(out_ptr), _c_vp(stream or 0)),
           "tlsb_unshard_records")


def unpack_records(records, n_periods):
    """Split the 3-plane device record layout (chi2 | depth | row+t0<<32 [| status]) copied to host."""
    rec = np.asarray(records)[: 3 * n_periods].reshape(3, n_periods)
    chi2 = rec[0].view(np.float64)
    depth = rec[1].view(np.float64)
    # packed = row | t0_index << 32 (little endian): the low and high halves as strided 32-bit views, widened once
    row = rec[2].view(np.uint32)[0::2].astype(np.int64)
    t0 = rec[2].view(np.int32)[1::2].astype(np.int64)
    return chi2, row, depth, t0
