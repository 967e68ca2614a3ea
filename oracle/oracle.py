"""TEST INFRASTRUCTURE — CPU oracle of the TLS period search.  NOT product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  Two restatements of
``/root/reference/transitleastsquares/core.py:96-188`` (``search_period``):

* :func:`search_periods_c` — ``liboracle.so`` (``tls_oracle.c``), all host threads;
* :func:`search_period_numpy` — a slow, line-by-line numpy/Python version used to
  cross-check the C one on small inputs.

Parity is PINNED: both are checked against outputs of the reference's own numba
path run in the build container (``tests/golden/*.npz`` via ``make_golden.py``).
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Templates(ctypes.Structure):
    _fields_ = [
        ("signal", ctypes.c_void_p), ("offset", ctypes.c_void_p), ("length", ctypes.c_void_p),
        ("width", ctypes.c_void_p), ("overshoot", ctypes.c_void_p), ("rows", ctypes.c_int64),
    ]


class _Params(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in (
        "transit_depth_min", "R_star_min", "R_star_max", "M_star_min", "M_star_max", "T0_fit_margin")]


def build(force=False):
    """Compile liboracle.so with the committed Makefile (building the checker is not using it)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("tls_oracle.c", "oracle_fold.c", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.tls_oracle_search.restype = ctypes.c_int
        _LIB.tls_oracle_threads.restype = ctypes.c_int
    return _LIB


def host_threads():
    return int(_lib().tls_oracle_threads())


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def search_periods_c(t, y, dy, periods, templates, params, threads=0):
    """All periods through liboracle.so.  ``templates`` = dict from
    ``tls_b200.transit.pack_templates``; ``params`` = dict of the six scalars.
    Returns (chi2 f8[P], row i8[P], depth f8[P]) in the order of ``periods``."""
    t = np.ascontiguousarray(t, np.float64)
    y = np.ascontiguousarray(y, np.float64)
    dy = np.ascontiguousarray(dy, np.float64)
    periods = np.ascontiguousarray(periods, np.float64)
    keep = {k: np.ascontiguousarray(v) for k, v in templates.items()}
    tp = _Templates(_ptr(keep["signal"]), _ptr(keep["offset"]), _ptr(keep["length"]),
                    _ptr(keep["width"]), _ptr(keep["overshoot"]), len(keep["width"]))
    pr = _Params(*[float(params[n]) for n, _ in _Params._fields_])
    P = len(periods)
    chi2 = np.empty(P, np.float64)
    row = np.empty(P, np.int64)
    depth = np.empty(P, np.float64)
    rc = _lib().tls_oracle_search(
        _ptr(t), _ptr(y), _ptr(dy), ctypes.c_int64(len(t)), _ptr(periods), ctypes.c_int64(P),
        ctypes.byref(tp), ctypes.byref(pr), ctypes.c_int(int(threads)),
        _ptr(chi2), _ptr(row), _ptr(depth))
    if rc != 0:
        raise RuntimeError("tls_oracle_search failed with code %d" % rc)
    return chi2, row, depth


# ---------------------------------------------------------------- numpy restatement
_G, _RSUN, _RJUP, _MSUN, _DAY = 6.673e-11, 695508000, 69911000, 1.989 * 10 ** 30, 86400


def _t14(R_s, M_s, P, small):
    """grid.py:9-32"""
    P = P * _DAY
    R_s = _RSUN * R_s
    M_s = _MSUN * M_s
    cube = ((4 * P) / (math.pi * _G * M_s)) ** (1 / 3)
    t14 = R_s * cube if small else (R_s + 2 * _RJUP) * cube
    out = t14 / P
    return 0.12 if out > 0.12 else out


def search_period_numpy(period, t, y, dy, templates, params):
    """One period, following core.py:96-188 statement by statement (slow)."""
    widths_by_row = templates["width"]
    uniq = np.unique(widths_by_row)                      # core.py:113
    M = int(max(uniq))
    if M % 2 != 0:
        M += 1
    r = 1.0 / period                                     # core.py:15-18 under fastmath
    x = t * r
    phases = x - np.floor(x)
    order = np.argsort(phases, kind="mergesort")         # core.py:120
    flux = y[order]
    err = dy[order]
    p_dy = np.append(err, err[:M])                       # core.py:126-132
    inv = 1 / p_dy ** 2
    data = np.append(flux, flux[:M])
    regular = np.sum(((1 - flux) ** 2) * 1 / err ** 2)   # core.py:21-25
    patched = np.sum(((1 - data) ** 2) * inv)
    edge = patched - regular
    n = len(y)
    dmax = _t14(params["R_star_max"], params["M_star_max"], period, False)
    dmin = _t14(params["R_star_min"], params["M_star_min"], period, True)
    naive = (max(t) - min(t)) / period                   # core.py:148-151
    corr = (naive + 1) / naive
    wmin = int(np.floor(dmin * n))
    wmax = int(np.ceil(dmax * n * corr))
    uniq = uniq[(uniq >= wmin) & (uniq <= wmax)]
    cs = np.cumsum(np.insert(data, 0, 0))                # helpers.py:70-73
    wdd = ((1 - data) ** 2) * inv
    best, best_row, best_depth = float("inf"), 0, 0
    margin = params["T0_fit_margin"]
    for W in uniq:
        W = int(W)
        row = int(np.argmax(widths_by_row == W))         # core.py:163-165
        signal = templates["signal"][templates["offset"][row]: templates["offset"][row] + templates["length"][row]]
        overshoot = templates["overshoot"][row]
        mean = 1 - (cs[W:] - cs[:-W]) / float(W)
        # core.py:79-93 (sequential recurrence)
        ootr = np.empty(len(data) - W + 1)
        ootr[0] = patched - np.sum(wdd[:W])
        for i in range(1, len(ootr)):
            ootr[i] = ootr[i - 1] + wdd[i - 1] - wdd[i - 1 + W]
        xth = 1                                          # core.py:50-55
        if margin > 0 and W > margin:
            xth = max(1, int(W / (1 / margin)))
        this = float(n)
        this_depth = 0
        q = 1 - signal
        for i in range(len(mean)):
            if mean[i] > params["transit_depth_min"] and i % xth == 0:
                target = mean[i] * overshoot
                rs = 1 / (0.5 / target)
                seg = data[i: i + len(signal)]
                res = np.sum(((seg - (1 - q * rs)) ** 2) * inv[i: i + len(signal)])
                stat = res + ootr[i] - edge
                if stat < this:
                    this, this_depth = stat, 1 - target
        if this < best:
            best, best_row, best_depth = this, row, this_depth
    return best, best_row, best_depth


# ---------------------------------------------------------------- final_T0_fit restatement
SIGNAL_DEPTH = 0.5  # tls_constants.py:71


def final_T0_fit_numpy(signal, depth, t, y, dy, period, T0_fit_margin, trials=None):
    """stats.py:135-204 statement by statement; returns ``(T0, residuals_total per trial, trials)``.

    ``fold`` (core.py:9-12) is written in the reciprocal-multiply form numba's fastmath build
    executes.  The ``dy`` overwrite of stats.py:191 is kept: after the first iteration the
    weights are the rolled flux rolled again, and because ``dy = dy[sort_index]`` is itself
    overwritten before use (stats.py:175 then :191), ``dy`` never reaches the residuals."""
    signal = np.asarray(signal, dtype=float)
    dur = len(signal)
    scale = SIGNAL_DEPTH / (1 - depth)
    model_in = 1 - ((1 - signal) / scale)
    n = np.size(y)
    if trials is None:
        if T0_fit_margin == 0:
            points = n
        else:
            points = int(n / (T0_fit_margin * dur))
        if points > n:
            points = n
        trials = np.linspace(start=np.min(t), stop=np.min(t) + period, num=points)
    ones = np.ones(len(y[dur:]))
    lowest, T0 = float("inf"), 0
    out = np.empty(len(trials))
    r = 1.0 / period
    for k, Tx in enumerate(trials):
        x = (t - Tx) * r
        phases = x - np.floor(x)
        idx = np.argsort(phases, kind="mergesort")
        flux = y[idx]
        roll = int(dur / 2) + 1
        flux = np.concatenate([flux[-roll:], flux[:-roll]])
        w = np.concatenate([flux[-roll:], flux[:-roll]])
        total = np.sum((flux[:dur] - model_in) ** 2 / w[:dur] ** 2) + np.sum((flux[dur:] - ones) ** 2 / w[dur:] ** 2)
        out[k] = total
        if total < lowest:
            lowest, T0 = total, Tx
    return T0, out, trials


# ---------------------------------------------------------------- spectra restatement
def running_median_numpy(data, kernel):
    """helpers.py:93-108"""
    idx = np.arange(kernel) + np.arange(len(data) - kernel + 1)[:, None]
    idx = idx.astype(np.int64)
    med = np.median(data[idx], axis=1)
    missing = len(data) - len(med)
    front = int(missing * 0.5)
    med = np.append(np.full(front, med[0]), med)
    return np.append(med, np.full(missing - front, med[-1]))


def spectra_numpy(chi2, oversampling_factor, kernel=None):
    """stats.py:105-132 statement by statement -> (SR, power_raw, power, SDE_raw, SDE)."""
    SR = np.min(chi2) / chi2
    SDE_raw = (1 - np.mean(SR)) / np.std(SR)
    power_raw = SR - np.mean(SR)
    scale = SDE_raw / np.max(power_raw)
    power_raw = power_raw * scale
    if kernel is None:
        kernel = oversampling_factor * 30  # tls_constants.py:100
        if kernel % 2 == 0:
            kernel = kernel + 1
    if len(power_raw) > 2 * kernel:
        power = power_raw - running_median_numpy(power_raw, kernel)
        power = power - np.mean(power)
        SDE = np.max(power / np.std(power))
        scale = SDE / np.max(power)
        power = power * scale
    else:
        power, SDE = power_raw, SDE_raw
    return SR, power_raw, power, SDE_raw, SDE
