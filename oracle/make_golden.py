"""TEST INFRASTRUCTURE — generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py            # all cases
    python oracle/make_golden.py cfg1_50ppm # one case

Each ``search_*.npz`` holds the exact inputs handed to the reference's
``core.search_period`` (core.py:96-188, numba path, via ``oracle/ref_shim.py``)
and its outputs (chi2, row, depth) for a list of periods.  Each ``power_*.npz``
holds inputs and the full results of the reference's ``.power()`` (main.py:51).
The K2 light curves come from the reference's own test fixtures
(``transitleastsquares/tests/EPIC*.csv``) with the preprocessing of the test
that uses them, so the reference's known answers apply to them.
"""
from __future__ import annotations

import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
OUT = os.path.join(REPO, "tests", "golden")

from oracle import ref_shim  # noqa: E402

ref = ref_shim.load()
from transitleastsquares.core import search_period  # noqa: E402
from transitleastsquares.transit import get_cache  # noqa: E402

from tls_b200 import transitleastsquares as mine  # noqa: E402
from tls_b200 import workloads  # noqa: E402
from tls_b200.transit import pack_templates  # noqa: E402

REF_TESTS = os.path.join(ref_shim.REFERENCE_ROOT, "transitleastsquares", "tests")


def k2_multi_planet():
    """tests/test_multi_planet.py:15-20"""
    import scipy.signal

    t, y = np.loadtxt(os.path.join(REF_TESTS, "EPIC201367065.csv"), delimiter=",", unpack=True)
    trend = scipy.signal.medfilt(y, 25)
    return t, y / trend, None


def k2_shapes():
    """tests/test_shapes.py:10-15"""
    import scipy.signal

    t, y = np.loadtxt(os.path.join(REF_TESTS, "EPIC206154641.csv"), delimiter=",", unpack=True)
    trend = scipy.signal.medfilt(y, 25)
    return t, y / trend, None


def prepared(t, y, dy, **kw):
    model = mine(t, y, dy, verbose=False)
    return model.prepare(verbose=False, **kw)


def run_reference_search(inp, periods, lc_arr=None, overview=None):
    lc_arr = inp.lc_arr if lc_arr is None else lc_arr
    overview = inp.overview if overview is None else overview
    chi2 = np.empty(len(periods))
    row = np.empty(len(periods), np.int64)
    depth = np.empty(len(periods))
    for k, p in enumerate(periods):
        out = search_period(p, inp.t, inp.y, inp.dy, lc_arr=lc_arr, lc_cache_overview=overview,
                            **inp.params)
        chi2[k], row[k], depth[k] = out[1], out[2], out[3]
    return chi2, row, depth


def save_search(name, inp, periods, note, lc_arr=None, overview=None):
    t0 = time.time()
    lc_arr = inp.lc_arr if lc_arr is None else lc_arr
    overview = inp.overview if overview is None else overview
    # cross-check that the template bank equals the reference's own get_cache where applicable
    chi2, row, depth = run_reference_search(inp, periods, lc_arr, overview)
    tp = pack_templates(lc_arr, overview)
    np.savez_compressed(
        os.path.join(OUT, "search_%s.npz" % name),
        t=inp.t, y=inp.y, dy=inp.dy, periods=np.asarray(periods, float),
        tp_signal=tp["signal"], tp_offset=tp["offset"], tp_length=tp["length"],
        tp_width=tp["width"], tp_overshoot=tp["overshoot"],
        params=np.array([inp.params[k] for k in PARAM_ORDER]),
        chi2=chi2, row=row, depth=depth, note=np.array(note),
    )
    n_inf = int(np.sum(~np.isfinite(chi2)))
    n_sent = int(np.sum(chi2 == len(inp.y)))
    print("%-22s N=%6d P=%6d R=%3d  inf=%d sentinel=%d  (%.1fs)" % (
        name, len(inp.y), len(periods), len(lc_arr), n_inf, n_sent, time.time() - t0))


PARAM_ORDER = ("transit_depth_min", "R_star_min", "R_star_max", "M_star_min", "M_star_max", "T0_fit_margin")


def case_cfg1_50ppm():
    t, y, dy, kw = workloads.lightcurve("cfg1")
    inp = prepared(t, y, dy, **kw)
    # the template bank must equal what the reference's get_cache builds from the same model
    M = int(np.max(inp.durations) * len(inp.y))
    M += M % 2
    ov, lc = get_cache(inp.durations, M, per=12.9, rp=0.03, a=23.1, inc=89.21, ecc=0, w=90,
                       u=[0.4804, 0.1867], limb_dark="quadratic", verbose=False)
    assert np.array_equal(ov, inp.overview) and all(np.array_equal(a, b) for a, b in zip(lc, inp.lc_arr))
    save_search("cfg1_50ppm", inp, inp.periods, "cfg-1: 90 d @ 30 min, 50 ppm, dy=None, full default grid")


def case_cfg1_500ppm():
    t, y, dy, kw = workloads.lightcurve("cfg1_500ppm")
    inp = prepared(t, y, dy, **kw)
    save_search("cfg1_500ppm", inp, inp.periods[::4], "cfg-1 shape at 500 ppm, every 4th period")


def case_cfg1_hetero():
    t, y, dy, kw = workloads.lightcurve("cfg1_500ppm", hetero=True)
    inp = prepared(t, y, dy, **kw)
    save_search("cfg1_hetero", inp, inp.periods[1::4], "cfg-1 shape, 500 ppm, dy = sigma*U(0.5,2) normalised; every 4th period")


def case_sentinel():
    t, y, dy, kw = workloads.lightcurve("cfg1")
    inp = prepared(t, y, dy, transit_depth_min=1000e-6, **kw)
    save_search("sentinel", inp, inp.periods[::16], "transit_depth_min=1000 ppm: every period returns the N sentinel")


def case_margins():
    t, y, dy, kw = workloads.lightcurve("cfg1_500ppm")
    for tag, margin in (("margin0", 0.0), ("margin003", 0.03), ("margin01", 0.2)):
        inp = prepared(t, y, dy, T0_fit_margin=margin, **kw)
        save_search(tag, inp, inp.periods[5::97], "T0_fit_margin=%g (clamped to [0,0.1])" % margin)


def case_small_and_ties():
    t, y, dy, kw = workloads.lightcurve("small")
    inp = prepared(t, y, dy, **kw)
    save_search("small", inp, inp.periods[::3], "N=720 quick case")
    # unsorted time stamps with exact duplicates => equal phases => stable-sort order matters
    rng = np.random.RandomState(7)
    t2 = np.concatenate([t, t[100:160]])
    y2 = np.concatenate([y, y[100:160] + rng.normal(0, 2e-4, 60)])
    perm = rng.permutation(len(t2))
    inp2 = prepared(t2[perm], y2[perm], None, **kw)
    save_search("ties_unsorted", inp2, inp2.periods[::7], "shuffled time stamps with 60 exact duplicates (phase ties)")
    # tiny
    t3 = np.linspace(1.0, 21.0, 240)
    y3 = 1 + np.random.RandomState(3).normal(0, 1e-3, 240)
    for c in (30, 90, 150, 210):
        y3[c:c + 3] -= 4e-3
    inp3 = prepared(t3, y3, None)
    save_search("tiny", inp3, inp3.periods[::5], "N=240")


def case_ragged_templates():
    """L < W rows: shorten some templates by hand."""
    t, y, dy, kw = workloads.lightcurve("cfg1_500ppm")
    inp = prepared(t, y, dy, **kw)
    lc = np.empty(len(inp.lc_arr), dtype=object)
    for r, s in enumerate(inp.lc_arr):
        cut = (r % 3)
        lc[r] = s[: len(s) - cut].copy() if len(s) - cut >= 3 else s.copy()
    save_search("ragged_L", inp, inp.periods[3::61], "templates trimmed so that L = W - (row %% 3)", lc_arr=lc)


def case_no_admissible():
    t, y, dy, kw = workloads.lightcurve("cfg1_500ppm")
    inp = prepared(t, y, dy, R_star_max=1.0, M_star_max=1.0, R_star_min=0.9, M_star_min=1.0, **kw)
    # narrow stellar limits: few widths admissible; then a bank holding only wide rows => none admissible
    save_search("narrow_limits", inp, inp.periods[::53], "R_star_min=0.9,R_star_max=1,M_star_min=M_star_max=1")
    wide = np.where(inp.overview["width_in_samples"] >= 200)[0]
    lc = np.empty(len(wide), dtype=object)
    for k, r in enumerate(wide):
        lc[k] = inp.lc_arr[r]
    save_search("no_admissible", inp, inp.periods[::211], "bank with only W>=200 rows: short periods have no admissible width (inf)",
                lc_arr=lc, overview=inp.overview[wide])


def case_cfg3():
    t, y, dy, kw = workloads.lightcurve("cfg3")
    inp = prepared(t, y, dy, **kw)
    sel = np.linspace(0, len(inp.periods) - 1, 48).astype(int)
    save_search("cfg3", inp, inp.periods[sel], "cfg-3: 27 d @ 2 min, 500 ppm, duration_grid_step=1.02; 48 periods of %d" % len(inp.periods))


def case_cfg2():
    t, y, dy, kw = workloads.lightcurve("cfg2")
    inp = prepared(t, y, dy, **kw)
    sel = list(np.linspace(0, len(inp.periods) - 1, 24).astype(int))
    for harmonic in (10.123, 20.246, 5.0615, 30.369, 3.3743):  # around the injected period and aliases
        k = int(np.argmin(np.abs(inp.periods - harmonic)))
        sel += [k - 2, k - 1, k, k + 1, k + 2]
    sel = np.unique(np.clip(sel, 0, len(inp.periods) - 1))
    save_search("cfg2", inp, inp.periods[sel], "cfg-2: 4 yr @ 30 min, 50 ppm; 24 spread periods + 25 around the injected period and its aliases, of %d" % len(inp.periods))


def case_k2():
    t, y, dy = k2_multi_planet()
    inp = prepared(t, y, dy)
    save_search("k2_epic201367065", inp, inp.periods[::9], "EPIC 201367065 / medfilt(25) (tests/test_multi_planet.py), every 9th period")
    t, y, dy = k2_shapes()
    for shape in ("box", "grazing"):
        inp = prepared(t, y, dy, transit_template=shape)
        save_search("k2_epic206154641_" + shape, inp, inp.periods[::17], "EPIC 206154641, transit_template=%s (tests/test_shapes.py)" % shape)


# ------------------------------------------------------------------ power() goldens
SCALARS = ("SDE", "SDE_raw", "chi2_min", "chi2red_min", "period", "period_uncertainty", "T0",
           "duration", "depth", "rp_rs", "snr", "odd_even_mismatch", "transit_count",
           "distinct_transit_count", "empty_transit_count", "FAP", "in_transit_count",
           "after_transit_count", "before_transit_count")
ARRAYS = ("periods", "power", "power_raw", "SR", "chi2", "chi2red", "transit_times",
          "per_transit_count", "transit_depths", "transit_depths_uncertainties", "snr_per_transit",
          "snr_pink_per_transit", "model_lightcurve_time", "model_lightcurve_model",
          "model_folded_phase", "folded_y", "folded_dy", "folded_phase", "model_folded_model",
          "depth_mean", "depth_mean_even", "depth_mean_odd")


def save_power(name, t, y, dy, note, **kw):
    t0 = time.time()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = ref.transitleastsquares(t, y, dy, verbose=False)
        res = model.power(use_threads=1, show_progress_bar=False, verbose=False, **kw)
    blob = dict(in_t=np.asarray(t, float), in_y=np.asarray(y, float),
                in_dy=np.zeros(0) if dy is None else np.asarray(dy, float),
                kwargs=np.array(repr(kw)), note=np.array(note))
    for k in SCALARS:
        blob["s_" + k] = np.asarray(res[k], dtype=float)
    for k in ARRAYS:
        blob["a_" + k] = np.asarray(res[k], dtype=float)
    np.savez_compressed(os.path.join(OUT, "power_%s.npz" % name), **blob)
    print("%-22s power(): period=%.6f SDE=%.4f depth=%.6f  (%.1fs)" % (
        name, res.period, res.SDE, res.depth, time.time() - t0))


def case_power():
    t, y, dy, kw = workloads.lightcurve("cfg1")
    save_power("cfg1_50ppm", t, y, dy, "cfg-1 end to end", **kw)
    t, y, dy, kw = workloads.lightcurve("small", hetero=True)
    save_power("small_hetero", t, y, dy, "N=720 with per-point dy", **kw)
    t, y, dy = k2_multi_planet()
    save_power("k2_epic201367065", t, y, dy, "tests/test_multi_planet.py first run")
    t, y, dy = k2_shapes()
    save_power("k2_epic206154641_box", t, y, dy, "tests/test_shapes.py box", transit_template="box")
    t, y, dy, kw = workloads.lightcurve("cfg1")
    save_power("sentinel", t, y, dy, "no-fit branch", transit_depth_min=1000e-6, T0_fit_margin=0.1)


# ------------------------------------------------------------------ the reference's own test scripts, end to end
def three_year_curve(seed=0):
    """The synthetic light curve shared by tests/test_synthetic.py:9-38, test_stats_gap.py:11-37,
    test_transit_depth_min.py:12-38 and test_uncertainties.py:9-38: 3 yr at 12 samples per day, an Earth
    around a Sun at 365.25 d (linear limb darkening), 5 ppm white noise.  The transit model is this repo's
    stand-in for batman (the reference tests call batman at this point)."""
    from tls_b200 import limbdark

    np.random.seed(seed=seed)
    start, days, samples_per_day = 48, 365.25 * 3, 12
    samples = int(days * samples_per_day)
    t = np.linspace(start, start + days, samples)
    ma = limbdark.TransitParams()
    ma.t0 = start + 20
    ma.per = 365.25
    ma.rp = 6371 / 696342
    ma.a = 217
    ma.inc = 90
    ma.ecc = 0
    ma.w = 90
    ma.u = [0.5]
    ma.limb_dark = "linear"
    flux = limbdark.TransitModel(ma, t).light_curve(ma)
    stdev = 10 ** -6 * 5
    y = flux + np.random.normal(0, stdev, int(samples))
    return t, y, stdev


def case_ref_tests():
    # tests/test_synthetic.py:38-48
    t, y, stdev = three_year_curve()
    y[1] = np.nan
    save_power("ref_synthetic", t, y, None, "tests/test_synthetic.py", period_min=360, period_max=370,
               transit_depth_min=10 * 10 ** -6, oversampling_factor=5, duration_grid_step=1.02)
    # tests/test_stats_gap.py:38-56 (a gap of NaNs in t and y)
    t, y, stdev = three_year_curve()
    y[1] = np.nan
    y[200:500] = np.nan
    t[200:500] = np.nan
    save_power("ref_stats_gap", t, y, None, "tests/test_stats_gap.py", period_min=360, period_max=370,
               transit_depth_min=10 * 10 ** -6, oversampling_factor=2, duration_grid_step=1.1, T0_fit_margin=1.2)
    # tests/test_uncertainties.py:40-56 (excess noise with matching dy at the end of the series)
    t, y, stdev = three_year_curve()
    y[10000:] = y[10000:] + np.random.normal(0, 10 * stdev, 3149)
    dy = np.full(len(y), stdev)
    dy[10000:] = 10 * stdev
    save_power("ref_uncertainties", t, y, dy, "tests/test_uncertainties.py", period_min=360, period_max=370,
               oversampling_factor=3, duration_grid_step=1.05, T0_fit_margin=0.2)
    # tests/test_transit_depth_min.py:39-48 (nothing is fitted)
    t, y, stdev = three_year_curve()
    y[1] = np.nan
    save_power("ref_transit_depth_min", t, y, None, "tests/test_transit_depth_min.py",
               transit_depth_min=1000 * 10 ** -6, period_min=360, period_max=370, oversampling_factor=5,
               duration_grid_step=1.02, T0_fit_margin=0.1)
    # tests/test_shapes.py:29-36
    t, y, dy = k2_shapes()
    save_power("k2_epic206154641_grazing", t, y, dy, "tests/test_shapes.py grazing", transit_template="grazing")


# ------------------------------------------------------------------ final_T0_fit goldens
def save_t0fit(name, t, y, signal, depth, period, margin, note):
    """Reference stats.final_T0_fit (stats.py:135-204, unmodified) -> T0; the per-trial residuals
    come from the oracle's restatement, which must pick the same T0."""
    from transitleastsquares.stats import final_T0_fit as ref_fit

    from oracle import oracle

    t0 = time.time()
    dy = np.full(len(y), np.std(y))
    T0_ref = ref_fit(signal=np.array(signal, float), depth=depth, t=t, y=y, dy=dy.copy(), period=period,
                     T0_fit_margin=margin, show_progress_bar=False, verbose=False)
    T0_orc, resid, trials = oracle.final_T0_fit_numpy(signal, depth, t, y, dy, period, margin)
    assert T0_ref == T0_orc, (name, T0_ref, T0_orc)
    np.savez_compressed(os.path.join(OUT, "t0fit_%s.npz" % name), t=t, y=y, signal=np.asarray(signal, float),
                        depth=np.float64(depth), period=np.float64(period), margin=np.float64(margin),
                        T0=np.float64(T0_ref), residuals=resid, trials=trials, note=np.array(note))
    print("%-22s N=%6d dur=%4d trials=%6d T0=%.6f  (%.1fs)" % (name, len(y), len(signal), len(trials), T0_ref, time.time() - t0))


def _best_of(inp, stride=1):
    """Best (signal, depth, period) of a search over every `stride`-th period (oracle C port)."""
    from oracle import oracle

    per = inp.periods[::stride]
    chi2, row, depth = oracle.search_periods_c(inp.t, inp.y, inp.dy, per, inp.templates, inp.params)
    k = int(np.argmin(chi2))
    return inp.lc_arr[int(row[k])], float(depth[k]), float(per[k])


def case_t0fit():
    t, y, dy, kw = workloads.lightcurve("cfg1")
    inp = prepared(t, y, dy, **kw)
    sig, depth, period = _best_of(inp)
    save_t0fit("cfg1", inp.t, inp.y, sig, depth, period, 0.01, "cfg-1 best model, default margin (points = N)")
    save_t0fit("cfg1_margin1", inp.t, inp.y, sig, depth, period, 1.0, "cfg-1, T0_fit_margin=1 (coarse scan)")
    t, y, dy, kw = workloads.lightcurve("small", hetero=True)
    inp = prepared(t, y, dy, **kw)
    sig, depth, period = _best_of(inp)
    save_t0fit("small_margin0", inp.t, inp.y, sig, depth, period, 0.0, "N=720, T0_fit_margin=0 (every sample)")
    # repeated and unsorted time stamps: the stable order of equal phases matters
    rng = np.random.RandomState(5)
    tt = np.round(inp.t, 1)
    perm = rng.permutation(len(tt))
    save_t0fit("ties_unsorted", tt[perm], inp.y[perm], sig, depth, 2.0, 0.01, "time stamps rounded to 0.1 d and shuffled, period 2 d")
    t, y, dy = k2_multi_planet()
    inp = prepared(t, y, dy)
    sig, depth, period = _best_of(inp, 9)
    save_t0fit("k2_epic201367065", inp.t, inp.y, sig, depth, period, 0.01, "EPIC 201367065 (irregular sampling)")
    t, y, dy, kw = workloads.lightcurve("cfg3")
    inp = prepared(t, y, dy, **kw)
    sig, depth, period = _best_of(inp, 20)
    save_t0fit("cfg3_margin02", inp.t, inp.y, sig, depth, period, 0.2, "cfg-3 (N=19440), margin 0.2")
    t, y, dy, kw = workloads.lightcurve("cfg2")
    inp = prepared(t, y, dy, **kw)
    row = int(np.argmin(np.abs(inp.overview["width_in_samples"] - 24)))
    save_t0fit("cfg2_margin1", inp.t, inp.y, inp.lc_arr[row], 1 - 8.4e-5, 10.123, 1.0, "cfg-2 (N=70128, streaming path), injected period, margin 1")


CASES = {
    "cfg1_50ppm": case_cfg1_50ppm, "cfg1_500ppm": case_cfg1_500ppm, "cfg1_hetero": case_cfg1_hetero,
    "sentinel": case_sentinel, "margins": case_margins, "small": case_small_and_ties,
    "ragged": case_ragged_templates, "no_admissible": case_no_admissible, "cfg3": case_cfg3,
    "cfg2": case_cfg2, "k2": case_k2, "power": case_power, "ref_tests": case_ref_tests, "t0fit": case_t0fit,
}

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    warnings.simplefilter("ignore")
    todo = sys.argv[1:] or list(CASES)
    for name in todo:
        CASES[name]()
