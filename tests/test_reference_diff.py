"""Differential tests of the HOST functions against the unmodified reference, imported from
/root/reference through oracle/ref_shim.py.  They run in the build container only (the reference
tree does not travel to the GPU box: skipped there); what travels are the goldens under
tests/golden/.  Bar: identical return values (bit for bit), identical exception type, identical
warning texts — these functions are the drop-in surface around the GPU search
(transitleastsquares/__init__.py:13-18)."""
import os
import sys
import warnings

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_shim.source_tree_available(), reason="needs /root/reference (build container only)")


@pytest.fixture(scope="module")
def ref():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return ref_shim.load()


def _run(fn, *a, **k):
    try:
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            r = fn(*a, **k)
        return r, [str(x.message) for x in w]
    except Exception as e:  # noqa: BLE001 - the exception type IS the behaviour under test
        return type(e), []


def _same(ra, rb):
    (va, wa), (vb, wb) = ra, rb
    if isinstance(va, type) or isinstance(vb, type):
        return va is vb
    ta = va if isinstance(va, tuple) else (va,)
    tb = vb if isinstance(vb, tuple) else (vb,)
    return wa == wb and len(ta) == len(tb) and all(
        np.array_equal(np.asarray(x, dtype=float), np.asarray(z, dtype=float), equal_nan=True) for x, z in zip(ta, tb))


def _check(name, f_ref, f_mine, *a, **k):
    ra, rb = _run(f_ref, *a, **k), _run(f_mine, *a, **k)
    assert _same(ra, rb), (name, k, str(ra)[:200], str(rb)[:200])


def test_period_grid_and_duration_grid(ref):
    """grid.py:35-131 and grid.py:134-150 over spans, stars, oversampling, limits (also the degenerate ones)."""
    from transitleastsquares import grid as rg

    from tls_b200 import grid as mg

    for span in (5, 27.4, 90, 1500):
        for R, M in ((1, 1), (0.5, 0.5), (2, 1.5), (0.1, 0.1), (3.5, 1)):
            for ov in (1, 3, 5):
                for pmin, pmax in ((0, float("inf")), (1, 10), (0.3, 3), (10, 5), (100, 200)):
                    for ntm in (1, 2, 3):
                        _check("period_grid", rg.period_grid, mg.period_grid, R_star=R, M_star=M, time_span=span,
                               period_min=pmin, period_max=pmax, oversampling_factor=ov, n_transits_min=ntm)
    for span in (27.4, 90, 400):
        per = rg.period_grid(1, 1, span)
        for shortest in (1 / 500, 1 / 4320, 1 / 70000, 0.01):
            for step in (1.02, 1.1, 1.5):
                _check("duration_grid", rg.duration_grid, mg.duration_grid, per, shortest=shortest, log_step=step)


def test_helpers_and_statistics(ref):
    """helpers.cleaned_array (:18-61), resample (:7-15), transit_mask (:64-67), stats.FAP (:8-24), core.fold (:9-12)."""
    from transitleastsquares import core as rc
    from transitleastsquares import helpers as rh
    from transitleastsquares import stats as rs

    from tls_b200 import helpers as mh
    from tls_b200 import stats as ms

    rng = np.random.RandomState(0)
    t = np.linspace(0, 10, 50)
    y = 1 + rng.normal(0, 1e-3, 50)
    dy = np.full(50, 1e-3)
    yo = y.astype(object)
    yo[3], yo[7], yo[9], yo[11] = None, np.nan, np.inf, -1.0
    to = t.astype(object)
    to[4] = None
    dyo = dy.copy()
    dyo[20], dyo[21] = 0, np.nan
    masked = np.ma.masked_invalid(np.where(y > 1, np.nan, t))
    for args in ((t, y), (t, yo), (to, yo), (t, y, dy), (to, yo, dyo), (masked, y), (list(t), list(y))):
        _check("cleaned_array", rh.cleaned_array, mh.cleaned_array, *args)
    for factor in (2.0, 3.0, 1.5, 10.0):
        for dyv in (None, dy):
            _check("resample", rh.resample, mh.resample, t, y, dyv, factor)
    for per, dur, T0 in ((2.0, 0.2, 0.3), (3.3, 0.5, 9.9), (0.7, 0.05, -4.0), (50, 1, 5)):
        _check("transit_mask", rh.transit_mask, mh.transit_mask, t, per, dur, T0)
        _check("fold", rc.fold, ms.fold, t, per, T0)
    for sde in (0, 3.0, 6.9, 7.0, 7.5, 8.3, 9.1, 12.0, 20.0, 100.0, np.nan):
        _check("FAP", rs.FAP, ms.FAP, sde)


KWARGS = [
    {}, {"use_threads": 0}, {"use_threads": "1"}, {"use_threads": 2.5}, {"period_min": 5, "period_max": 2},
    {"period_min": -1}, {"R_star": -1}, {"M_star": 0}, {"R_star_min": 2, "R_star_max": 1},
    {"M_star_min": 2, "M_star_max": 1}, {"R_star": 5}, {"M_star": 5}, {"transit_template": "box"},
    {"transit_template": "grazing"}, {"transit_template": "foo"}, {"transit_template": 3}, {"n_transits_min": 1},
    {"n_transits_min": 0}, {"n_transits_min": 2.5}, {"n_transits_min": "2"}, {"T0_fit_margin": 0.5},
    {"T0_fit_margin": -1}, {"T0_fit_margin": 0}, {"oversampling_factor": 0}, {"oversampling_factor": 2.5},
    {"oversampling_factor": 7}, {"duration_grid_step": 1.0}, {"duration_grid_step": 0.9}, {"duration_grid_step": 1.3},
    {"transit_depth_min": 0}, {"transit_depth_min": -1e-6}, {"limb_dark": "linear", "u": [0.5]},
    {"limb_dark": "nonlinear", "u": [0.1, 0.2, 0.3, 0.4]}, {"u": [0.4, 0.4]}, {"per": 5}, {"rp": 0.2}, {"a": 20},
    {"show_progress_bar": False}, {"verbose": False}, {"period_max": 100}, {"period_min": 0.1},
]


def test_validate_args_and_inputs(ref):
    """validate.py:49-181: every keyword, its default, its clamp and its ValueError; validate.py:9-46 on good
    and malformed arrays."""
    from transitleastsquares import validate as rv

    from tls_b200 import validate as mv

    rng = np.random.RandomState(1)
    t = np.linspace(0, 30, 800)
    y = 1 + rng.normal(0, 1e-4, 800)

    class Bag(object):
        pass

    def attrs(validate_args, kw):
        b = Bag()
        b.t, b.y, b.dy = t, y, np.full(800, np.std(y))
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                out, _ = validate_args(b, dict(kw))
        except Exception as e:  # noqa: BLE001
            return type(e)
        return {k: v for k, v in vars(out).items() if k not in ("t", "y", "dy")}

    for kw in KWARGS:
        a, b = attrs(rv.validate_args, kw), attrs(mv.validate_args, kw)
        if isinstance(a, type) or isinstance(b, type):
            assert a is b, (kw, a, b)
        else:
            assert set(a) == set(b), (kw, set(a) ^ set(b))
            for k in a:
                assert np.all(np.asarray(a[k] == b[k])), (kw, k, a[k], b[k])

    tn, yn = t.copy(), y.copy()
    yn[5], tn[9] = np.nan, np.inf
    good = [(t, y, None), (t, y, np.full(800, 2.0)), (tn, yn, None), (t[::-1], y, None), (list(t), list(y), None),
            (np.r_[t[:5], t[3], t[5:-1]], y, None)]
    for args in good:
        _check("validate_inputs", rv.validate_inputs, mv.validate_inputs, *args)
    # malformed: both must refuse (the reference trips over an IndexError inside its cleaner for unequal
    # lengths before reaching its own size check, validate.py:41-42; here that check is what fires)
    bad = [(t[:-1], y, None), (t, y, np.ones(799)), (t, -y, None), (t, y, -np.ones(800)), (t, y * 0, None),
           (t, y + np.inf, None)]
    for args in bad:
        ra, rb = _run(rv.validate_inputs, *args)[0], _run(mv.validate_inputs, *args)[0]
        assert isinstance(ra, type) and isinstance(rb, type), (ra, rb)


def test_template_bank_is_bit_identical(ref):
    """transit.py:98-160 (get_cache: reference transit -> slice / lerp / trim -> overview) with the same
    transit model underneath (the stand-in for batman): every template and every overview field equal."""
    from transitleastsquares import grid as rg
    from transitleastsquares import transit as rt

    from tls_b200 import transit as mt

    per = rg.period_grid(1, 1, 90.0)
    planets = ((13.4, 0.103, 23.1, 89.21), (2.5, 2 ** 0.5, 3, 75.0))  # default, grazing (tls_constants.py:40-66)
    for step, n in ((1.1, 4320), (1.02, 19440)):
        dur = rg.duration_grid(per, shortest=1 / n, log_step=step)
        mw = int(np.max(dur) * n)
        mw += mw % 2
        for law, u, p in (("quadratic", [0.4804, 0.1867], planets[0]), ("linear", [0.5], planets[0]),
                          ("nonlinear", [0.1, 0.2, 0.3, 0.1], planets[0]), ("quadratic", [0.4804, 0.1867], planets[1])):
            kw = dict(durations=dur, maxwidth_in_samples=mw, per=p[0], rp=p[1], a=p[2], inc=p[3], ecc=0, w=90, u=u,
                      limb_dark=law, verbose=False)
            oa, la = rt.get_cache(**kw)
            ob, lb = mt.get_cache(**kw)
            assert oa.dtype == ob.dtype and len(la) == len(lb)
            for f in oa.dtype.names:
                np.testing.assert_array_equal(oa[f], ob[f])
            for x, z in zip(la, lb):
                np.testing.assert_array_equal(x, z)


def test_oracle_against_the_numba_search_on_random_inputs(ref):
    """Beyond the committed goldens: the C oracle (oracle/tls_oracle.c) against the reference's own numba
    ``core.search_period`` (core.py:96-188) on random light curves — irregular and unsorted time stamps, ties,
    per-point dy, every T0_fit_margin regime, several gates — rows exact, chi2 / depth to 1e-9."""
    from transitleastsquares.core import search_period

    from oracle import oracle
    from tls_b200 import transitleastsquares as mine

    rng = np.random.RandomState(11)
    checked = 0
    for case in range(30):
        n = int(rng.choice([60, 150, 400, 900]))
        span = float(rng.choice([8.0, 27.0, 90.0]))
        t = np.sort(rng.uniform(0.5, 0.5 + span, n)) if case % 2 else np.linspace(0.5, 0.5 + span, n)
        if case % 5 == 0:
            t = np.round(t, 1)  # ties
        if case % 4 == 3:
            t = t[rng.permutation(n)]  # unsorted
        ppm = float(rng.choice([50e-6, 500e-6, 3e-3]))
        y = 1 + rng.normal(0, ppm, n)
        per, dur = rng.uniform(1.0, span / 3), rng.uniform(0.05, 0.3)
        y[np.abs((t - 0.7) % per) < dur] -= rng.uniform(2, 8) * ppm
        dy = None if case % 3 else ppm * rng.uniform(0.5, 2.0, n)
        kw = dict(T0_fit_margin=float(rng.choice([0.0, 0.01, 0.1])), transit_depth_min=float(rng.choice([10e-6, 1e-4, 2e-3])),
                  oversampling_factor=int(rng.choice([1, 3])), duration_grid_step=float(rng.choice([1.1, 1.3])))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                inp = mine(t, y, dy, verbose=False).prepare(verbose=False, **kw)
            except ValueError:
                continue
        periods = inp.periods[np.linspace(0, len(inp.periods) - 1, min(25, len(inp.periods))).astype(int)]
        chi2, row, depth = oracle.search_periods_c(inp.t, inp.y, inp.dy, periods, inp.templates, inp.params)
        for k, p in enumerate(periods):
            out = search_period(p, inp.t, inp.y, inp.dy, lc_arr=inp.lc_arr, lc_cache_overview=inp.overview, **inp.params)
            assert int(out[2]) == int(row[k]), (case, p, out, chi2[k], row[k])
            if np.isfinite(out[1]):
                np.testing.assert_allclose(chi2[k], out[1], rtol=1e-9)
                np.testing.assert_allclose(depth[k], out[3], rtol=1e-9, atol=1e-15)
            else:
                assert chi2[k] == out[1]
            checked += 1
    assert checked >= 400


def test_oracle_spectra_and_t0_fit_against_the_reference_on_random_inputs(ref):
    """oracle.spectra_numpy vs stats.spectra (stats.py:105-132) and oracle.final_T0_fit_numpy vs
    stats.final_T0_fit (stats.py:135-204) beyond the goldens: short and long chi2 rows (with and without the
    median detrending), inf entries, every oversampling; T0 of random templates, margins and periods."""
    from transitleastsquares import stats as rs

    from oracle import oracle

    rng = np.random.RandomState(5)
    for P, ov in ((40, 1), (181, 1), (182, 1), (183, 1), (600, 2), (1500, 3), (3000, 5), (547, 3)):
        chi2 = 1000 - rng.rand(P) * 30
        chi2[rng.randint(P)] -= 200
        if P % 2:
            chi2[rng.randint(P)] = np.inf
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = rs.spectra(chi2.copy(), ov)
            got = oracle.spectra_numpy(chi2.copy(), ov)
        for a, b in zip(want, got):
            np.testing.assert_allclose(np.asarray(b, dtype=float), np.asarray(a, dtype=float), rtol=1e-12, atol=1e-12, equal_nan=True)
    for case in range(10):
        n = int(rng.choice([200, 500, 1200]))
        t = np.sort(rng.uniform(1.0, 40.0, n)) if case % 2 else np.linspace(1.0, 40.0, n)
        y = 1 + rng.normal(0, 2e-4, n)
        period = float(rng.uniform(1.5, 12.0))
        L = int(rng.randint(5, 40))
        signal = 1 - 0.5 * np.sin(np.linspace(0, np.pi, L)) ** 2 * 1e-3
        depth = 1 - rng.uniform(1e-4, 2e-3)
        y[np.abs((t - 2.2) % period) < 0.15] -= 1e-3
        margin = float(rng.choice([0.0, 0.01, 0.1, 1.0]))
        dy = np.full(n, np.std(y))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = rs.final_T0_fit(signal=signal.copy(), depth=depth, t=t, y=y, dy=dy.copy(), period=period,
                                   T0_fit_margin=margin, show_progress_bar=False, verbose=False)
            got = oracle.final_T0_fit_numpy(signal.copy(), depth, t, y, dy.copy(), period, margin)[0]
        assert got == want, (case, got, want)


def test_oracle_against_the_numba_search_on_the_gpu_fuzz_cases(ref):
    """The randomised GPU test (tests/test_gpu_fuzz.py) compares the CUDA search with the C oracle on the GPU box, where
    the reference does not exist.  Here, where it does, the SAME light curves and periods go through the reference's own
    numba ``core.search_period`` (core.py:96-188) and the oracle: rows exact (ties in value excepted, as in the GPU test),
    chi2 / depth to 1e-9.  Together the two tests tie the CUDA results on those inputs to the unmodified reference."""
    from transitleastsquares.core import search_period

    import test_gpu_fuzz
    from oracle import oracle

    checked = 0
    for seed in range(28):  # the generator is deterministic per seed
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            inp, periods = test_gpu_fuzz._case(seed)
        periods = periods[:: max(1, len(periods) // 12)]
        chi2, row, depth = oracle.search_periods_c(inp.t, inp.y, inp.dy, periods, inp.templates, inp.params)
        for k, p in enumerate(periods):
            out = search_period(p, inp.t, inp.y, inp.dy, lc_arr=inp.lc_arr, lc_cache_overview=inp.overview, **inp.params)
            if np.isfinite(out[1]):
                np.testing.assert_allclose(chi2[k], out[1], rtol=1e-9, err_msg="seed %d period %r" % (seed, p))
                tie = abs(chi2[k] - out[1]) <= 1e-13 * abs(out[1])
                assert int(out[2]) == int(row[k]) or tie, (seed, p, out, chi2[k], row[k])
                if int(out[2]) == int(row[k]):
                    np.testing.assert_allclose(depth[k], out[3], rtol=1e-9, atol=1e-15)
            else:
                assert chi2[k] == out[1] and int(out[2]) == int(row[k])
            checked += 1
    assert checked >= 300
