"""End to end on the GPU: ``transitleastsquares(t, y, dy).power(**kw)`` (CUDA search through the C
ABI + host post-processing) against the reference's own ``.power()`` results (tests/golden/power_*.npz,
made by oracle/make_golden.py from the unmodified reference).  Tolerance: BASELINE.json's 1e-5
relative on floats; argmax period and transit counts exact."""
import os
import warnings

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _golden(name):
    z = np.load(os.path.join(GOLDEN, "power_%s.npz" % name))
    kw = eval(str(z["kwargs"]), {"__builtins__": {}})
    dy = z["in_dy"] if len(z["in_dy"]) else None
    return z, kw, dy


@pytest.mark.parametrize("name", ["cfg1_50ppm", "small_hetero", "k2_epic201367065", "k2_epic206154641_box", "sentinel",
                                  "k2_epic206154641_grazing", "ref_synthetic", "ref_stats_gap", "ref_uncertainties",
                                  "ref_transit_depth_min"])
def test_power_matches_reference(name):
    from tls_b200 import native, transitleastsquares

    assert native.device_count() > 0
    z, kw, dy = _golden(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = transitleastsquares(z["in_t"], z["in_y"], dy, verbose=False).power(show_progress_bar=False, verbose=False, **kw)
    np.testing.assert_array_equal(res.periods, z["a_periods"])
    np.testing.assert_allclose(res.chi2, z["a_chi2"], rtol=RTOL)
    np.testing.assert_allclose(res.SR, z["a_SR"], rtol=RTOL, atol=1e-9)
    np.testing.assert_allclose(res.power, z["a_power"], rtol=RTOL, atol=1e-6)
    assert np.argmax(res.power) == np.argmax(z["a_power"])
    for key in ("SDE", "SDE_raw", "chi2_min", "chi2red_min", "period", "T0", "duration", "depth", "rp_rs", "snr", "FAP"):
        want, got = float(z["s_" + key]), float(np.asarray(res[key], dtype=float))
        if np.isnan(want):
            assert np.isnan(got), key
        else:
            np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-9, err_msg=key)
    np.testing.assert_allclose(np.asarray(res.transit_times, dtype=float), z["a_transit_times"], rtol=RTOL, equal_nan=True)


def test_reference_known_answers_multi_planet():
    """transitleastsquares/tests/test_multi_planet.py:22-29 (3 decimals; the template comes from
    this repo's limb-darkening model instead of batman)."""
    from tls_b200 import transitleastsquares

    z, kw, dy = _golden("k2_epic201367065")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = transitleastsquares(z["in_t"], z["in_y"], dy, verbose=False).power(show_progress_bar=False, verbose=False)
    np.testing.assert_almost_equal(max(res.power), 45.49085809486116, decimal=3)
    np.testing.assert_almost_equal(max(res.power_raw), 42.93056655774114, decimal=3)
    np.testing.assert_almost_equal(min(res.power), -0.6175100139942546, decimal=3)
    np.testing.assert_almost_equal(min(res.power_raw), -0.3043720539933344, decimal=3)


def _run(name):
    from tls_b200 import transitleastsquares

    z, kw, dy = _golden(name)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return transitleastsquares(z["in_t"], z["in_y"], dy, verbose=False).power(show_progress_bar=False, verbose=False, **kw)


def test_reference_known_answers_shapes():
    """transitleastsquares/tests/test_shapes.py:24-36 (box and grazing templates need no transit model,
    so the reference's 5 decimals hold)."""
    res = _run("k2_epic206154641_box")
    np.testing.assert_almost_equal(res.duration, 0.06111785726416931, decimal=5)
    np.testing.assert_almost_equal(res.rp_rs, 0.08836981203437415, decimal=5)
    res = _run("k2_epic206154641_grazing")
    np.testing.assert_almost_equal(res.duration, 0.08948265482047034, decimal=5)
    np.testing.assert_almost_equal(min(res.chi2red), 0.06759475703796078, decimal=5)


def test_reference_known_answers_stats_gap_uncertainties_synthetic():
    """test_stats_gap.py:57-85, test_uncertainties.py:57, test_synthetic.py:50-63 (light curves made with this
    repo's transit model instead of batman; decimals as in the reference unless the model enters)."""
    res = _run("ref_stats_gap")
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
    res = _run("ref_uncertainties")
    np.testing.assert_almost_equal(res.SDE, 5.292594615900944, decimal=3)
    res = _run("ref_synthetic")
    np.testing.assert_almost_equal(res.period_uncertainty, 0.216212529678387, decimal=5)
    assert res.per_transit_count[0] == 7 and len(res.transit_times) == 3
    np.testing.assert_almost_equal(res.period, 365.2582192473641, decimal=5)
    np.testing.assert_almost_equal(res.transit_times[0], 68.00349264912924, decimal=5)


def test_reference_known_answers_transit_depth_min():
    """test_transit_depth_min.py:50-71: nothing is fitted, every period returns the sentinel N."""
    res = _run("ref_transit_depth_min")
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
