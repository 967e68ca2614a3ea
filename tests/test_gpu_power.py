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


@pytest.mark.parametrize("name", ["cfg1_50ppm", "small_hetero", "k2_epic201367065", "k2_epic206154641_box", "sentinel"])
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
