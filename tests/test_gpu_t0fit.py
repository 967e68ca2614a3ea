"""final_T0_fit on the GPU (``tlsb_final_t0_fit`` / ``tlsb_final_t0_fit_lc`` through the C ABI)
against the reference's own ``stats.final_T0_fit`` (stats.py:135-204): T0 must be the SAME trial
epoch (bit-exact choice), per-trial residuals within 1e-9 relative of the oracle's restatement
(the product bar of BASELINE.json is 1e-5)."""
import numpy as np
import pytest

from conftest import load_t0fit_golden, t0fit_goldens
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", t0fit_goldens())
def test_t0_fit_matches_reference(name):
    from tls_b200 import native, stats

    g = load_t0fit_golden(name)
    dy = np.full(len(g["y"]), np.std(g["y"]))
    model_in, trials = stats.t0_fit_inputs(g["signal"], g["depth"], g["t"], g["y"], g["period"], g["margin"])
    np.testing.assert_array_equal(trials, g["trials"])
    best, resid = native.final_t0_fit(g["t"], g["y"], dy, model_in, g["period"], trials)
    np.testing.assert_allclose(resid, g["residuals"], rtol=1e-9, atol=0)
    assert best == int(np.argmin(g["residuals"]))
    assert trials[best] == g["T0"]
    # the drop-in function itself
    T0 = stats.final_T0_fit(g["signal"], g["depth"], g["t"], g["y"], dy, g["period"], g["margin"], False, False)
    assert T0 == g["T0"]


def test_handle_api_equals_one_shot_and_live_oracle():
    from tls_b200 import native, stats

    g = load_t0fit_golden("small_margin0")
    dy = np.full(len(g["y"]), np.std(g["y"]))
    rng = np.random.RandomState(11)
    s = native.Searcher()
    s.set_lightcurve(g["t"], g["y"], dy)
    for period in (1.7, 3.3, 9.1):
        depth = 1 - rng.uniform(1e-4, 5e-3)
        model_in, trials = stats.t0_fit_inputs(g["signal"], depth, g["t"], g["y"], period, 0.05)
        b1, r1 = s.final_t0_fit(model_in, period, trials)
        b2, r2 = native.final_t0_fit(g["t"], g["y"], dy, model_in, period, trials)
        np.testing.assert_array_equal(r1, r2)
        assert b1 == b2
        T0, want, _ = oracle.final_T0_fit_numpy(g["signal"], depth, g["t"], g["y"], dy, period, 0.05)
        np.testing.assert_allclose(r1, want, rtol=1e-9)
        assert trials[b1] == T0
    assert s.t0_fit_ms > 0
    s.close()


def test_t0_fit_after_search_on_the_same_handle():
    """The multi-planet / power() sequence: search, then T0 fit, on one handle; the search results
    are not disturbed and a second search still works."""
    from conftest import load_search_golden
    from tls_b200 import native, stats

    g = load_search_golden("small")
    s = native.Searcher()
    s.set_inputs(g["t"], g["y"], g["dy"], g["templates"], g["params"])
    s.set_periods(g["periods"])
    s.search_async()
    chi2, row, depth, _ = s.results()
    k = int(np.argmin(chi2))
    off, ln = int(g["templates"]["offset"][row[k]]), int(g["templates"]["length"][row[k]])
    signal = g["templates"]["signal"][off:off + ln]
    model_in, trials = stats.t0_fit_inputs(signal, depth[k], g["t"], g["y"], g["periods"][k], 0.01)
    best, resid = s.final_t0_fit(model_in, g["periods"][k], trials)
    T0, want, _ = oracle.final_T0_fit_numpy(signal, depth[k], g["t"], g["y"], g["dy"], g["periods"][k], 0.01)
    assert trials[best] == T0
    np.testing.assert_allclose(resid, want, rtol=1e-9)
    s.search_async()
    chi2b, rowb, depthb, _ = s.results()
    np.testing.assert_array_equal(chi2, chi2b)
    np.testing.assert_array_equal(row, rowb)
    s.close()


def test_t0_fit_argument_errors():
    from tls_b200 import native

    g = load_t0fit_golden("small_margin0")
    dy = np.full(len(g["y"]), 1e-3)
    with pytest.raises(RuntimeError, match="dur"):
        native.final_t0_fit(g["t"], g["y"], dy, np.ones(len(g["y"]) + 1), 2.0, g["trials"])
    with pytest.raises(RuntimeError, match="period"):
        native.final_t0_fit(g["t"], g["y"], dy, np.ones(8), -1.0, g["trials"])
    s = native.Searcher()
    with pytest.raises(RuntimeError, match="light curve"):
        s.final_t0_fit(np.ones(8), 2.0, g["trials"])
    s.close()
