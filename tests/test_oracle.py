"""The oracle (oracle/tls_oracle.c + the numpy restatement) against golden vectors
produced by the reference's own numba path (oracle/make_golden.py)."""
import numpy as np
import pytest

from conftest import assert_search_parity, load_search_golden, load_t0fit_golden, search_goldens, t0fit_goldens
from oracle import oracle


@pytest.mark.parametrize("name", search_goldens())
def test_c_oracle_matches_reference(name):
    g = load_search_golden(name)
    got = oracle.search_periods_c(g["t"], g["y"], g["dy"], g["periods"], g["templates"], g["params"])
    # the oracle is held to a much tighter bar than the product: 1e-9 relative
    assert_search_parity(got, g, rtol=1e-9, label=name)


@pytest.mark.parametrize("name", ["tiny", "small", "ties_unsorted", "ragged_L", "no_admissible"])
def test_numpy_oracle_matches_reference(name):
    g = load_search_golden(name)
    idx = np.linspace(0, len(g["periods"]) - 1, 12).astype(int)
    out = [oracle.search_period_numpy(g["periods"][k], g["t"], g["y"], g["dy"], g["templates"], g["params"]) for k in idx]
    got = (np.array([o[0] for o in out]), np.array([o[1] for o in out]), np.array([o[2] for o in out]))
    sub = dict(g, chi2=g["chi2"][idx], row=g["row"][idx], depth=g["depth"][idx])
    assert_search_parity(got, sub, rtol=1e-9, label=name)


def test_single_thread_equals_all_threads():
    g = load_search_golden("small")
    a = oracle.search_periods_c(g["t"], g["y"], g["dy"], g["periods"], g["templates"], g["params"], threads=1)
    b = oracle.search_periods_c(g["t"], g["y"], g["dy"], g["periods"], g["templates"], g["params"], threads=0)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize("name", [n for n in t0fit_goldens() if n != "cfg2_margin1"])
def test_numpy_t0_fit_matches_reference(name):
    """oracle.final_T0_fit_numpy against T0 returned by the reference's stats.final_T0_fit."""
    g = load_t0fit_golden(name)
    dy = np.full(len(g["y"]), np.std(g["y"]))
    T0, resid, trials = oracle.final_T0_fit_numpy(g["signal"], g["depth"], g["t"], g["y"], dy, g["period"], g["margin"])
    assert T0 == g["T0"]
    np.testing.assert_array_equal(trials, g["trials"])
    np.testing.assert_allclose(resid, g["residuals"], rtol=1e-12)


@pytest.mark.parametrize("name", ["cfg1_50ppm", "small_hetero", "k2_epic201367065", "k2_epic206154641_box"])
def test_numpy_spectra_matches_reference(name):
    """oracle.spectra_numpy against the reference's own SR / power_raw / power / SDE (power_*.npz)."""
    import os
    from conftest import GOLDEN

    z = np.load(os.path.join(GOLDEN, "power_%s.npz" % name))
    kw = eval(str(z["kwargs"]), {"__builtins__": {}})
    SR, pr, pw, sde_raw, sde = oracle.spectra_numpy(z["a_chi2"], kw.get("oversampling_factor", 3))
    np.testing.assert_allclose(SR, z["a_SR"], rtol=1e-13)
    np.testing.assert_allclose(pr, z["a_power_raw"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(pw, z["a_power"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose([sde_raw, sde], [float(z["s_SDE_raw"]), float(z["s_SDE"])], rtol=1e-12)
