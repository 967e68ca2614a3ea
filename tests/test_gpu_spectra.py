"""stats.spectra on the GPU (``tlsb_spectra`` through the C ABI) against the reference's own
outputs (tests/golden/power_*.npz) and the oracle's restatement: SR, power_raw, power within
1e-9 (bar: 1e-5), SDE within 1e-9, arg-max index exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["cfg1_50ppm", "small_hetero", "k2_epic201367065", "k2_epic206154641_box"])
def test_spectra_matches_reference(name):
    from tls_b200 import native, stats

    z = np.load(os.path.join(GOLDEN, "power_%s.npz" % name))
    kw = eval(str(z["kwargs"]), {"__builtins__": {}})
    osf = kw.get("oversampling_factor", 3)
    SR, pr, pw, sde_raw, sde, amax = native.spectra(z["a_chi2"], stats.median_window(osf))
    np.testing.assert_allclose(SR, z["a_SR"], rtol=1e-12)
    np.testing.assert_allclose(pr, z["a_power_raw"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(pw, z["a_power"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose([sde_raw, sde], [float(z["s_SDE_raw"]), float(z["s_SDE"])], rtol=1e-10)
    assert amax == int(np.argmax(z["a_power"]))
    got = stats.spectra(z["a_chi2"], osf)  # the drop-in function
    np.testing.assert_array_equal(got[2], pw)


@pytest.mark.parametrize("P,window", [(50, 91), (183, 91), (184, 91), (700, 31), (5000, 301), (3001, 82), (20011, 91)])
def test_spectra_matches_oracle_on_random_rows(P, window):
    """Short rows (no detrending, stats.py:130-131), even windows (numpy.median averages the middle
    two), +inf entries (periods without an admissible duration), several curves per call."""
    from tls_b200 import native

    rng = np.random.RandomState(P + window)
    rows = 4000.0 + rng.normal(0, 3, (3, P)) - 40 * np.exp(-0.5 * ((np.arange(P) - P // 3) / 2.0) ** 2)
    rows[1, :: max(1, P // 7)] = np.inf
    rows[2] = np.round(rows[2])  # many exact ties inside the median windows
    SR, pr, pw, sde_raw, sde, amax = native.spectra(rows, window)
    for c in range(3):
        w = oracle.spectra_numpy(rows[c], None, kernel=window)
        np.testing.assert_allclose(SR[c], w[0], rtol=1e-12)
        np.testing.assert_allclose(pr[c], w[1], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(pw[c], w[2], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose([sde_raw[c], sde[c]], [w[3], w[4]], rtol=1e-10)
        assert amax[c] == int(np.argmax(w[2]))


def test_spectra_argument_errors():
    from tls_b200 import native

    with pytest.raises(RuntimeError, match="median window"):
        native.spectra(np.ones(10) + np.arange(10), 0)
