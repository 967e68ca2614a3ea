import glob
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")

PARAM_ORDER = ("transit_depth_min", "R_star_min", "R_star_max", "M_star_min", "M_star_max", "T0_fit_margin")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def search_goldens():
    return sorted(os.path.basename(p)[len("search_"):-4] for p in glob.glob(os.path.join(GOLDEN, "search_*.npz")))


def load_search_golden(name):
    """Inputs + reference outputs of core.search_period (made by oracle/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "search_%s.npz" % name))
    templates = dict(signal=z["tp_signal"], offset=z["tp_offset"], length=z["tp_length"],
                     width=z["tp_width"], overshoot=z["tp_overshoot"])
    params = dict(zip(PARAM_ORDER, [float(v) for v in z["params"]]))
    return dict(t=z["t"], y=z["y"], dy=z["dy"], periods=z["periods"], templates=templates,
                params=params, chi2=z["chi2"], row=z["row"], depth=z["depth"])


def assert_search_parity(got, want, rtol=1e-5, label=""):
    """The bar of BASELINE.json: argmin rows bit-exact, chi2/depth within 1e-5 relative.
    Sentinel (N) and inf values must be reproduced exactly."""
    chi2, row, depth = got
    special = ~np.isfinite(want["chi2"]) | (want["chi2"] == float(len(want["y"])))
    np.testing.assert_array_equal(np.asarray(row), want["row"], err_msg=label + " rows")
    np.testing.assert_array_equal(np.asarray(chi2)[special], want["chi2"][special], err_msg=label + " sentinel/inf chi2")
    np.testing.assert_array_equal(np.asarray(depth)[special], want["depth"][special], err_msg=label + " sentinel depth")
    np.testing.assert_allclose(np.asarray(chi2)[~special], want["chi2"][~special], rtol=rtol, atol=0, err_msg=label + " chi2")
    np.testing.assert_allclose(np.asarray(depth)[~special], want["depth"][~special], rtol=rtol, atol=0, err_msg=label + " depth")


@pytest.fixture(scope="session")
def has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def t0fit_goldens():
    return sorted(os.path.basename(p)[len("t0fit_"):-4] for p in glob.glob(os.path.join(GOLDEN, "t0fit_*.npz")))


def load_t0fit_golden(name):
    """Inputs of stats.final_T0_fit + the reference's T0 and the oracle's per-trial residuals."""
    z = np.load(os.path.join(GOLDEN, "t0fit_%s.npz" % name))
    return {k: (z[k] if z[k].ndim else z[k].item()) for k in z.files}
