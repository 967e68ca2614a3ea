"""Input and keyword validation of the drop-in API (host side).

Mirrors the checks, defaults and error texts of
``/root/reference/transitleastsquares/validate.py`` (``validate_inputs`` :9-46,
``validate_args`` :49-181) so user code sees the same ``ValueError``s/warnings."""
from __future__ import annotations

import multiprocessing
import warnings

import numpy as np

from . import constants as C
from .helpers import cleaned_array, impact_to_inclination


def validate_inputs(t, y, dy):
    """Clean the arrays, normalise ``dy`` to mean 1 (or fill it with std(y)) and
    run the range checks of validate.py:9-46."""
    if dy is None:
        t, y = cleaned_array(t, y)
    else:
        t, y, dy = cleaned_array(t, y, dy)
        dy = dy / np.mean(dy)  # weights, not absolute errors (validate.py:18)

    if np.max(t) - np.min(t) <= 0:
        raise ValueError("Time duration must positive")
    if np.size(y) < 3 or np.size(t) < 3:
        raise ValueError("Too few values in data set")
    mean_flux = np.mean(y)
    if mean_flux > 1.01 or mean_flux < 0.99:
        warnings.warn(
            "Warning: The mean flux should be normalized to 1, but it was found to be "
            + str(mean_flux)
        )
    if np.min(y) < 0:
        raise ValueError("Flux values must be positive")
    if np.max(y) >= float("inf"):
        raise ValueError("Flux values must be finite")
    if dy is None:
        dy = np.full(len(y), np.std(y))  # validate.py:39-40
    if np.size(t) != np.size(y) or np.size(t) != np.size(dy):
        raise ValueError("Arrays (t, y, dy) must be of the same dimensions")
    if t.ndim != 1:
        raise ValueError("Inputs (t, y, dy) must be 1-dimensional")
    return t, y, dy


def _require_positive_finite(value, name):
    if value <= 0 or value >= float("inf"):
        raise ValueError(name + " must be positive")


def validate_args(self, kwargs):
    """kwargs -> attributes on ``self`` with the reference defaults (validate.py:49-181)."""
    get = kwargs.get
    self.verbose = get("verbose", True)
    for key in kwargs:
        if key not in C.VALID_PARAMETERS and key not in C.EXTRA_PARAMETERS:
            warnings.warn("Ignoring unknown parameter: " + str(key))

    self.show_progress_bar = get("show_progress_bar", True)
    self.transit_depth_min = get("transit_depth_min", C.TRANSIT_DEPTH_MIN)
    self.R_star = get("R_star", C.R_STAR)
    self.M_star = get("M_star", C.M_STAR)
    self.oversampling_factor = get("oversampling_factor", C.OVERSAMPLING_FACTOR)
    self.period_max = get("period_max", float("inf"))
    self.period_min = get("period_min", 0)
    self.n_transits_min = get("n_transits_min", C.N_TRANSITS_MIN)
    self.R_star_min = get("R_star_min", C.R_STAR_MIN)
    self.R_star_max = get("R_star_max", C.R_STAR_MAX)
    self.M_star_min = get("M_star_min", C.M_STAR_MIN)
    self.M_star_max = get("M_star_max", C.M_STAR_MAX)
    self.duration_grid_step = get("duration_grid_step", C.DURATION_GRID_STEP)
    self.use_threads = get("use_threads", multiprocessing.cpu_count())
    self.T0_fit_margin = get("T0_fit_margin", C.T0_FIT_MARGIN)

    default = C.TEMPLATES["default"]
    self.per = get("per", default["per"])
    self.rp = get("rp", default["rp"])
    self.a = get("a", default["a"])
    if "b" in kwargs:  # an impact parameter overrules the inclination (validate.py:86-90)
        self.b = get("b")
        self.inc = impact_to_inclination(b=self.b, semimajor_axis=self.a)
    else:
        self.inc = get("inc", default["inc"])
    self.ecc = get("ecc", C.DEFAULT_ECC)
    self.w = get("w", C.DEFAULT_W)
    self.u = get("u", C.DEFAULT_U)
    self.limb_dark = get("limb_dark", C.DEFAULT_LIMB_DARK)

    # validate.py:101-125: presets; "default" overrides user per/rp/a/inc
    self.transit_template = get("transit_template", "default")
    if self.transit_template == "default":
        self.per, self.rp, self.a, self.inc = (
            default["per"], default["rp"], default["a"], default["inc"],
        )
    elif self.transit_template == "grazing":
        self.b = C.TEMPLATES["grazing"]["b"]
        self.inc = impact_to_inclination(b=self.b, semimajor_axis=self.a)
    elif self.transit_template == "box":
        box = C.TEMPLATES["box"]
        self.per, self.rp, self.a, self.b = box["per"], box["rp"], box["a"], box["b"]
        self.inc, self.u, self.limb_dark = box["inc"], box["u"], box["limb_dark"]
    else:
        raise ValueError(
            'Unknown transit_template. Known values: "default", "grazing", "box"'
        )

    # validate.py:127-180
    _require_positive_finite(self.R_star, "R_star")
    if self.R_star_min > self.R_star:
        raise ValueError("R_star_min <= R_star is required")
    _require_positive_finite(self.R_star_min, "R_star_min")
    if self.R_star_max < self.R_star:
        raise ValueError("R_star_max >= R_star is required")
    _require_positive_finite(self.R_star_max, "R_star_max")
    _require_positive_finite(self.M_star, "M_star")
    if self.M_star_min > self.M_star:
        raise ValueError("M_star_min <= M_star is required")
    _require_positive_finite(self.M_star_min, "M_star_min")
    if self.M_star_max < self.M_star:
        raise ValueError("M_star_max >= M_star required")
    _require_positive_finite(self.M_star_max, "M_star_max")
    if self.period_min < 0:
        raise ValueError("period_min >= 0 required")
    if self.period_min >= self.period_max:
        raise ValueError("period_min < period_max required")
    if not isinstance(self.n_transits_min, int):
        raise ValueError("n_transits_min must be an integer value")
    if self.n_transits_min < 1:
        raise ValueError("n_transits_min must be an integer value >= 1")
    if not isinstance(self.use_threads, int) or self.use_threads < 1:
        raise ValueError("use_threads must be an integer value >= 1")
    if self.T0_fit_margin < 0:
        self.T0_fit_margin = 0
    elif self.T0_fit_margin > 0.1:
        self.T0_fit_margin = 0.1
    return self, kwargs
