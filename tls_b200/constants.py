"""Numeric constants of the TLS search, by value.

Every number here is an input to the arithmetic that the parity tests pin, so
each carries the reference location it mirrors
(``/root/reference/transitleastsquares/tls_constants.py``).
"""
from os import path

VERSION = "tls-b200 0.1 (hot path of TLS 1.0.31 on sm_100a)"

# physical constants (SI)                      tls_constants.py:20-25
G = 6.673e-11
R_sun = 695508000
R_earth = 6371000
R_jup = 69911000
M_sun = 1.989 * 10 ** 30
SECONDS_PER_DAY = 86400

# search defaults                              tls_constants.py:28-44
TRANSIT_DEPTH_MIN = 10 * 10 ** -6
NUMERICAL_STABILITY_CUTOFF = 0.01 * 10 ** -6
R_STAR = 1.0
M_STAR = 1.0
OVERSAMPLING_FACTOR = 3
N_TRANSITS_MIN = 2
M_STAR_MIN = 0.1
M_STAR_MAX = 1.0
R_STAR_MIN = 0.13
R_STAR_MAX = 3.5
DURATION_GRID_STEP = 1.1

# template presets                             tls_constants.py:47-66
TEMPLATES = {
    "default": dict(per=12.9, rp=0.03, a=23.1, inc=89.21),
    "grazing": dict(b=0.99),
    "box": dict(per=29, rp=0.1, a=26.9, b=0, inc=90, u=[0], limb_dark="linear"),
}
DEFAULT_U = [0.4804, 0.1867]
DEFAULT_LIMB_DARK = "quadratic"
DEFAULT_ECC = 0
DEFAULT_W = 90

# template construction                        tls_constants.py:71,78,89-90
SIGNAL_DEPTH = 0.5
FRACTIONAL_TRANSIT_DURATION_MAX = 0.12
SUPERSAMPLE_SIZE = 10000
OVERSAMPLE_MODEL_LIGHT_CURVE = 5

# post-processing                              tls_constants.py:100,109,114,118
SDE_MEDIAN_KERNEL_SIZE = 30
T0_FIT_MARGIN = 0.01
PROGRESSBAR_THRESHOLD = 5000
MINIMUM_PERIOD_GRID_SIZE = 100

# accepted power() keywords                    tls_constants.py:121-148
VALID_PARAMETERS = (
    "R_star R_star_min R_star_max M_star M_star_min M_star_max period_min period_max "
    "n_transits_min per rp a inc b ecc w u limb_dark duration_grid_step transit_depth_min "
    "oversampling_factor T0_fit_margin use_threads show_progress_bar transit_template verbose"
).split()

# keywords that only exist in this implementation (device selection)
EXTRA_PARAMETERS = ("device", "devices", "dist")

resources_dir = path.join(path.dirname(__file__), "data")
