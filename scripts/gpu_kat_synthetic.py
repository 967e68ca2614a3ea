#!/usr/bin/env python
"""The reference's known answer that depends on the transit MODEL (tests/test_synthetic.py:50:
chi2_min = 8831.654060613922 to 5 decimals, made with genuine batman): the test's data and search reproduced with
this repository's two independent models — the closed-form Mandel & Agol expressions and the radial quadrature.
usage: python scripts/gpu_kat_synthetic.py"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
from tls_b200 import limbdark, transit, transitleastsquares

WANT = dict(chi2_min=8831.654060613922, chi2red_min=0.6719152511118321, period=365.2582192473641)
keep = limbdark.TransitModel.__init__.__defaults__
for name, closed in (("closed form (tls_b200.mandelagol)", True), ("radial quadrature, 384 nodes (tls_b200.limbdark)", False)):
    limbdark.TransitModel.__init__.__defaults__ = (384, closed)
    transit.clear_caches()
    np.random.seed(seed=0)
    start, days, spd = 48, 365.25 * 3, 12
    samples = int(days * spd)
    t = np.linspace(start, start + days, samples)
    ma = limbdark.TransitParams()
    ma.t0, ma.per, ma.rp, ma.a, ma.inc, ma.ecc, ma.w, ma.u, ma.limb_dark = start + 20, 365.25, 6371 / 696342, 217, 90, 0, 90, [0.5], "linear"
    y = limbdark.TransitModel(ma, t).light_curve(ma) + np.random.normal(0, 5e-6, samples)
    y[1] = np.nan
    res = transitleastsquares(t, y, verbose=False).power(period_min=360, period_max=370, transit_depth_min=10e-6, oversampling_factor=5,
                                                         duration_grid_step=1.02, verbose=False, use_threads=1, show_progress_bar=False)
    print("%-50s chi2_min %.9f (want %.9f, diff %+.3e, rel %.2e)  chi2red_min %.12f  period %.10f" % (
        name, res.chi2_min, WANT["chi2_min"], res.chi2_min - WANT["chi2_min"], abs(res.chi2_min - WANT["chi2_min"]) / WANT["chi2_min"],
        res.chi2red_min, res.period))
limbdark.TransitModel.__init__.__defaults__ = keep
