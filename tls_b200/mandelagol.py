"""Closed-form transit light curve for the quadratic (and linear, uniform) limb-darkening law.

The reference builds its template bank with ``batman`` (``/root/reference/transitleastsquares/transit.py:8-42``;
un-vendored, unpinned ``setup.py:41``, absent from this image).  For its default law batman evaluates the analytic
model of Mandel & Agol (2002, ApJ 580, L171): eq. (1) for a uniform source and eq. (7) with Table 1 for the
quadratic law, in terms of the complete elliptic integrals K, E and Pi.  This module restates those published
formulas (Kreidberg 2015 describes batman's use of them) with numpy/scipy:

    F(p, z) = 1 - [ (1 - c2) lambda_e + c2 (lambda_d + 2/3 Theta(p - z)) - c4 eta_d ] / (1 - c2/3 - c4/2),
    c2 = u1 + 2 u2,  c4 = -u2,

p = planet radius, z = projected separation (stellar radii).  K and E come from ``scipy.special.ellipk/ellipe``
(double precision, not the polynomial approximations of the classic occultquad code) and Pi from Carlson's
symmetric forms, Pi(n, k) = R_F(0, 1-k^2, 1) + (n/3) R_J(0, 1-k^2, 1, 1-n) for the convention
Pi(n, k) = int_0^{pi/2} dtheta / ((1 - n sin^2 theta) sqrt(1 - k^2 sin^2 theta)) that Table 1 uses.

:mod:`tls_b200.limbdark` (a quadrature over stellar radius, any law) is the independent cross-check:
``tests/test_host.py::test_closed_form_mandel_agol_equals_the_quadrature``.
"""
from __future__ import annotations

import numpy as np
from scipy import special

__all__ = ["quadratic_flux", "uniform_lambda"]


def _ellip_pi(n, k2):
    """Complete elliptic integral of the third kind, Pi(n, k) with k^2 = k2 < 1 and n < 1."""
    y = 1.0 - k2
    return special.elliprf(0.0, y, 1.0) + n / 3.0 * special.elliprj(0.0, y, 1.0, 1.0 - n)


def uniform_lambda(p, z):
    """lambda_e(p, z): fraction of a UNIFORM stellar disc the planet covers (Mandel & Agol 2002, eq. 1)."""
    z = np.asarray(z, dtype=float)
    lam = np.zeros_like(z)
    if p <= 0:
        return lam
    full = z <= p - 1.0                      # planet covers the whole star
    inside = (z <= 1.0 - p) & ~full          # planet entirely on the disc
    part = (z > abs(1.0 - p)) & (z < 1.0 + p)
    lam[full] = 1.0
    lam[inside] = p * p
    if np.any(part):
        zz = z[part]
        k1 = np.arccos(np.clip((1.0 - p * p + zz * zz) / (2.0 * zz), -1.0, 1.0))
        k0 = np.arccos(np.clip((p * p + zz * zz - 1.0) / (2.0 * p * zz), -1.0, 1.0))
        root = np.sqrt(np.clip((4.0 * zz * zz - (1.0 + zz * zz - p * p) ** 2) / 4.0, 0.0, None))
        lam[part] = (p * p * k0 + k1 - root) / np.pi
    return lam


def _lambda1_eta1(p, z):
    """Table 1, cases II and VIII: the planet's disc crosses the stellar limb."""
    a, b = (z - p) ** 2, (z + p) ** 2
    q = p * p - z * z
    k2 = (1.0 - a) / (4.0 * z * p)
    K, E = special.ellipk(k2), special.ellipe(k2)
    Pi = _ellip_pi((a - 1.0) / a, k2)
    lam = (((1.0 - b) * (2.0 * b + a - 3.0) - 3.0 * q * (b - 2.0)) * K
           + 4.0 * p * z * (z * z + 7.0 * p * p - 4.0) * E
           - 3.0 * (q / a) * Pi) / (9.0 * np.pi * np.sqrt(p * z))
    k1 = np.arccos(np.clip((1.0 - p * p + z * z) / (2.0 * z), -1.0, 1.0))
    k0 = np.arccos(np.clip((p * p + z * z - 1.0) / (2.0 * p * z), -1.0, 1.0))
    eta2 = 0.5 * p * p * (p * p + 2.0 * z * z)
    eta = (k1 + 2.0 * eta2 * k0 - 0.25 * (1.0 + 5.0 * p * p + z * z) * np.sqrt(np.clip((1.0 - a) * (b - 1.0), 0.0, None))) / (2.0 * np.pi)
    return lam, eta


def _lambda2(p, z):
    """Table 1, cases III and IX: the planet's disc lies inside the stellar disc and does not cover its centre
    (III) or covers it (IX)."""
    a, b = (z - p) ** 2, (z + p) ** 2
    q = p * p - z * z
    inv_k2 = (4.0 * z * p) / (1.0 - a)  # (1/k)^2 < 1 here
    K, E = special.ellipk(inv_k2), special.ellipe(inv_k2)
    Pi = _ellip_pi((a - b) / a, inv_k2)
    return 2.0 / (9.0 * np.pi * np.sqrt(1.0 - a)) * (
        (1.0 - 5.0 * z * z + p * p + q * q) * K + (1.0 - a) * (z * z + 7.0 * p * p - 4.0) * E - 3.0 * (q / a) * Pi)


def _lambda_eta_at_z_equal_p(p):
    """Table 1, cases V (p < 1/2), VI (p = 1/2) and VII (p > 1/2): the planet's limb passes through the stellar
    centre, z = p.  (The moduli follow the published erratum / the classic occultquad code: 2p below 1/2, 1/(2p) above.)"""
    if abs(p - 0.5) < 1e-14:
        return 1.0 / 3.0 - 4.0 / (9.0 * np.pi), 3.0 / 32.0
    if p < 0.5:
        m = 4.0 * p * p
        lam = 1.0 / 3.0 + 2.0 / (9.0 * np.pi) * (4.0 * (2.0 * p * p - 1.0) * special.ellipe(m) + (1.0 - 4.0 * p * p) * special.ellipk(m))
        return float(lam), 1.5 * p ** 4
    m = 1.0 / (4.0 * p * p)
    lam = (1.0 / 3.0 + 16.0 * p / (9.0 * np.pi) * (2.0 * p * p - 1.0) * special.ellipe(m)
           - (32.0 * p ** 4 - 20.0 * p * p + 3.0) / (9.0 * np.pi * p) * special.ellipk(m))
    z = np.array([p])
    k1 = np.arccos(np.clip((1.0 - p * p + z * z) / (2.0 * z), -1.0, 1.0))
    k0 = np.arccos(np.clip((p * p + z * z - 1.0) / (2.0 * p * z), -1.0, 1.0))
    a, b = 0.0, 4.0 * p * p
    eta = (k1 + 2.0 * (1.5 * p ** 4) * k0 - 0.25 * (1.0 + 6.0 * p * p) * np.sqrt(max((1.0 - a) * (b - 1.0), 0.0))) / (2.0 * np.pi)
    return float(lam), float(eta[0])


def quadratic_flux(z, p, u1, u2=0.0):
    """Relative flux (1 = unocculted) of a star with I(mu) = 1 - u1 (1 - mu) - u2 (1 - mu)^2 behind a dark disc of
    radius ``p`` at separations ``z`` (array).  u2 = 0 is the linear law, u1 = u2 = 0 the uniform disc."""
    z = np.abs(np.asarray(z, dtype=float))
    p = float(abs(p))
    flux = np.ones_like(z)
    if p == 0.0:
        return flux
    # Around z = p the general expressions of Table 1 lose digits (q/a ~ 1/(z - p) against a diverging Pi): inside
    # |z - p| < delta the curve is the parabola through the closed form at p - delta, p (cases V-VII) and p + delta
    delta = 3e-5
    near = np.abs(z - p) < delta
    if np.any(near) and p > delta and abs(1.0 - 2.0 * p) > 2.0 * delta:
        f_lo, f_hi = _quadratic_flux_general(np.array([p - delta, p + delta]), p, u1, u2)
        lam_p, eta_p = _lambda_eta_at_z_equal_p(p)
        c2, c4 = u1 + 2.0 * u2, -u2
        lam_e_p = float(uniform_lambda(p, np.array([p]))[0])
        f_mid = 1.0 - ((1.0 - c2) * lam_e_p + c2 * (lam_p + 0.0) - c4 * eta_p) / (1.0 - c2 / 3.0 - c4 / 2.0)
        # Theta(p - z) jumps at z = p while lambda_d jumps by -2/3: their sum is continuous; f_mid takes the z -> p+ side
        x = (z[near] - p) / delta
        flux = _quadratic_flux_general(np.where(near, p + 2.0 * delta, z), p, u1, u2)
        flux[near] = f_mid + 0.5 * (f_hi - f_lo) * x + (0.5 * (f_hi + f_lo) - f_mid) * x * x
        return flux
    return _quadratic_flux_general(z, p, u1, u2)


def _quadratic_flux_general(z, p, u1, u2):
    flux = np.ones_like(z)
    c2, c4 = u1 + 2.0 * u2, -u2
    omega4 = 1.0 - c2 / 3.0 - c4 / 2.0
    lam_e = uniform_lambda(p, z)
    lam_d = np.zeros_like(z)
    eta_d = np.zeros_like(z)
    tiny = 1e-12  # the measure-zero special cases of Table 1 (z = p, z = 1 - p, z = 0) are taken as limits

    # XI: the star is entirely behind the planet
    total = z <= p - 1.0
    eta_d[total] = 0.5
    # X: concentric (also the limit of IX)
    centre = (z < tiny) & ~total
    lam_d[centre] = -2.0 / 3.0 * (1.0 - min(p, 1.0) ** 2) ** 1.5
    eta_d[centre] = 0.5 * p ** 4
    # II / VIII: the limb is crossed
    limb = (z > abs(1.0 - p) + tiny) & (z < 1.0 + p) & ~centre & ~total
    # z = p exactly (cases V, VII) is the continuous limit of its neighbours: step off it
    zz = np.where(np.abs(z - p) < tiny, p + 2.0 * tiny, z)
    if np.any(limb):
        lam_d[limb], eta_d[limb] = _lambda1_eta1(p, zz[limb])
    # III / IX: entirely on the disc (p < 1)
    inner = (z <= abs(1.0 - p) + tiny) & (z <= 1.0 - p + tiny) & ~centre & ~total & ~limb
    if np.any(inner):
        zi = np.minimum(zz[inner], 1.0 - p - tiny) if p < 1.0 else zz[inner]
        lam_d[inner] = _lambda2(p, zi)
        eta_d[inner] = 0.5 * p * p * (p * p + 2.0 * zi * zi)
    theta = (p > z).astype(float)
    blocked = ((1.0 - c2) * lam_e + c2 * (lam_d + 2.0 / 3.0 * theta) - c4 * eta_d) / omega4
    touching = z < 1.0 + p
    flux[touching] = 1.0 - blocked[touching]
    return flux
