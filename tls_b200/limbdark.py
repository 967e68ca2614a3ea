"""Limb-darkened transit light-curve model (CPU, numpy), written from scratch.

The reference builds its transit template by calling the third-party package
``batman`` (``/root/reference/transitleastsquares/transit.py:14-25``); batman is
an un-vendored, unpinned dependency (``setup.py:41``) that is absent from this
image, so the template builder here needs its own model.  This module restates
the *published* model that batman implements (Mandel & Agol 2002; Kreidberg
2015): a dark planet disc of radius ``rp`` (stellar radii) at projected
separation ``z`` blocks the part of a limb-darkened stellar disc it overlaps.

For the quadratic, linear and uniform laws ``TransitModel`` evaluates the closed-form
elliptic-integral expressions of that paper (:mod:`tls_b200.mandelagol`, what batman
evaluates for its default law).  For every other law, and as the independent cross-check
of the closed form (they agree to 1e-11, ``tests/test_host.py``), the blocked flux is
computed for *any* radial intensity profile I(r) by a 1-D quadrature over stellar radius r::

    blocked(z) = int_{0}^{1} I(r) * 2*kappa(r; z, rp) * r dr
    kappa      = pi                         if r <= rp - z      (ring fully covered)
               = 0                          if r <= z - rp or r >= z + rp
               = arccos((r^2+z^2-rp^2)/(2 r z))   otherwise

with a sin^2 substitution that removes the square-root end-point
singularities, and Gauss-Legendre nodes in the substituted variable.  With the
default 384 nodes the result is converged to ~1e-13 for the template shapes
TLS uses, i.e. far below the 1e-8 trimming threshold of ``transit.py:144-149``.

Host-side, once per search (~10 ms); not part of the GPU hot path.
"""
from __future__ import annotations

import functools
import os

import numpy as np

__all__ = ["TransitParams", "TransitModel", "separation", "occulted_flux", "LAWS"]

_TWO_PI = 2.0 * np.pi


def _profile(law: str, u):
    """Return I(mu) for a batman-style law name and coefficient list."""
    u = [float(v) for v in np.atleast_1d(u)] if u is not None else []
    if law == "uniform":
        return lambda mu: np.ones_like(mu)
    if law == "linear":
        (c1,) = u
        return lambda mu: 1.0 - c1 * (1.0 - mu)
    if law == "quadratic":
        c1, c2 = u
        return lambda mu: 1.0 - c1 * (1.0 - mu) - c2 * (1.0 - mu) ** 2
    if law == "squareroot":
        c1, c2 = u
        return lambda mu: 1.0 - c1 * (1.0 - mu) - c2 * (1.0 - np.sqrt(mu))
    if law == "logarithmic":
        c1, c2 = u

        def f(mu):
            safe = np.where(mu > 0, mu, 1.0)
            return 1.0 - c1 * (1.0 - mu) - c2 * mu * np.log(safe)

        return f
    if law == "exponential":
        c1, c2 = u

        def f(mu):
            safe = np.where(mu > 0, mu, 1.0)
            term = np.where(mu > 0, c2 / (1.0 - np.exp(safe)), 0.0)
            return 1.0 - c1 * (1.0 - mu) - term

        return f
    if law == "power2":
        c1, c2 = u
        return lambda mu: 1.0 - c1 * (1.0 - mu ** c2)
    if law == "nonlinear":
        c1, c2, c3, c4 = u
        return lambda mu: (
            1.0
            - c1 * (1.0 - np.sqrt(mu))
            - c2 * (1.0 - mu)
            - c3 * (1.0 - mu ** 1.5)
            - c4 * (1.0 - mu ** 2)
        )
    raise ValueError("unsupported limb darkening law: " + str(law))


LAWS = (
    "uniform",
    "linear",
    "quadratic",
    "squareroot",
    "logarithmic",
    "exponential",
    "power2",
    "nonlinear",
)


@functools.lru_cache(maxsize=16)
def _gl(n):
    """Gauss-Legendre nodes and weights.  numpy solves an n x n eigenvalue problem for them
    (0.5 s per order, most of a first ``.power()`` call), so the two orders this module uses
    ship as a table (``data/gauss_legendre.npz`` = ``leggauss(256)`` and ``leggauss(384)``,
    bit for bit); any other order is computed."""
    table = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "gauss_legendre.npz")
    x = w = None
    if os.path.exists(table):
        with np.load(table) as z:
            if "x%d" % n in z.files:
                x, w = np.array(z["x%d" % n]), np.array(z["w%d" % n])
    if x is None:
        x, w = np.polynomial.legendre.leggauss(n)
    x.setflags(write=False)
    w.setflags(write=False)
    return x, w


def _mu_of_r(r):
    return np.sqrt(np.clip(1.0 - r * r, 0.0, None))


def _segment_integral(func, a, b, nodes):
    """int_a^b func(r) dr for arrays a, b (same shape) with r = a + (b-a) sin^2(theta/2)."""
    x, w = nodes
    theta = 0.5 * np.pi * (x + 1.0)  # [0, pi]
    wt = 0.5 * np.pi * w
    s2 = np.sin(0.5 * theta) ** 2
    jac = 0.5 * np.sin(theta)
    span = (b - a)[..., None]
    r = a[..., None] + span * s2
    return np.sum(func(r) * (span * jac * wt), axis=-1)


def occulted_flux(z, rp, law="quadratic", u=(0.4804, 0.1867), n_nodes=384):
    """Relative flux (1 = unocculted) for separations ``z`` (array, stellar radii)."""
    z = np.abs(np.asarray(z, dtype=float))
    rp = float(abs(rp))
    inten = _profile(law, u)
    nodes = _gl(n_nodes)

    # total stellar flux: int_0^1 I 2 pi r dr with r = sin(phi)
    x, w = _gl(256)
    phi = 0.25 * np.pi * (x + 1.0)
    total = _TWO_PI * np.sum(inten(np.cos(phi)) * np.sin(phi) * np.cos(phi) * w) * 0.25 * np.pi

    flux = np.ones_like(z)
    touching = z < 1.0 + rp
    if not np.any(touching) or rp == 0.0:
        return flux
    zt = z[touching]
    blocked = np.zeros_like(zt)

    # rings completely covered by the planet (only when the planet overlaps the centre)
    full_hi = np.clip(rp - zt, 0.0, 1.0)
    has_full = full_hi > 0
    if np.any(has_full):
        a = np.zeros(np.count_nonzero(has_full))
        b = full_hi[has_full]
        blocked[has_full] += _segment_integral(
            lambda r: inten(_mu_of_r(r)) * _TWO_PI * r, a, b, nodes
        )

    # partially covered rings
    lo = np.clip(np.abs(zt - rp), 0.0, 1.0)
    hi = np.clip(zt + rp, 0.0, 1.0)
    part = (hi > lo) & (zt > 0)
    if np.any(part):
        zz = zt[part][..., None]

        def integrand(r):
            rs = np.where(r > 0, r, 1.0)
            c = (r * r + zz * zz - rp * rp) / (2.0 * rs * zz)
            kappa = np.arccos(np.clip(c, -1.0, 1.0))
            return inten(_mu_of_r(r)) * 2.0 * kappa * r

        blocked[part] += _segment_integral(integrand, lo[part], hi[part], nodes)

    flux[touching] = 1.0 - blocked / total
    return flux


def _kepler_E(M, ecc, iters=60):
    """Solve Kepler's equation E - e sin E = M (Newton, vectorised)."""
    E = np.where(ecc < 0.8, M, np.pi * np.ones_like(M))
    for _ in range(iters):
        dE = (E - ecc * np.sin(E) - M) / (1.0 - ecc * np.cos(E))
        E = E - dE
        if np.max(np.abs(dE)) < 1e-15:
            break
    return E


def separation(t, t0, per, a, inc, ecc=0.0, w=90.0):
    """Sky-projected star-planet separation (stellar radii); +inf when the planet
    is behind the star (only primary transits are modelled, as in the reference's
    use of batman, ``transit.py:14-25``)."""
    t = np.asarray(t, dtype=float)
    inc_r = np.deg2rad(inc)
    w_r = np.deg2rad(w)
    # time of periastron from the time of inferior conjunction
    f_conj = 0.5 * np.pi - w_r
    E_conj = 2.0 * np.arctan(np.sqrt((1.0 - ecc) / (1.0 + ecc)) * np.tan(0.5 * f_conj))
    M_conj = E_conj - ecc * np.sin(E_conj)
    tp = t0 - per * M_conj / _TWO_PI
    M = _TWO_PI * ((t - tp) / per - np.floor((t - tp) / per))
    if ecc < 1e-5:
        f = M
    else:
        E = _kepler_E(M, ecc)
        f = 2.0 * np.arctan2(
            np.sqrt(1.0 + ecc) * np.sin(0.5 * E), np.sqrt(1.0 - ecc) * np.cos(0.5 * E)
        )
    r_orb = a * (1.0 - ecc * ecc) / (1.0 + ecc * np.cos(f))
    s = np.sin(w_r + f)
    d = r_orb * np.sqrt(np.clip(1.0 - (s * np.sin(inc_r)) ** 2, 0.0, None))
    in_front = s * np.sin(inc_r) > 0.0
    return np.where(in_front, d, np.inf)


class TransitParams(object):
    """Plain attribute bag with the fields the reference fills (``transit.py:14-23``)."""

    def __init__(self):
        self.t0 = 0.0
        self.per = 1.0
        self.rp = 0.1
        self.a = 10.0
        self.inc = 90.0
        self.ecc = 0.0
        self.w = 90.0
        self.u = [0.4804, 0.1867]
        self.limb_dark = "quadratic"


class TransitModel(object):
    """``TransitModel(params, t).light_curve(params)`` — the two calls the reference
    makes (``transit.py:24-25``)."""

    def __init__(self, params, t, n_nodes=384, closed_form=True):
        self.t = np.asarray(t, dtype=float)
        self.n_nodes = n_nodes
        self.closed_form = closed_form  # quadratic / linear / uniform: Mandel & Agol's analytic model (what batman evaluates)

    def light_curve(self, params):
        z = separation(
            self.t, params.t0, params.per, params.a, params.inc, params.ecc, params.w
        )
        finite = np.isfinite(z)
        flux = np.ones_like(self.t)
        if np.any(finite):
            law = params.limb_dark
            if self.closed_form and law in ("quadratic", "linear", "uniform"):
                from . import mandelagol

                u = [float(v) for v in np.atleast_1d(params.u)] if law != "uniform" else []
                u1 = u[0] if len(u) > 0 else 0.0
                u2 = u[1] if len(u) > 1 else 0.0
                flux[finite] = mandelagol.quadratic_flux(z[finite], params.rp, u1, u2)
            else:
                flux[finite] = occulted_flux(z[finite], params.rp, law, params.u, self.n_nodes)
        return flux
