"""TEST INFRASTRUCTURE — not part of the product.

Imports the *unmodified* reference package from ``/root/reference`` so that its
own numba-compiled hot path (``transitleastsquares/core.py``) can be run in the
build container to pin the oracle and to generate the golden vectors under
``tests/golden/``.  The reference imports the third-party ``batman`` package at
module top (``transit.py:2``); batman is not installed in this image, so a
stand-in module backed by :mod:`tls_b200.limbdark` (same published model,
independent implementation) is placed in ``sys.modules`` first.  Only the
template *values* depend on it; ``core.search_period`` itself is untouched.

``/root/reference`` does not exist on the GPU box: nothing that runs there may
import this file (``available()`` tells).
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "transitleastsquares"))


def _install_batman_standin():
    if "batman" in sys.modules:
        return
    if _REPO not in sys.path:
        sys.path.insert(0, _REPO)
    from tls_b200 import limbdark

    mod = types.ModuleType("batman")
    mod.TransitParams = limbdark.TransitParams
    mod.TransitModel = limbdark.TransitModel
    mod.__doc__ = "stand-in for batman-package backed by tls_b200.limbdark"
    sys.modules["batman"] = mod


def load():
    """Return the reference package (``import transitleastsquares``)."""
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    _install_batman_standin()
    if "astroquery" not in sys.modules:
        # catalog.py imports astroquery lazily inside functions; nothing to stub.
        pass
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import transitleastsquares  # noqa: E402

    return transitleastsquares
