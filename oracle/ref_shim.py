"""TEST INFRASTRUCTURE — not part of the product.

Imports the *unmodified* reference package from ``/root/reference`` so that its
own numba-compiled hot path (``transitleastsquares/core.py``) can be run in the
build container to pin the oracle and to generate the golden vectors under
``tests/golden/``.  The reference imports the third-party ``batman`` package at
module top (``transit.py:2``); batman is not installed in this image, so a
stand-in module backed by :mod:`tls_b200.limbdark` (same published model,
independent implementation) is placed in ``sys.modules`` first.  Only the
template *values* depend on it; ``core.search_period`` itself is untouched.

``/root/reference`` does not exist on the GPU box.  There the unmodified copy that
``oracle/vendor_ref.py`` put under ``oracle/_ref/`` (git-ignored, shipped with the gpurun
snapshot) is imported instead, for one purpose only: timing the reference's own numba
hot path as the CPU baseline of ``bench.py`` (``oracle/time_reference.py``).
``available()`` tells whether either copy is there.
"""
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)


VENDORED_ROOT = os.path.join(_HERE, "_ref")


def reference_root():
    """Where the reference package can be imported from: the read-only tree, else the vendored copy, else None."""
    for root in (REFERENCE_ROOT, VENDORED_ROOT):
        if os.path.isfile(os.path.join(root, "transitleastsquares", "core.py")):
            return root
    return None


def available():
    return reference_root() is not None


def source_tree_available():
    """The read-only reference tree itself (build container only): what golden vectors are generated from."""
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
    root = reference_root()
    if root is None:
        raise RuntimeError("reference package found neither at %s nor vendored under %s" % (REFERENCE_ROOT, VENDORED_ROOT))
    _install_batman_standin()
    if root not in sys.path:
        sys.path.insert(0, root)
    import transitleastsquares  # noqa: E402

    return transitleastsquares
