#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE — not part of the product.

Puts an UNMODIFIED copy of the reference package (``/root/reference/transitleastsquares``, the
python modules and their data tables, not its tests) under ``oracle/_ref/`` so that the reference's own
numba hot path can be timed on the GPU box, where ``/root/reference`` does not exist.  ``oracle/_ref/``
is an output directory: git-ignored (the reference's sources never enter this repository's history)
but shipped with the gpurun snapshot, like the built ``.so`` files.  ``__graft_entry__.build()`` runs
this whenever the reference tree is present.

usage: python oracle/vendor_ref.py           -> oracle/_ref/transitleastsquares/, oracle/_ref/VENDORED.json
"""
import hashlib
import json
import os
import shutil
import sys

SRC = "/root/reference/transitleastsquares"
HERE = os.path.dirname(os.path.abspath(__file__))
DST_ROOT = os.path.join(HERE, "_ref")
DST = os.path.join(DST_ROOT, "transitleastsquares")


def available():
    return os.path.isdir(SRC)


def vendored():
    return os.path.isfile(os.path.join(DST, "core.py"))


def vendor(force=False):
    """Copy the package (top-level files only: modules + csv tables).  Returns the destination or None."""
    if not available():
        return DST if vendored() else None
    names = sorted(n for n in os.listdir(SRC) if os.path.isfile(os.path.join(SRC, n)))
    if vendored() and not force:
        same = all(os.path.exists(os.path.join(DST, n)) and
                   os.path.getsize(os.path.join(DST, n)) == os.path.getsize(os.path.join(SRC, n)) for n in names)
        if same:
            return DST
    os.makedirs(DST, exist_ok=True)
    digest = {}
    for n in names:
        shutil.copyfile(os.path.join(SRC, n), os.path.join(DST, n))
        with open(os.path.join(DST, n), "rb") as f:
            digest[n] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST_ROOT, "VENDORED.json"), "w") as f:
        json.dump({"source": SRC, "files": digest,
                   "note": "unmodified copy made by oracle/vendor_ref.py; batman is supplied by oracle/ref_shim.py"}, f, indent=1)
    return DST


if __name__ == "__main__":
    out = vendor(force="--force" in sys.argv)
    print(out or "reference tree not present and nothing vendored")
