#!/usr/bin/env python
"""Build experimental variants of libtlsb200.so (compile-time switches) for an A/B run on the GPU box.
usage: scripts/build_variants.py name:DEF=VAL,DEF=VAL [name:...]   ->  tls_b200/variants/lib_<name>.so
Prints registers / spills of the resident cfg-1 kernel for each."""
import os, re, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tls_b200 import build
os.makedirs(os.path.join(REPO, "tls_b200", "variants"), exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    defines = [d for d in defs.split(",") if d]
    out = os.path.join(REPO, "tls_b200", "variants", "lib_%s.so" % name)
    import contextlib, io
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            build.build(force=True, verbose=True, defines=defines, out=out)
    except RuntimeError as exc:
        print(name, "FAILED\n", str(exc)[-2000:]); continue
    class proc: stderr = buf.getvalue()
    lines = proc.stderr.splitlines()
    for i, l in enumerate(lines):
        if "Compiling entry function" in l and re.search(r"search_kernelILi256ELb1ELb1ELi[79]|search_tiled_kernelILi512ELb1ELi[579]|search_kernelILi256ELb1ELb0ELi5", l):
            kn = re.search(r"tlsb_\w+?kernelI\w+?EEE", l).group(0)
            print("%-10s %-46s %s | %s" % (name, kn, lines[i + 2].strip(), lines[i + 3].strip()[:60]))
