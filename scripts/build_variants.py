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
    cmd = [build.nvcc_path()] + build.NVCC_FLAGS + ["-Xptxas", "-v"] + ["-D" + d for d in defines] + ["-o", out] + build.SRC
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    proc = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if proc.returncode:
        print(name, "FAILED\n", proc.stderr[-2000:]); continue
    lines = proc.stderr.splitlines()
    for i, l in enumerate(lines):
        if "Compiling entry function" in l and re.search(r"search_kernelILi256ELb1ELb1ELi[79]|search_tiled_kernelILi512ELb1ELi[579]|search_kernelILi256ELb1ELb0ELi5", l):
            kn = re.search(r"tlsb_\w+?kernelI\w+?EEE", l).group(0)
            print("%-10s %-46s %s | %s" % (name, kn, lines[i + 2].strip(), lines[i + 3].strip()[:60]))
