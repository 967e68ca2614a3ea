#!/usr/bin/env python
"""Which kernels of two builds of libtlsb200.so differ?  Compares the SASS of every kernel (cuobjdump -sass),
function by function.  usage: scripts/sass_diff.py old.so new.so
Used to show that a change (a new instantiation, host-side code, an experiment behind an environment variable)
leaves the machine code of the measured kernels untouched when no GPU is at hand to re-run them."""
import re, subprocess, sys


def kernels(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out = {}
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        name, _, body = part.partition("\n")
        out[name.strip()] = body
    return out


a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
short = lambda k: re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+\d\d", "", k)[:90]
changed = [k for k in a if k in b and a[k] != b[k]]
print("%d kernels in %s, %d in %s" % (len(a), sys.argv[1], len(b), sys.argv[2]))
print("identical: %d" % sum(1 for k in a if k in b and a[k] == b[k]))
for k in changed:
    print("  CHANGED ", short(k))
for k in b:
    if k not in a:
        print("  NEW     ", short(k))
for k in a:
    if k not in b:
        print("  REMOVED ", short(k))
sys.exit(1 if changed else 0)
