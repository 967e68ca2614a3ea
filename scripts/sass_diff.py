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
        m = re.search(r"\d+(tlsb_\w+?_kernel)(I.*?EEE)?", name.strip())  # kernel + template arguments, without the
        key = m.group(1) + (m.group(2) or "") if m else name.strip()       # per-translation-unit namespace hash
        out[key] = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(.*?;)", body)        # the instruction stream only
    return out


a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
short = lambda k: k[:90]
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
