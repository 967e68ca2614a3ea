#!/usr/bin/env python
"""A few searches of a workload with per-point uncertainties (for ncu).  usage: python scripts/prof_hetero.py [workload]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tls_b200 import native, transitleastsquares, workloads
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
t, y, dy, kw = workloads.lightcurve(wl, hetero=True)
inp = transitleastsquares(t, y, dy, verbose=False).prepare(verbose=False, **kw)
s = native.Searcher()
s.set_inputs(inp.t, inp.y, inp.dy, inp.templates, inp.params)
s.set_periods(inp.periods)
for _ in range(4):
    s.search_async(); s.results()
print(wl, "hetero kernel %.3f ms" % s.kernel_ms, s.layout)
s.close()
