#!/usr/bin/env python
"""Kernel time of the search with per-point uncertainties (the general-weights kernel) next to the
equal-weights one, same light curves.  usage: python scripts/gpu_hetero_bench.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tls_b200 import native, transitleastsquares, workloads

for wl, stride in (("cfg1", 1), ("cfg1_500ppm", 1), ("cfg3", 1), ("cfg2", 24)):
    for hetero in (False, True):
        t, y, dy, kw = workloads.lightcurve(wl, hetero=hetero)
        inp = transitleastsquares(t, y, dy, verbose=False).prepare(verbose=False, **kw)
        per = inp.periods[::stride]
        s = native.Searcher()
        s.set_inputs(inp.t, inp.y, inp.dy, inp.templates, inp.params)
        s.set_periods(per)
        ms = []
        for _ in range(4):
            s.search_async(); s.results(); ms.append(s.kernel_ms)
        lay = s.layout
        print("%-12s %-7s %8.3f ms  %9.0f periods/s  path %s block %d threads %d" % (
            wl, "dy[k]" if hetero else "dy=None", min(ms), len(per) / min(ms) * 1e3, lay["path"], lay["block"], lay["threads"]), flush=True)
        s.close()
