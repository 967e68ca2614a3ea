#!/usr/bin/env python
"""Filter-pass counters per workload: gate survivors, finalists (fp64 evaluations), queue overflows, kernel ms.
usage: scripts/gpu_filter_stats.py [workloads...]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tls_b200 import native, transitleastsquares, workloads

for name in sys.argv[1:] or ["cfg1", "cfg1_500ppm"]:
    t, y, dy, kw = workloads.lightcurve(name)
    inp = transitleastsquares(t, y, dy, verbose=False).prepare(verbose=False, **kw)
    per = inp.periods if name != "cfg2" else inp.periods[::31]
    s = native.Searcher()
    s.set_inputs(inp.t, inp.y, inp.dy, inp.templates, inp.params)
    s.set_periods(per)
    for on in (True, False) if len(per) * len(inp.t) < 5e7 and "--off" in os.environ.get("TLSB_STATS_ARGS", "") else (True,):
        s.set_filter(on, True)
        for _ in range(2):
            s.search_async(); s.results()
        st = s.filter_stats
        P = len(per)
        print("%-12s filter=%d path %s  kernel %.3f ms  per period: survivors %.0f  finalists %.1f (%.3f%%)  overflows %.2f" % (
            name, on, s.path, s.kernel_ms, st["candidates"] / P, st["finalists"] / P,
            100.0 * st["finalists"] / max(1, st["candidates"]), st["overflows"] / P))
    s.close()
