#!/usr/bin/env python
"""Where the wall clock of .power() goes (cProfile, top entries).  usage: python scripts/gpu_power_breakdown.py [workload]"""
import cProfile, os, pstats, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
from tls_b200 import transitleastsquares, workloads
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
t, y, dy, kw = workloads.lightcurve(wl)
m = transitleastsquares(t, y, dy, verbose=False)
m.power(show_progress_bar=False, verbose=False, **kw)
pr = cProfile.Profile(); pr.enable(); t0 = time.perf_counter()
m.power(show_progress_bar=False, verbose=False, **kw)
wall = time.perf_counter() - t0; pr.disable()
print(wl, "power() wall %.3f s" % wall)
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
