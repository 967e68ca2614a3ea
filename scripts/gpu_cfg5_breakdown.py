#!/usr/bin/env python
"""cfg-5 (three planets on the 4-yr curve, mask + rerun x3): where the wall clock goes, per section.
usage: python scripts/gpu_cfg5_breakdown.py [workload]"""
import os, sys, time, warnings, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import numpy as np
from tls_b200 import search_planets, workloads
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
planets = [7.1, 23.4, 101.7] if wl == "cfg2" else [4.1, 9.4, 17.7]
t, y, dy, kw = workloads.lightcurve(wl, planets=planets)
for rep in range(2):
    tm = []
    t0 = time.perf_counter(); found = search_planets(t, y, n_planets=3, timings=tm, **kw); wall = time.perf_counter() - t0
print(wl, "search_planets wall %.3f s" % wall, [round(float(r.period), 4) for r in found])
keys = sorted(set(k for d in tm for k in d))
for k in keys:
    print("  %-11s %s  sum %.3f s (%.0f%%)" % (k, " ".join("%.3f" % d.get(k, 0) for d in tm), sum(d.get(k, 0) for d in tm), 100 * sum(d.get(k, 0) for d in tm) / wall))
