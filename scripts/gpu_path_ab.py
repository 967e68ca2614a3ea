#!/usr/bin/env python
"""Kernel time of one workload through each layout that can hold it.  usage: python scripts/gpu_path_ab.py workload [hetero]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tls_b200 import native, transitleastsquares, workloads
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
hetero = len(sys.argv) > 2 and sys.argv[2] == "hetero"
t, y, dy, kw = workloads.lightcurve(wl, hetero=hetero)
inp = transitleastsquares(t, y, dy, verbose=False).prepare(verbose=False, **kw)
for path in ("auto", "resident", "tiled", "streaming"):
    s = native.Searcher()
    try:
        s.set_inputs(inp.t, inp.y, inp.dy, inp.templates, inp.params)
        s.set_periods(inp.periods)
        s.set_path(path)
        ms = []
        for _ in range(4):
            s.search_async(); s.results(); ms.append(s.kernel_ms)
        lay = s.layout
        print("%-12s %-6s forced %-9s %8.3f ms  %s" % (wl, "dy[k]" if hetero else "none", path, min(ms),
              {k: lay[k] for k in ("path", "threads", "ctas_per_sm", "chunk", "block", "tiled_widths")}), flush=True)
    except Exception as exc:
        print(wl, path, "->", str(exc)[:100])
    finally:
        s.close()
