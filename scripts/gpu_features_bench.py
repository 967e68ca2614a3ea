#!/usr/bin/env python
"""Wall-clock numbers of the pieces around the period search (run on the B200 box):
.power() end to end, the T0-fit kernel, batch_power (cfg-4 shape) and search_planets (cfg-5 shape).
usage: python scripts/gpu_features_bench.py [out.json] [--big]"""
import json, os, sys, time, warnings
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
warnings.simplefilter("ignore")
from tls_b200 import batch_power, native, search_planets, stats, transitleastsquares, workloads

out = {}
big = "--big" in sys.argv
for wl in ["cfg1", "cfg3"] + (["cfg2"] if big else []):
    t, y, dy, kw = workloads.lightcurve(wl)
    m = transitleastsquares(t, y, dy, verbose=False)
    best = 1e9
    for rep in range(3 if wl != "cfg2" else 2):
        t0 = time.perf_counter(); res = m.power(show_progress_bar=False, verbose=False, **kw); best = min(best, time.perf_counter() - t0)
    out["power_wall_s_" + wl] = best
    out["power_result_" + wl] = dict(SDE=float(res.SDE), period=float(res.period), T0=float(res.T0), n_periods=len(res.periods))
    print(wl, "power() %.3f s" % best, out["power_result_" + wl], flush=True)

# T0-fit kernel alone (cfg-1: 4320 trial epochs)
t, y, dy, kw = workloads.lightcurve("cfg1")
inp = transitleastsquares(t, y, dy, verbose=False).prepare(verbose=False, **kw)
s = native.Searcher(); s.set_inputs(inp.t, inp.y, inp.dy, inp.templates, inp.params)
model_in, trials = stats.t0_fit_inputs(inp.lc_arr[20], 0.9999, inp.t, inp.y, 10.123, 0.01)
for _ in range(3):
    s.final_t0_fit(model_in, 10.123, trials)
out["t0_fit_kernel_ms_cfg1"] = s.t0_fit_ms; out["t0_fit_trials_cfg1"] = len(trials)
print("T0 fit kernel %.3f ms for %d trials" % (s.t0_fit_ms, len(trials)), flush=True)
s.close()

# cfg-4 shape: B curves shaped as cfg-1, own planets and noise
B = 256 if big else 64
rng = np.random.RandomState(1000)
t = np.linspace(3.14, 93.14, 4320)
ys = np.empty((B, len(t)))
for c in range(B):
    per = rng.uniform(1, 40); ppm = 10 ** rng.uniform(np.log10(50), np.log10(500))
    ys[c] = workloads.inject(t, per, 3.14 + rng.uniform(0, per)) + rng.normal(0, ppm * 1e-6, len(t))
batch_power(t, ys[:2])  # warm-up (template cache, buffers)
t0 = time.perf_counter(); res = batch_power(t, ys); wall = time.perf_counter() - t0
P = len(res.periods)
out["batch"] = dict(curves=B, periods=P, wall_s=wall, curves_per_s=B / wall, periods_per_s=B * P / wall, median_SDE=float(np.median(res.SDE)))
print("batch_power: %d curves x %d periods in %.3f s = %.1f curves/s, %.3g periods/s" % (B, P, wall, B / wall, B * P / wall), flush=True)

# cfg-5 shape: three planets, three successive searches
wl = "cfg2" if big else "cfg1"
t, y, dy, kw = workloads.lightcurve(wl, planets=[7.1, 23.4, 101.7] if big else [4.1, 9.4, 17.7])
t0 = time.perf_counter(); found = search_planets(t, y, n_planets=3, **kw); wall = time.perf_counter() - t0
out["multi_planet"] = dict(workload=wl, wall_s=wall, periods=[float(r.period) for r in found], SDE=[float(r.SDE) for r in found])
print("search_planets(%s): %.3f s" % (wl, wall), out["multi_planet"], flush=True)
if len(sys.argv) > 1 and not sys.argv[1].startswith("--"):
    json.dump(out, open(sys.argv[1], "w"), indent=1)
