import cProfile, pstats, sys, time, warnings, numpy as np
sys.path.insert(0, '.')
warnings.simplefilter("ignore")
from tls_b200 import batch_power, workloads
B = 64
rng = np.random.RandomState(1000)
t = np.linspace(3.14, 93.14, 4320)
ys = np.empty((B, len(t)))
for c in range(B):
    per = rng.uniform(1, 40); ppm = 10 ** rng.uniform(np.log10(50), np.log10(500))
    ys[c] = workloads.inject(t, per, 3.14 + rng.uniform(0, per)) + rng.normal(0, ppm * 1e-6, len(t))
batch_power(t, ys[:2])
t0 = time.perf_counter(); res = batch_power(t, ys); print("wall %.3f s -> %.1f curves/s" % (time.perf_counter() - t0, B / (time.perf_counter() - t0)))
pr = cProfile.Profile(); pr.enable(); res = batch_power(t, ys); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
