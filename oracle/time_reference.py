#!/usr/bin/env python
"""TEST / BENCH INFRASTRUCTURE — not part of the product.

Times the reference's OWN implementation of the hot path — the unmodified numba ``core.search_period``
(core.py:96-188) — on this machine's host cores, the way the reference itself parallelises it
(main.py:141-163): a warmed ``multiprocessing.Pool(processes).imap_unordered(partial(search_period, ...))``
over the trial periods, plus a one-core serial loop (main.py:165-183).  The package is imported from
``/root/reference`` in the build container and from the unmodified copy under ``oracle/_ref/``
(``oracle/vendor_ref.py``) on the GPU box; ``batman`` comes from ``oracle/ref_shim.py`` (it only shapes the
templates, which both sides receive as the same arrays).

Runs as its own process (bench.py spawns it) so that the fork-based Pool never inherits a CUDA context or
torchrun's OMP_NUM_THREADS.  Prints one JSON object.

usage: python oracle/time_reference.py --workload cfg1 [--oversampling 3] [--seconds 10] [--procs N] [--min-periods 2000]
"""
import argparse
import json
import multiprocessing
import os
import sys
import time
import warnings
from functools import partial

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402


def spread(periods, n):
    n = int(max(1, min(len(periods), n)))
    return periods[np.linspace(0, len(periods) - 1, n).astype(int)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg1")
    ap.add_argument("--oversampling", type=int, default=3)
    ap.add_argument("--seconds", type=float, default=10.0, help="target wall time of the pooled sample")
    ap.add_argument("--serial-seconds", type=float, default=3.0)
    ap.add_argument("--procs", type=int, default=0, help="pool size (0 = os.cpu_count())")
    ap.add_argument("--min-periods", type=int, default=2000, help="lower bound of the pooled sample (if the grid has that many)")
    ap.add_argument("--max-periods", type=int, default=0, help="cap the grid first (the b200 arm's --max-periods * gpus)")
    ap.add_argument("--steps", type=int, default=1, help="repeat the pooled sample this many times (one warmed pool)")
    args = ap.parse_args()

    from oracle import ref_shim

    if not ref_shim.available():
        print(json.dumps({"unavailable": "reference package neither at /root/reference nor vendored under oracle/_ref"}))
        return 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ref_shim.load()
        from transitleastsquares import core as ref_core  # the reference's own module

        from tls_b200 import transitleastsquares as host, workloads

        t, y, dy, kw = workloads.lightcurve(args.workload)
        kw = dict(kw)
        kw["oversampling_factor"] = args.oversampling
        inp = host(t, y, dy, verbose=False).prepare(verbose=False, **kw)
    periods = inp.periods
    if args.max_periods and args.max_periods < len(periods):
        periods = spread(periods, args.max_periods)
    prm = inp.params
    call = partial(
        ref_core.search_period, t=inp.t, y=inp.y, dy=inp.dy, transit_depth_min=prm["transit_depth_min"],
        R_star_min=prm["R_star_min"], R_star_max=prm["R_star_max"], M_star_min=prm["M_star_min"],
        M_star_max=prm["M_star_max"], lc_arr=inp.lc_arr, lc_cache_overview=inp.overview,
        T0_fit_margin=prm["T0_fit_margin"])

    # numba JIT in the parent, before any fork: the children inherit the compiled code
    t0 = time.perf_counter()
    call(float(periods[len(periods) // 2]))
    jit_s = time.perf_counter() - t0

    # one core, serial loop (main.py:165-183)
    probe = spread(periods, 16)
    t0 = time.perf_counter()
    for p in probe:
        call(float(p))
    rate1 = len(probe) / (time.perf_counter() - t0)
    serial = spread(periods, max(16, rate1 * args.serial_seconds))
    t0 = time.perf_counter()
    for p in serial:
        call(float(p))
    serial_rate = len(serial) / (time.perf_counter() - t0)

    # all cores: warmed fork pool, imap_unordered (main.py:141-163)
    procs = args.procs or (os.cpu_count() or 1)
    ctx = multiprocessing.get_context("fork")
    with ctx.Pool(processes=procs) as pool:
        warm = spread(periods, procs * 8)
        for _ in pool.imap_unordered(call, [float(p) for p in warm]):
            pass
        t0 = time.perf_counter()
        for _ in pool.imap_unordered(call, [float(p) for p in warm]):
            pass
        est = len(warm) / (time.perf_counter() - t0)
        n = int(min(len(periods), max(args.min_periods, est * args.seconds)))
        sample = [float(p) for p in spread(periods, n)]
        step_s = []
        for _ in range(max(1, args.steps)):
            t0 = time.perf_counter()
            got = 0
            for _ in pool.imap_unordered(call, sample):
                got += 1
            step_s.append(time.perf_counter() - t0)
            assert got == len(sample)
    import numba

    out = {
        "kind": "reference", "impl": "unmodified transitleastsquares.core.search_period (numba %s, numpy %s)" % (
            numba.__version__, np.__version__),
        "imported_from": ref_shim.reference_root(), "version": getattr(ref, "__version__", None) or "1.0.31",
        "workload": args.workload, "n_points": int(len(inp.y)), "periods_in_grid": int(len(periods)),
        "pool": {"value": len(sample) * len(step_s) / float(sum(step_s)), "unit": "periods/s", "cores": procs,
                 "periods_per_step": len(sample), "steps": len(step_s), "ms_per_step": 1e3 * float(np.mean(step_s)),
                 "how": "warmed multiprocessing.Pool(%d).imap_unordered(partial(search_period, ...)), main.py:141-163" % procs},
        "serial": {"value": serial_rate, "unit": "periods/s", "cores": 1, "periods": len(serial), "how": "serial loop, main.py:165-183"},
        "jit_first_call_s": jit_s, "host_cpu_count": os.cpu_count(),
    }
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
