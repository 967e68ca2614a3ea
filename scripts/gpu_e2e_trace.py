#!/usr/bin/env python
"""Where the end-to-end call spends its time: Python marshalling vs the stages of tlsb_search_periods (TLSB_TRACE=1).
usage: TLSB_TRACE=1 scripts/gpu_e2e_trace.py [workload]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tls_b200 import native, transitleastsquares, workloads

name = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
t, y, dy, kw = workloads.lightcurve(name)
inp = transitleastsquares(t, y, dy, verbose=False).prepare(verbose=False, **kw)
pin = {}
for k, arr in (("t", inp.t), ("y", inp.y), ("dy", inp.dy), ("periods", inp.periods)):
    tt = torch.from_numpy(np.ascontiguousarray(arr, np.float64).copy()).pin_memory()
    pin[k] = (tt, tt.numpy())


def call():
    return native.search_periods(pin["t"][1], pin["y"][1], pin["dy"][1], pin["periods"][1], inp.templates, inp.params, devices=[0])


for _ in range(5):
    call()
torch.cuda.synchronize()
n = 20
t0 = time.perf_counter()
for _ in range(n):
    call()
torch.cuda.synchronize()
print("e2e call: %.1f us per call" % (1e6 * (time.perf_counter() - t0) / n))
t0 = time.perf_counter()
for _ in range(200):
    pk = native._Packed(pin["t"][1], pin["y"][1], pin["dy"][1], inp.templates, inp.params)
print("python _Packed: %.1f us" % (1e6 * (time.perf_counter() - t0) / 200))
