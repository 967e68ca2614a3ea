#!/usr/bin/env python
"""Where power(dist=...) spends its search section at world > 1 (torchrun): ShardedSearch constructor, step + wait,
results, close.  usage: torchrun --nproc-per-node N scripts/gpu_dist_breakdown.py [workload]"""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tls_b200 import transitleastsquares, workloads
from tls_b200.distributed import ShardedSearch

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
t, y, dy, kw = workloads.lightcurve(name)
inp = transitleastsquares(t, y, dy, verbose=False).prepare(verbose=False, **kw)
for it in range(3):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    job = ShardedSearch(inp.t, inp.y, inp.dy, inp.templates, inp.params, inp.periods, rank=rank, world=world, device=local, dist=dist)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    job.step(torch.cuda.current_stream(local))
    torch.cuda.synchronize(); t2 = time.perf_counter()
    out = job.results()
    t3 = time.perf_counter()
    job.close()
    t4 = time.perf_counter()
    if rank == 0:
        print("iter %d: ctor %.1f ms  step+wait %.1f ms (kernel %.1f ms)  results %.1f ms  close %.1f ms  total %.1f ms" % (
            it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), job.kernel_ms, 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t4 - t0)))
dist.destroy_process_group()
