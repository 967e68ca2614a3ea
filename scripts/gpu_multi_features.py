#!/usr/bin/env python
"""cfg-4 and cfg-5 shapes on N GPUs, one process per GPU (run under torchrun on the B200 box):
batch_power(dist=...) on curve shards and search_planets(dist=...) on period shards, each checked
against the same call on one GPU (rank 0).
usage: python -m torch.distributed.run --nproc-per-node N scripts/gpu_multi_features.py [out.json] [--big]"""
import json, os, sys, time, warnings
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
warnings.simplefilter("ignore")
import torch
import torch.distributed as dist

from tls_b200 import batch_power, search_planets, workloads

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
big = "--big" in sys.argv
out = {"world": world}

# cfg-4 shape: B curves shaped as cfg-1, own planets and noise; curve c -> rank c mod world
B = 1000 if big else 64 * world
rng = np.random.RandomState(1000)
t = np.linspace(3.14, 93.14, 4320)
ys = np.empty((B, len(t)))
for c in range(B):
    per = rng.uniform(1, 40); ppm = 10 ** rng.uniform(np.log10(50), np.log10(500))
    ys[c] = workloads.inject(t, per, 3.14 + rng.uniform(0, per)) + rng.normal(0, ppm * 1e-6, len(t))
batch_power(t, ys[: 2 * world], dist=dist, device=local)  # warm-up
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter(); res = batch_power(t, ys, dist=dist, device=local); torch.cuda.synchronize(); dist.barrier()
wall = time.perf_counter() - t0
P = len(res.periods)
out["batch"] = dict(curves=B, periods=P, wall_s=wall, curves_per_s=B / wall, periods_per_s=B * P / wall,
                    median_SDE=float(np.median(res.SDE)))
if rank == 0:
    n_chk = min(B, 16)
    one = batch_power(t, ys[:n_chk], device=local)
    out["batch"]["max_rel_SDE_diff_vs_one_gpu"] = float(np.max(np.abs(one.SDE - res.SDE[:n_chk]) / np.abs(one.SDE)))
    out["batch"]["periods_equal_vs_one_gpu"] = bool(np.array_equal(one.period, res.period[:n_chk]))
    print("batch_power on %d GPUs: %d curves x %d periods in %.3f s = %.1f curves/s, %.3g periods/s; vs one GPU: %s" % (
        world, B, P, wall, B / wall, B * P / wall, {k: v for k, v in out["batch"].items() if "vs_one" in k}), flush=True)
dist.barrier()

# cfg-5 shape: three planets, three successive searches, periods sharded over the ranks
wl = "cfg2" if big else "cfg1"
planets = [7.1, 23.4, 101.7] if big else [4.1, 9.4, 17.7]
t, y, dy, kw = workloads.lightcurve(wl, planets=planets)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter(); found = search_planets(t, y, n_planets=3, dist=dist, device=local, **kw); torch.cuda.synchronize(); dist.barrier()
wall = time.perf_counter() - t0
out["multi_planet"] = dict(workload=wl, wall_s=wall, periods=[float(r.period) for r in found], SDE=[float(r.SDE) for r in found])
if rank == 0:
    t1 = time.perf_counter(); one = search_planets(t, y, n_planets=3, device=local, **kw); wall1 = time.perf_counter() - t1
    out["multi_planet"]["one_gpu_wall_s"] = wall1
    out["multi_planet"]["same_as_one_gpu"] = bool(len(one) == len(found) and all(
        a.period == b.period and abs(a.SDE - b.SDE) <= 1e-9 * abs(b.SDE) and a.T0 == b.T0 for a, b in zip(found, one)))
    print("search_planets(%s) on %d GPUs: %.3f s (one GPU %.3f s)" % (wl, world, wall, wall1), out["multi_planet"], flush=True)
    if len(sys.argv) > 1 and not sys.argv[1].startswith("--"):
        json.dump(out, open(sys.argv[1], "w"), indent=1)
dist.barrier()
dist.destroy_process_group()
