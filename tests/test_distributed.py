"""Host logic of the one-process-per-GPU path (tls_b200/distributed.py) on CPU: interleaved
period partition, the kernel's record layout, ONE all-gather (gloo, world_size 2 and 3), and the
un-interleave.  The per-rank search is stood in by the oracle (test infrastructure)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import REPO, load_search_golden
from tls_b200 import distributed as D


def test_partition_covers_every_period_once():
    for n, world in ((10, 1), (10, 3), (9679, 8), (5, 8)):
        seen = np.concatenate([D.shard_indices(n, r, world) for r in range(world)])
        assert sorted(seen) == list(range(n))
        assert max(len(D.shard_indices(n, r, world)) for r in range(world)) == D.shard_capacity(n, world)


def test_pack_unpack_round_trip():
    rng = np.random.RandomState(0)
    n, world = 1001, 4
    chi2 = rng.rand(n) * 1e4
    chi2[3] = np.inf
    depth = rng.rand(n)
    row = rng.randint(0, 200, n)
    t0 = rng.randint(-1, 70000, n)
    cap = D.shard_capacity(n, world)
    shards = []
    for r in range(world):
        idx = D.shard_indices(n, r, world)
        shards.append(D.pack_records(chi2[idx], row[idx], depth[idx], t0[idx], cap))
    got = D.unpack_gathered(np.concatenate(shards), n, world)
    for a, b in zip(got, (chi2, row, depth, t0)):
        np.testing.assert_array_equal(a, b)
    assert D.gathered_status(np.concatenate(shards), n, world) == 0
    shards[2][3 * len(D.shard_indices(n, 2, world))] = 5  # a shard flags 5 uncertain periods
    assert D.gathered_status(np.concatenate(shards), n, world) == 5


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import torch
    import torch.distributed as dist

    from oracle import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_search_golden("small")
    n = len(g["periods"])
    idx = D.shard_indices(n, rank, world)
    chi2, row, depth = oracle.search_periods_c(g["t"], g["y"], g["dy"], g["periods"][idx], g["templates"], g["params"], threads=1)
    rec = torch.from_numpy(D.pack_records(chi2, row, depth, np.full(len(idx), -1), D.shard_capacity(n, world)))
    gathered = D.all_gather_records(rec, dist, world)  # the one collective of the search
    full = D.unpack_gathered(gathered.numpy(), n, world)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), chi2=full[0], row=full[1], depth=full[2])
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_over_gloo_equals_single_process(world, tmp_path):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = load_search_golden("small")
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        np.testing.assert_array_equal(z["row"], g["row"])
        np.testing.assert_allclose(z["chi2"], g["chi2"], rtol=1e-9)
        np.testing.assert_allclose(z["depth"], g["depth"], rtol=1e-9)


def _batch_worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    import torch.distributed as dist

    from tls_b200 import batch

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_rows = 7
    rows = np.full((n_rows, 3), np.nan)
    mine = batch.shard_curves(n_rows, rank, world)
    rows[mine] = np.arange(n_rows * 3, dtype=float).reshape(n_rows, 3)[mine]  # what this rank computed
    full = batch._all_gather_rows(rows, mine, n_rows, dist, None)
    np.save(os.path.join(out_dir, "batch_rank%d.npy" % rank), full)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_batch_summaries_all_gather_over_gloo(world, tmp_path):
    """Curves are dealt to the ranks round-robin; ONE all-gather returns every curve's row to every rank."""
    import torch.multiprocessing as mp

    from tls_b200 import batch

    seen = np.concatenate([batch.shard_curves(7, r, world) for r in range(world)])
    assert sorted(seen) == list(range(7))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_batch_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    want = np.arange(21, dtype=float).reshape(7, 3)
    for rank in range(world):
        np.testing.assert_array_equal(np.load(os.path.join(str(tmp_path), "batch_rank%d.npy" % rank)), want)


def _t0_worker(rank, world, port, out_dir):
    """The sharded final_T0_fit (stats.py:135-204 with trial epoch k -> rank k mod world): the per-rank residuals come
    from the numpy oracle here (no GPU), the all-gather and the strict-'<' scan are the product's."""
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import torch.distributed as dist

    from conftest import load_t0fit_golden
    from oracle import oracle
    from tls_b200 import native, stats

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_t0fit_golden("small_margin0")

    dy = np.full(len(g["y"]), np.std(g["y"]))

    def fake_device_fit(t, y, dy, model_in, period, trials, device=None):  # what tlsb_final_t0_fit_lc returns for these trials
        _, resid, _ = oracle.final_T0_fit_numpy(g["signal"], float(g["depth"]), t, y, dy, period, float(g["margin"]), trials=trials)
        return int(np.argmin(resid)), resid

    native.final_t0_fit, keep = fake_device_fit, native.final_t0_fit
    try:
        T0 = stats.final_T0_fit(g["signal"], float(g["depth"]), g["t"], g["y"], dy, float(g["period"]),
                                float(g["margin"]), False, False, dist=dist)
    finally:
        native.final_t0_fit = keep
    np.save(os.path.join(out_dir, "t0_rank%d.npy" % rank), np.array([T0]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_t0_fit_over_gloo_gives_the_reference_epoch(world, tmp_path):
    import torch.multiprocessing as mp

    from conftest import load_t0fit_golden

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_t0_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    want = float(load_t0fit_golden("small_margin0")["T0"])
    for rank in range(world):
        assert float(np.load(os.path.join(str(tmp_path), "t0_rank%d.npy" % rank))[0]) == want
