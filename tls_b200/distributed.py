"""One process per GPU: shard the trial periods, search locally, all-gather the records.

The reference parallelises the period loop with a process pool (main.py:140-163); periods
are independent, so here period ``k`` goes to rank ``k mod world`` (interleaved: the cost per
period drifts along the grid) and the only exchange is ONE all-gather of the
per-period records ``{chi2 f64, depth f64, row i32 | t0_index i32}`` (24 bytes) at the end —
NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests of this host logic.  There is no
reduction: the SDE needs the whole chi2 array (stats.py:105-107).

``torch`` is plumbing here (device buffers, streams, ``torch.distributed``); the search
itself is the CUDA library behind ``include/tlsb200.h``.
"""
from __future__ import annotations

import numpy as np

WORDS_PER_PERIOD = 3  # chi2 | depth | packed(row, t0 index): three planes of 8-byte words,
                      # followed by ONE status word per shard (include/tlsb200.h: tlsb_search_async)


def record_words(capacity):
    """8-byte words of one shard's record buffer."""
    return WORDS_PER_PERIOD * capacity + 1


def shard_indices(n_periods, rank, world):
    """Indices of the periods rank ``rank`` searches (interleaved partition)."""
    return np.arange(rank, n_periods, world)


def shard_capacity(n_periods, world):
    """Record slots per rank in the gathered buffer (the largest shard)."""
    return (n_periods + world - 1) // world


def pack_records(chi2, row, depth, t0_index, capacity):
    """Host-side mirror of the kernel's record layout: int64[3 * capacity], three planes with
    the plane stride equal to the shard's own period count (as the kernel writes them)."""
    n = len(chi2)
    out = np.zeros(record_words(capacity), dtype=np.int64)
    out[0:n] = np.ascontiguousarray(chi2, np.float64).view(np.int64)
    out[n:2 * n] = np.ascontiguousarray(depth, np.float64).view(np.int64)
    packed = (np.asarray(row, np.int64) & 0xFFFFFFFF) | (np.asarray(t0_index, np.int64) << 32)
    out[2 * n:3 * n] = packed
    return out


def unpack_gathered(gathered, n_periods, world):
    """Undo the interleaved partition: ``gathered`` is int64[world * (3 * capacity + 1)]
    (rank-major).  Returns (chi2, row, depth, t0_index) in the order of the job's period list."""
    cap = shard_capacity(n_periods, world)
    g = np.asarray(gathered, dtype=np.int64).reshape(world, record_words(cap))
    if n_periods == cap * world:  # equal shards: three transposes instead of a loop over the ranks
        planes = g[:, : WORDS_PER_PERIOD * cap].reshape(world, WORDS_PER_PERIOD, cap)
        chi2 = np.ascontiguousarray(planes[:, 0, :].T).reshape(-1).view(np.float64)
        depth = np.ascontiguousarray(planes[:, 1, :].T).reshape(-1).view(np.float64)
        packed = np.ascontiguousarray(planes[:, 2, :].T).reshape(-1)
        return chi2, packed & 0xFFFFFFFF, depth, packed >> 32
    chi2 = np.empty(n_periods, np.float64)
    depth = np.empty(n_periods, np.float64)
    row = np.empty(n_periods, np.int64)
    t0 = np.empty(n_periods, np.int64)
    for r in range(world):  # rank r holds periods r, r + world, ...: strided slices, no index arrays
        n = len(range(r, n_periods, world))
        chi2[r::world] = g[r, 0:n].view(np.float64)
        depth[r::world] = g[r, n:2 * n].view(np.float64)
        packed = g[r, 2 * n:3 * n]
        row[r::world] = packed & 0xFFFFFFFF
        t0[r::world] = packed >> 32
    return chi2, row, depth, t0


def gathered_status(gathered, n_periods, world):
    """Sum of the shards' status words (non-zero: redo the search with the exact host plan)."""
    cap = shard_capacity(n_periods, world)
    g = np.asarray(gathered, dtype=np.int64).reshape(world, record_words(cap))
    return int(sum(g[r, WORDS_PER_PERIOD * len(shard_indices(n_periods, r, world))] for r in range(world)))


def all_gather_records(local_records, dist, world):
    """One collective: every rank ends with all ranks' record buffers (rank-major)."""
    import torch

    out = torch.empty(world * local_records.numel(), dtype=local_records.dtype, device=local_records.device)
    dist.all_gather_into_tensor(out, local_records)
    return out


def all_gather_interleaved(local_values, n_total, dist, device=None):
    """``local_values`` holds elements rank, rank + world, ... of an array of ``n_total`` float64; ONE all-gather
    returns the whole array in its original order on every rank (NCCL: through the GPU; gloo: on the CPU)."""
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    cap = shard_capacity(n_total, world)
    pad = np.full(cap, np.nan)
    pad[: len(local_values)] = local_values
    if dist.get_backend() == "gloo":
        dev = "cpu"
    else:
        dev = "cuda:%d" % (torch.cuda.current_device() if device is None else device)
    src = torch.from_numpy(pad).to(dev)
    out = torch.empty(world * cap, dtype=src.dtype, device=dev)
    dist.all_gather_into_tensor(out, src)
    out = out.cpu().numpy().reshape(world, cap)
    full = np.empty(n_total)
    for r in range(world):
        n = len(range(r, n_total, world))
        full[r::world] = out[r, :n]
    return full


class ShardedSearch(object):
    """The period search of one light curve on ``world`` GPUs, one process per GPU.

    ``step(stream)`` launches this rank's share (plan + search kernels) on ``stream`` and, when
    ``world > 1``, the all-gather of the records behind it; everything is asynchronous.  With
    ``world > 1`` pass torch's CURRENT stream of the device (or nothing, on the default stream): the
    collective is ordered behind the current stream, not behind an arbitrary one."""

    def __init__(self, t, y, dy, templates, params, periods, rank=0, world=1, device=0, dist=None):
        import torch

        from . import native

        self.rank, self.world, self.dist = rank, world, dist
        self.all_periods = np.ascontiguousarray(periods, np.float64)
        self.index = shard_indices(len(self.all_periods), rank, world)
        self.local_periods = np.ascontiguousarray(self.all_periods[self.index])
        self.n_local = len(self.local_periods)
        self.capacity = shard_capacity(len(self.all_periods), world)
        self.searcher = native.Searcher.acquire(device=device)
        self.searcher.set_inputs(t, y, dy, templates, params)
        self.searcher.set_periods(self.local_periods)
        self.records = torch.zeros(record_words(self.capacity), dtype=torch.int64, device="cuda:%d" % device)
        self.gathered = self.final = self._final_host = None
        self._stream = None  # the stream of the most recent step(): what a read of the records has to wait for
        if world > 1:
            self.gathered = torch.empty(world * self.records.numel(), dtype=torch.int64, device=self.records.device)
            # the job's records in period order (3 planes + the summed status word), un-interleaved on the device
            self.final = torch.empty(WORDS_PER_PERIOD * len(self.all_periods) + 1, dtype=torch.int64, device=self.records.device)
            self._final_host = torch.empty(self.final.numel(), dtype=torch.int64).pin_memory()

    def step(self, stream=None):
        ptr = stream.cuda_stream if stream is not None else 0
        self._stream = stream
        self.searcher.search_async(stream=ptr, records_ptr=self.records.data_ptr())
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.gathered, self.records)

    def _wait(self):
        """Order the host behind the stream the last step ran on (it need not be torch's current stream)."""
        import torch

        if self._stream is not None:
            self._stream.synchronize()
        else:
            torch.cuda.default_stream(self.records.device).synchronize()

    @property
    def launch_count(self):
        return self.searcher.launch_count

    @property
    def kernel_ms(self):
        return self.searcher.kernel_ms

    @property
    def resident(self):
        return self.searcher.resident

    def local_results(self):
        """This rank's (chi2, row, depth, t0_index), in the order of ``local_periods``."""
        from . import native

        self._wait()
        rec = self.records.cpu().numpy()
        if rec[WORDS_PER_PERIOD * self.n_local] != 0:
            rec = self._redo_exact(lambda: self.records.cpu().numpy())
        return native.unpack_records(rec, self.n_local)

    def results(self):
        """All periods' (chi2, row, depth, t0_index) in the job's period order (every rank).  The gathered shards are
        un-interleaved ON THE DEVICE (``tlsb_unshard_records``, main.py:190-196) and come back in one copy of
        24 bytes per period + the summed status word."""
        from . import native

        if self.world == 1:
            return self.local_results()
        n = len(self.all_periods)
        rec = self._fetch_final()
        if rec[WORDS_PER_PERIOD * n] != 0:  # every rank sees the same sum of flags: a collective decision
            rec = self._redo_exact(self._fetch_final)
        return native.unpack_records(rec, n)

    def _fetch_final(self):
        """Un-interleave on the device, then ONE device-to-host copy into pinned memory."""
        import torch

        from . import native

        dev = self.records.device
        cur = torch.cuda.current_stream(dev)
        if self._stream is not None and self._stream != cur:
            cur.wait_stream(self._stream)
        native.unshard_records(self.gathered.data_ptr(), len(self.all_periods), self.world, self.final.data_ptr(),
                               stream=cur.cuda_stream)
        with torch.cuda.stream(cur):
            self._final_host.copy_(self.final, non_blocking=True)
        cur.synchronize()
        return self._final_host.numpy()

    def reload(self, t, y, dy, templates, params, stream=None):
        """New inputs from HOST buffers for the next ``step`` (the end-to-end path): light curve, template bank and this
        rank's periods go up again in ONE call of the C ABI, asynchronously on the stream the step will use.  The
        buffers must stay unchanged until the step's results have been read."""
        ptr = stream.cuda_stream if stream is not None else 0
        self.searcher.set_inputs_async(t, y, dy, templates, params, self.local_periods, stream=ptr)

    def _redo_exact(self, fetch):
        """The device plan flagged T14 limits too close to an integer: every rank settles its own
        flagged periods on the host and re-searches only those that changed
        (include/tlsb200.h: tlsb_resolve_plan); then the records are gathered again."""
        import torch

        stream = torch.cuda.current_stream(self.records.device)
        self.searcher.resolve_plan(stream=stream.cuda_stream, records_ptr=self.records.data_ptr())
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.gathered, self.records)
        return fetch()

    def gather_host(self, local_out):
        """All-gather host-side results (the e2e path: results already copied back)."""
        import torch

        chi2, row, depth = local_out[:3]
        t0 = local_out[3] if len(local_out) > 3 else np.full(len(chi2), -1, np.int64)
        rec = torch.from_numpy(pack_records(chi2, row, depth, t0, self.capacity)).to(self.records.device, non_blocking=False)
        out = all_gather_records(rec, self.dist, self.world)
        return unpack_gathered(out.cpu().numpy(), len(self.all_periods), self.world)

    def close(self):
        self.searcher.release()


def search_periods_distributed(t, y, dy, periods, templates, params, dist, device=None):
    """Drop-in for the period loop when ``torch.distributed`` is initialised with one process
    per GPU: returns the full (chi2, row, depth) on every rank."""
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    if device is None:
        device = torch.cuda.current_device()
    job = ShardedSearch(t, y, dy, templates, params, periods, rank=rank, world=world, device=device, dist=dist)
    try:
        job.step(torch.cuda.current_stream(device))
        chi2, row, depth, _ = job.results()
    finally:
        job.close()
    return chi2, row, depth
