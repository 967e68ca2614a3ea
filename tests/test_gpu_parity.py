"""Parity of the CUDA path (through the C ABI) with the reference, on the committed golden
vectors (reference numba outputs) and against the C oracle on seeded inputs.

Bar (BASELINE.json north_star): argmin rows bit-exact, chi2 and depth within 1e-5 relative."""
import numpy as np
import pytest

from conftest import assert_search_parity, load_search_golden, search_goldens

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # the stated tolerance; the kernels are fp64 end to end and land near 1e-12


def _native():
    from tls_b200 import native

    assert native.device_count() > 0, "no CUDA device: the GPU tests must not fall back to anything"
    return native


@pytest.mark.parametrize("name", search_goldens())
def test_cuda_matches_reference_golden(name):
    native = _native()
    g = load_search_golden(name)
    got = native.search_periods(g["t"], g["y"], g["dy"], g["periods"], g["templates"], g["params"])
    assert_search_parity(got, g, rtol=RTOL, label=name)
    # and report how close we actually are
    fin = np.isfinite(g["chi2"]) & (g["chi2"] != len(g["y"]))
    if fin.any():
        err = np.max(np.abs(got[0][fin] - g["chi2"][fin]) / g["chi2"][fin])
        assert err < 1e-9, "fp64 path drifted: %g" % err


def test_handle_api_equals_one_shot_and_is_repeatable():
    native = _native()
    g = load_search_golden("small")
    one = native.search_periods(g["t"], g["y"], g["dy"], g["periods"], g["templates"], g["params"], return_t0_index=True)
    s = native.Searcher()
    s.set_inputs(g["t"], g["y"], g["dy"], g["templates"], g["params"])
    s.set_periods(g["periods"])
    counts = []
    for _ in range(3):  # the scheduler counter must reset itself between launches
        s.search_async()
        counts.append(s.launch_count)
        res = s.results()
        for a, b in zip(one, res):
            np.testing.assert_array_equal(a, b)
    # plan + search; once results() has read the plan's status word back clean, the handle keeps the device plan for
    # as long as periods, bank, N, span and stellar limits stay the same (TLSB_MEMO=0 switches that off)
    assert counts == [2, 1, 1]
    assert s.kernel_ms > 0
    s.close()


@pytest.mark.parametrize("name", ["small", "no_admissible", "cfg1_hetero"])
def test_device_plan_equals_exact_host_plan_and_repair(name):
    """T14 limits on the device (plan kernel) vs the exact host plan (libm pow), and the repair of
    periods the device flags as too close to an integer: mode 2 flags every period (nothing differs,
    nothing is searched again), mode 3 also makes every 7th device range wrong (those periods are
    searched again with the exact range)."""
    native = _native()
    g = load_search_golden(name)
    s = native.Searcher()
    s.set_inputs(g["t"], g["y"], g["dy"], g["templates"], g["params"])
    s.set_periods(g["periods"])
    out = {}
    repairs = {}
    for mode in (0, 1, 2, 3):
        s.set_plan_mode(mode)
        s.search_async()
        out[mode] = s.results()
        repairs[mode] = s.plan_repairs
    assert repairs[0] == repairs[1] == repairs[2] == 0
    if name != "no_admissible":
        assert repairs[3] >= len(g["periods"]) // 8   # every 7th period with an admissible width
    assert s.plan_fallbacks == (1 if len(g["periods"]) > 4096 else 0)  # more flags than the kernel lists: whole exact plan
    for mode in (1, 2, 3):
        for a, b in zip(out[0], out[mode]):
            np.testing.assert_array_equal(a, b)
    assert_search_parity(out[0][:3], g, rtol=RTOL, label=name)
    s.close()


def test_resolve_plan_on_a_caller_owned_record_buffer():
    """The asynchronous API: a sabotaged device plan, then tlsb_resolve_plan on the caller's buffer."""
    import torch

    native = _native()
    g = load_search_golden("small")
    P = len(g["periods"])
    s = native.Searcher()
    s.set_inputs(g["t"], g["y"], g["dy"], g["templates"], g["params"])
    s.set_periods(g["periods"])
    rec = torch.zeros(3 * P + 1, dtype=torch.int64, device="cuda")
    s.set_plan_mode(3)
    s.search_async(records_ptr=rec.data_ptr())
    torch.cuda.synchronize()
    assert int(rec[3 * P].item()) == P            # every period flagged
    wrong = native.unpack_records(rec.cpu().numpy(), P)
    s.resolve_plan(records_ptr=rec.data_ptr())
    assert int(rec[3 * P].item()) == 0
    fixed = native.unpack_records(rec.cpu().numpy(), P)
    assert not np.array_equal(wrong[0], fixed[0])  # the sabotage did change some results ...
    assert_search_parity(fixed[:3], g, rtol=RTOL, label="resolved")   # ... and the repair restored them
    s.close()


def test_equal_weight_specialisation_matches_general_path():
    """dy=None gives equal weights and a specialised kernel; nudging one dy by one ulp forces
    the general kernel on (numerically) the same problem."""
    native = _native()
    g = load_search_golden("cfg1_500ppm")
    sel = slice(0, None, 8)
    dy2 = g["dy"].copy()
    dy2[17] = np.nextafter(dy2[17], 2.0)
    a = native.search_periods(g["t"], g["y"], g["dy"], g["periods"][sel], g["templates"], g["params"])
    b = native.search_periods(g["t"], g["y"], dy2, g["periods"][sel], g["templates"], g["params"])
    np.testing.assert_array_equal(a[1], b[1])
    np.testing.assert_allclose(a[0], b[0], rtol=1e-10)
    np.testing.assert_allclose(a[2], b[2], rtol=1e-10)


def test_period_order_does_not_matter():
    native = _native()
    g = load_search_golden("small")
    rng = np.random.RandomState(0)
    perm = rng.permutation(len(g["periods"]))
    a = native.search_periods(g["t"], g["y"], g["dy"], g["periods"], g["templates"], g["params"])
    b = native.search_periods(g["t"], g["y"], g["dy"], g["periods"][perm], g["templates"], g["params"])
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x[perm], y)


def test_empty_period_list():
    native = _native()
    g = load_search_golden("tiny")
    chi2, row, depth = native.search_periods(g["t"], g["y"], g["dy"], np.zeros(0), g["templates"], g["params"])
    assert len(chi2) == len(row) == len(depth) == 0


def test_bad_arguments_raise():
    native = _native()
    g = load_search_golden("tiny")
    with pytest.raises(RuntimeError):
        native.search_periods(g["t"][:2], g["y"][:2], g["dy"][:2], g["periods"], g["templates"], g["params"])


@pytest.mark.parametrize("workload,hetero,count", [("cfg1", False, 400), ("cfg1_500ppm", True, 300), ("cfg3", False, 40),
                                                    ("cfg1", True, 300), ("cfg3", True, 24)])
def test_cuda_matches_c_oracle_on_seeded_inputs(workload, hetero, count):
    """Same seeded inputs into the CUDA path and the CPU oracle (oracle/ is the checker only)."""
    native = _native()
    from oracle import oracle
    from tls_b200 import transitleastsquares, workloads

    t, y, dy, kw = workloads.lightcurve(workload, hetero=hetero)
    inp = transitleastsquares(t, y, dy, verbose=False).prepare(verbose=False, **kw)
    sel = np.linspace(0, len(inp.periods) - 1, count).astype(int)
    periods = inp.periods[sel]
    want = oracle.search_periods_c(inp.t, inp.y, inp.dy, periods, inp.templates, inp.params)
    got = native.search_periods(inp.t, inp.y, inp.dy, periods, inp.templates, inp.params)
    ref = dict(y=inp.y, chi2=want[0], row=want[1], depth=want[2])
    assert_search_parity(got, ref, rtol=RTOL, label=workload)


def _search_with_path(native, g, path, chunk=0):
    s = native.Searcher()
    s.set_inputs(g["t"], g["y"], g["dy"], g["templates"], g["params"])
    s.set_periods(g["periods"])
    s.set_path(path, chunk)
    s.search_async()
    out = s.results()
    info = dict(s.layout, **s.sort_info)
    s.close()
    return out, info


@pytest.mark.parametrize("name", [n for n in search_goldens() if n not in ("cfg2",)])
def test_tiled_path_matches_reference_golden(name):
    """Every golden through the TILED kernel (bulk-copy staged chunks) with a chunk capacity small
    enough that the folded curve spans several chunks; results must equal the automatic path's."""
    native = _native()
    g = load_search_golden(name)
    if len(g["periods"]) > 600:
        sel = np.linspace(0, len(g["periods"]) - 1, 600).astype(int)
        g = dict(g, periods=g["periods"][sel], chi2=g["chi2"][sel], row=g["row"][sel], depth=g["depth"][sel])
    n = len(g["y"])
    out, info = _search_with_path(native, g, "tiled", chunk=max(256, n // 3))
    assert info["path"] == "tiled"
    # the fold is sorted on chip in several phase segments; only clustered phases may fall back
    assert info["n_segments"] >= 2 and info["segment_capacity"] < n
    if name in ("small", "cfg1_50ppm", "cfg3", "k2_epic201367065"):
        assert info["global_sort_periods"] <= len(g["periods"]) // 10
    assert_search_parity(out[:3], g, rtol=RTOL, label=name + " tiled")
    auto, _ = _search_with_path(native, g, "auto")
    np.testing.assert_array_equal(out[1], auto[1])
    np.testing.assert_array_equal(out[3], auto[3])  # same best window start
    np.testing.assert_allclose(out[0], auto[0], rtol=1e-12)


@pytest.mark.parametrize("name", ["small", "cfg1_50ppm", "cfg1_hetero", "ragged_L", "ties_unsorted", "margin0", "cfg3"])
def test_tiled_path_with_widths_too_wide_for_a_chunk(name):
    """A chunk capped BELOW the widest window: the narrow widths are searched from staged chunks, the
    widest ones in the extra pass that reads the folded curve from the L2 scratch (the layout of a
    4-year curve with per-point uncertainties).  Same results as the automatic path."""
    native = _native()
    g = load_search_golden(name)
    if len(g["periods"]) > 400:
        sel = np.linspace(0, len(g["periods"]) - 1, 400).astype(int)
        g = dict(g, periods=g["periods"][sel], chi2=g["chi2"][sel], row=g["row"][sel], depth=g["depth"][sel])
    widths = np.unique(np.asarray(g["templates"]["width"]))
    cap = int(widths[len(widths) * 2 // 3] * 1.2) + 100  # roughly the widest third of the bank does not fit
    if cap >= widths[-1]:
        pytest.skip("bank too narrow to split")
    out, info = _search_with_path(native, g, "tiled", chunk=-cap)
    assert info["path"] == "tiled" and info["chunk"] <= cap
    assert 1 <= info["tiled_widths"] < len(widths), info
    assert_search_parity(out[:3], g, rtol=RTOL, label=name + " tiled+L2")
    auto, _ = _search_with_path(native, g, "auto")
    np.testing.assert_array_equal(out[1], auto[1])
    np.testing.assert_array_equal(out[3], auto[3])
    np.testing.assert_allclose(out[0], auto[0], rtol=1e-12)


@pytest.mark.parametrize("name", ["small", "cfg1_hetero", "ragged_L", "ties_unsorted", "cfg3"])
def test_streaming_path_matches_reference_golden(name):
    """The last-resort layout (everything through L1/L2) stays correct."""
    native = _native()
    g = load_search_golden(name)
    if len(g["periods"]) > 300:
        sel = np.linspace(0, len(g["periods"]) - 1, 300).astype(int)
        g = dict(g, periods=g["periods"][sel], chi2=g["chi2"][sel], row=g["row"][sel], depth=g["depth"][sel])
    out, info = _search_with_path(native, g, "streaming")
    assert info["path"] == "streaming"
    assert_search_parity(out[:3], g, rtol=RTOL, label=name + " streaming")


def test_automatic_layout_choice():
    native = _native()
    for name, want in (("small", "resident"), ("cfg1_50ppm", "resident"), ("cfg3", "tiled"), ("cfg2", "tiled")):
        g = load_search_golden(name)
        g = dict(g, periods=g["periods"][:8])
        _, info = _search_with_path(native, g, "auto")
        assert info["path"] == want, (name, info)


def test_forced_resident_path_refuses_what_does_not_fit():
    native = _native()
    g = load_search_golden("cfg2")
    s = native.Searcher()
    s.set_inputs(g["t"], g["y"], g["dy"], g["templates"], g["params"])
    s.set_periods(g["periods"][:4])
    s.set_path("resident")
    with pytest.raises(RuntimeError, match="resident"):
        s.search_async()
    s.close()


def test_on_chip_sort_equals_global_sort(monkeypatch):
    """Tiled path with the segmented on-chip sort vs the same path sorting in global scratch."""
    native = _native()
    g = load_search_golden("cfg3")
    g = dict(g, periods=g["periods"][::7])
    on, info_on = _search_with_path(native, g, "tiled")
    assert info_on["n_segments"] >= 2
    monkeypatch.setenv("TLSB_ONCHIP_SORT", "0")
    off, info_off = _search_with_path(native, g, "tiled")
    assert info_off["n_segments"] == 0
    np.testing.assert_array_equal(on[1], off[1])
    np.testing.assert_array_equal(on[3], off[3])
    np.testing.assert_allclose(on[0], off[0], rtol=1e-12)


def test_clustered_phases_fall_back_to_the_global_sort():
    """Three samples per day, trial period one day: every sample sits on one of three phases, a
    segment of the on-chip sort overflows and that period is sorted in global scratch instead -
    with the same results as the resident path."""
    native = _native()
    g = load_search_golden("small")
    n = len(g["y"])
    rng = np.random.RandomState(8)
    k = np.arange(n)
    t = (k % 3) / 3.0 + k // 3 + rng.uniform(0, 1e-7, n)
    y = 1.0 + rng.normal(0, 2e-4, n)
    g = dict(g, t=t, y=y, dy=np.full(n, np.std(y)), periods=np.array([1.0, 2.0, 3.3, 0.5, 7.7]))
    tiled, info = _search_with_path(native, g, "tiled", chunk=max(256, n // 3))
    assert info["n_segments"] >= 2
    assert 1 <= info["global_sort_periods"] <= 3, info
    auto, info2 = _search_with_path(native, g, "auto")
    assert info2["path"] == "resident"
    np.testing.assert_array_equal(tiled[1], auto[1])
    np.testing.assert_array_equal(tiled[3], auto[3])
    np.testing.assert_allclose(tiled[0], auto[0], rtol=1e-12)


@pytest.mark.parametrize("n,world", [(9679, 2), (9679, 8), (77, 3), (5, 8), (64, 4)])
def test_device_unshard_equals_the_host_unpack(n, world):
    """tlsb_unshard_records (main.py:190-196 on the device): a rank-major all-gathered buffer of interleaved shards,
    built here on one GPU exactly as ranks would write it, comes out in the job's period order; the host-side
    unpack of the gloo tests (tls_b200.distributed.unpack_gathered) is the checker.  Bit for bit."""
    import torch

    native = _native()
    from tls_b200 import distributed as D

    rng = np.random.default_rng(n * 31 + world)
    chi2 = rng.normal(4000, 10, n)
    depth = rng.uniform(0.99, 1.0, n)
    row = rng.integers(0, 60, n)
    t0 = rng.integers(-1, 5000, n)
    cap = D.shard_capacity(n, world)
    shards = []
    for r in range(world):
        idx = D.shard_indices(n, r, world)
        rec = D.pack_records(chi2[idx], row[idx], depth[idx], t0[idx], cap)
        rec[3 * len(idx)] = r + 1  # status words: the sum must come through
        shards.append(rec)
    gathered = np.concatenate(shards)
    want = D.unpack_gathered(gathered, n, world)
    g_dev = torch.from_numpy(gathered).cuda()
    out = torch.empty(3 * n + 1, dtype=torch.int64, device="cuda")
    native.unshard_records(g_dev.data_ptr(), n, world, out.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    host = out.cpu().numpy()
    got = native.unpack_records(host, n)
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    assert host[3 * n] == world * (world + 1) // 2
    np.testing.assert_array_equal(got[0], chi2)
    np.testing.assert_array_equal(got[3], t0)


def test_one_call_async_upload_equals_the_three_setters_and_get_results_guards_its_buffer():
    import torch

    native = _native()
    g = load_search_golden("small")
    s = native.Searcher()
    s.set_inputs(g["t"], g["y"], g["dy"], g["templates"], g["params"])
    s.set_periods(g["periods"])
    s.search_async()
    want = s.results()
    s2 = native.Searcher()
    stream = torch.cuda.current_stream()
    for _ in range(3):  # repeated reloads: the staged template copy must not be rewritten under an upload in flight
        s2.set_inputs_async(g["t"], g["y"], g["dy"], g["templates"], g["params"], g["periods"], stream=stream.cuda_stream)
        s2.search_async(stream=stream.cuda_stream)
        got = s2.results(stream=stream.cuda_stream)
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)
    # a search into a caller-supplied buffer must not let tlsb_get_results hand back the OLDER search in the handle's buffer
    mine = torch.zeros(3 * len(g["periods"]) + 1, dtype=torch.int64, device="cuda")
    s2.search_async(stream=stream.cuda_stream, records_ptr=mine.data_ptr())
    with pytest.raises(RuntimeError, match="records_dev|own buffer"):
        s2.results(stream=stream.cuda_stream)
    torch.cuda.synchronize()
    got = native.unpack_records(mine.cpu().numpy(), len(g["periods"]))
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)
    s.close()
    s2.close()


@pytest.mark.parametrize("hetero", [False, True])
def test_resident_layout_with_one_big_cta_per_sm(hetero):
    """A curve whose folded arrays fit shared memory once but not twice per SM (146 d @ 30 min, N = 7000) takes the
    resident kernel as ONE 512-thread CTA per SM with the barrier-free ring schedule (tlsb_device.cuh: sweep_filter) -
    with equal weights and with per-point dy.  Results against the C oracle (core.py:96-188)."""
    native = _native()
    from oracle import oracle
    from tls_b200 import transitleastsquares, workloads

    n = 7000
    t = np.linspace(3.14, 3.14 + n / 48.0, n)
    np.random.seed(11)
    y = workloads.inject(t, 7.77, 3.14) + np.random.normal(0, 300e-6, n)
    dy = 300e-6 * np.random.uniform(0.5, 2.0, n) if hetero else None
    inp = transitleastsquares(t, y, dy, verbose=False).prepare(verbose=False)
    periods = inp.periods[np.linspace(0, len(inp.periods) - 1, 60).astype(int)]
    s = native.Searcher()
    s.set_inputs(inp.t, inp.y, inp.dy, inp.templates, inp.params)
    s.set_periods(periods)
    s.search_async()
    got = s.results()
    lay = s.layout
    s.close()
    assert lay["path"] == "resident" and lay["threads"] == 512 and lay["ctas_per_sm"] == 1, lay
    want = oracle.search_periods_c(inp.t, inp.y, inp.dy, periods, inp.templates, inp.params)
    assert_search_parity(got[:3], dict(y=inp.y, chi2=want[0], row=want[1], depth=want[2]), rtol=RTOL, label="N=7000 resident 512x1")
