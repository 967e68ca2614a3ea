"""The fp32 filter pass (equal weights): an fp32 correlation with a rigorous error bound decides which gate
survivors can still be the minimum; only those are evaluated in fp64.  The claim is that nothing changes:
results with the filter on are BIT-IDENTICAL to results with the filter off (every survivor through the
exact fp64 evaluation), on every layout, and both match the reference goldens (core.py:57-74)."""
import numpy as np
import pytest

from conftest import assert_search_parity, load_search_golden, search_goldens

pytestmark = pytest.mark.gpu


def _uniform(g):
    return np.all(g["dy"] == g["dy"][0])


def _run(g, on, path="auto", chunk=0, stats=False, periods=None):
    from tls_b200 import native

    s = native.Searcher()
    try:
        s.set_inputs(g["t"], g["y"], g["dy"], g["templates"], g["params"])
        s.set_periods(g["periods"] if periods is None else periods)
        s.set_path(path, chunk)
        s.set_filter(on, stats)
        s.search_async()
        out = s.results()
        return out, (s.filter_stats if stats else None), s.path
    finally:
        s.close()


@pytest.mark.parametrize("name", [n for n in search_goldens()])
def test_filter_on_equals_filter_off_bit_for_bit(name):
    g = load_search_golden(name)  # per-point weights: the filter runs two fp32 correlations (resident layouts)
    on, _, path = _run(g, True)
    off, _, _ = _run(g, False)
    for a, b, what in zip(on, off, ("chi2", "row", "depth", "t0_index")):
        np.testing.assert_array_equal(a, b, err_msg="%s differs with the filter on (%s, %s path)" % (what, name, path))
    assert_search_parity(on[:3], g, rtol=1e-5, label=name)


@pytest.mark.parametrize("name,path,chunk", [("cfg1_500ppm", "tiled", 1536), ("small", "tiled", 512),
                                              ("ragged_L", "tiled", 700), ("cfg3", "tiled", 0),
                                              ("ties_unsorted", "tiled", 600), ("cfg1_hetero", "tiled", 1536)])
def test_filter_on_equals_off_on_the_tiled_layout(name, path, chunk):
    g = load_search_golden(name)
    on, _, used = _run(g, True, path, chunk)
    off, _, _ = _run(g, False, path, chunk)
    assert used == "tiled"
    for a, b, what in zip(on, off, ("chi2", "row", "depth", "t0_index")):
        np.testing.assert_array_equal(a, b, err_msg="%s differs with the filter on (%s, tiled)" % (what, name))
    assert_search_parity(on[:3], g, rtol=1e-5, label=name)


def test_filter_passes_few_candidates_on_noisy_data():
    """cfg-1 at 500 ppm: about half of all offsets pass the gate; the filter must leave the fp64 evaluation
    a small fraction of them (the bound is rigorous, not tuned: a regression here means a slow kernel,
    never a wrong one)."""
    from tls_b200 import transitleastsquares, workloads

    t, y, dy, kw = workloads.lightcurve("cfg1_500ppm")
    inp = transitleastsquares(t, y, dy, verbose=False).prepare(verbose=False, **kw)
    g = dict(t=inp.t, y=inp.y, dy=inp.dy, templates=inp.templates, params=inp.params, periods=inp.periods[::16])
    on, st, path = _run(g, True, stats=True)
    off, _, _ = _run(g, False)
    assert path == "resident"
    for a, b in zip(on, off):
        np.testing.assert_array_equal(a, b)
    assert st["candidates"] > 1000 * len(g["periods"])
    assert st["finalists"] < 0.05 * st["candidates"], st
    assert st["overflows"] == 0, st


def _oracle(g, periods=None):
    from oracle import oracle

    return oracle.search_periods_c(g["t"], g["y"], g["dy"], g["periods"] if periods is None else periods, g["templates"], g["params"])


@pytest.mark.parametrize("name,path,chunk", [("small", "auto", 0), ("cfg1_500ppm", "auto", 0), ("small", "tiled", 512),
                                              ("cfg1_500ppm", "tiled", 1536), ("cfg1_hetero", "auto", 0), ("cfg1_hetero", "tiled", 1536)])
@pytest.mark.parametrize("case", ["scale_0.97", "scale_1.02", "offset_+3e-3", "offset_-3e-3", "trend", "scale_0.5"])
def test_fp32_gate_on_flux_that_is_not_normalised(name, path, chunk, case):
    """The fp32 gate tests detrended cumulative sums cs32[k] = fl32(cs[k] - k mu) against a threshold lowered by a rigorous
    error bound (tlsb_device.cuh: Gate32); candidates it lets through are settled by the exact fp64 gate in bound_one.
    Flux that is not normalised to 1 makes the cumulative sums grow linearly (what the detrending is for) and, with a
    trend, leaves large excursions after it (what the error bound is for).  Whatever the gate's resolution: filter on ==
    filter off bit for bit, and both equal the C oracle (core.py:58 gate, rows exact, chi2 to 1e-9)."""
    g = dict(load_search_golden(name))
    y = g["y"].copy()
    if case.startswith("scale_"):
        y = y * float(case.split("_")[1])
    elif case.startswith("offset_"):
        y = y + float(case.split("_")[1])
    else:
        y = y * (1.0 + 4e-3 * (g["t"] - g["t"].mean()) / (g["t"].max() - g["t"].min()))
    g["y"] = y
    periods = g["periods"][:: max(1, len(g["periods"]) // 24)]
    on, _, used = _run(g, True, path, chunk, periods=periods)
    off, _, _ = _run(g, False, path, chunk, periods=periods)
    for a, b, what in zip(on, off, ("chi2", "row", "depth", "t0_index")):
        np.testing.assert_array_equal(a, b, err_msg="%s differs with the filter on (%s, %s, %s)" % (what, name, case, used))
    want = _oracle(g, periods)
    np.testing.assert_array_equal(on[1], want[1], err_msg="rows (%s, %s, %s)" % (name, case, used))
    fin = np.isfinite(want[0])
    np.testing.assert_array_equal(on[0][~fin], want[0][~fin])
    np.testing.assert_allclose(on[0][fin], want[0][fin], rtol=1e-9, atol=0)
    np.testing.assert_allclose(on[2], want[2], rtol=1e-9, atol=1e-300)


def test_memo_switch_changes_nothing_but_the_launch_count():
    """TLSB_MEMO=0 makes every call redo the plan and the derived template arrays (bench.py's e2e figure); results are
    identical either way."""
    import os

    from tls_b200 import native

    g = load_search_golden("small")
    outs = {}
    for memo in ("1", "0"):
        os.environ["TLSB_MEMO"] = memo
        try:
            s = native.Searcher()
            s.set_inputs(g["t"], g["y"], g["dy"], g["templates"], g["params"])
            s.set_periods(g["periods"])
            counts = []
            for _ in range(3):
                s.set_inputs(g["t"], g["y"], g["dy"], g["templates"], g["params"])  # the same bank again
                s.search_async()
                counts.append(s.launch_count)
                outs.setdefault(memo, []).append(s.results())
            s.close()
        finally:
            os.environ.pop("TLSB_MEMO", None)
        assert counts == ([2, 1, 1] if memo == "1" else [2, 2, 2]), (memo, counts)
    for a, b in zip(outs["1"], outs["0"]):
        for x, z in zip(a, b):
            np.testing.assert_array_equal(x, z)


@pytest.mark.parametrize("name", [n for n in search_goldens()])
def test_unequal_weights_filter_equals_the_all_fp64_kernel(name, monkeypatch):
    """Per-point dy: the filter layouts (fp32 gate, two fp32 correlations, exact fp64 for finalists) against the
    all-fp64 kernels (TLSB_WFILTER=0: tap_block / block_min on every gate survivor).  The two sum the taps in different
    orders, so chi2 agrees to rounding (1e-11), rows and t0 exactly."""
    g = load_search_golden(name)
    if _uniform(g):
        pytest.skip("equal weights")
    new, _, path = _run(g, True)
    monkeypatch.setenv("TLSB_WFILTER", "0")
    old, _, _ = _run(g, True)
    np.testing.assert_array_equal(new[1], old[1], err_msg="rows (%s, %s)" % (name, path))
    np.testing.assert_array_equal(new[3], old[3], err_msg="t0 index (%s, %s)" % (name, path))
    fin = np.isfinite(old[0])
    np.testing.assert_array_equal(new[0][~fin], old[0][~fin])
    np.testing.assert_allclose(new[0][fin], old[0][fin], rtol=1e-11, atol=0)
    np.testing.assert_allclose(new[2], old[2], rtol=1e-11, atol=1e-300)
