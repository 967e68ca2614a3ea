"""BASELINE.json's full sizes through size-independent properties (the oracle cannot finish these
grids in seconds): the whole cfg-1 grid and the whole Kepler-long cfg-2 grid (186,681 periods of
70,128 samples), plus an oracle spot check spread over each grid.

Properties of the statistic: scaling every dy by c divides chi2 by c^2 and leaves
rows, depths and window starts untouched (powers of two: exact); the order of the trial periods
does not matter; a search restricted to a subset of the periods reproduces the full search."""
import numpy as np
import pytest

from conftest import assert_search_parity

pytestmark = pytest.mark.gpu


def _inputs(workload, hetero=False):
    from tls_b200 import transitleastsquares, workloads

    t, y, dy, kw = workloads.lightcurve(workload, hetero=hetero)
    return transitleastsquares(t, y, dy, verbose=False).prepare(verbose=False, **kw)


def _search(inp, periods, dy=None, path="auto"):
    from tls_b200 import native

    s = native.Searcher()
    s.set_inputs(inp.t, inp.y, inp.dy if dy is None else dy, inp.templates, inp.params)
    s.set_periods(periods)
    s.set_path(path)
    s.search_async()
    out = s.results()
    info = dict(s.layout, **s.sort_info)
    s.close()
    return out, info


@pytest.mark.parametrize("workload,want_path", [("cfg1", "resident"), ("cfg2", "tiled")])
def test_whole_grid_properties(workload, want_path):
    from oracle import oracle

    inp = _inputs(workload)
    P = len(inp.periods)
    full, info = _search(inp, inp.periods)
    assert info["path"] == want_path
    chi2, row, depth, t0 = full
    n = len(inp.y)
    assert np.all(chi2 <= n) and np.all(chi2 > 0)          # a fitted model beats the sentinel N (core.py:46)
    assert np.all((row >= 0) & (row < len(inp.templates["width"])))
    fitted = chi2 < n
    assert np.all((depth[fitted] > 0) & (depth[fitted] < 1)) and np.all(depth[~fitted] == 0)
    # the injected planet is the global minimum (10.123 d or a harmonic of it)
    best = inp.periods[np.argmin(chi2)]
    ratio = best / 10.123
    assert min(abs(ratio - k) for k in (0.5, 1.0, 2.0, 3.0)) < 0.01, best

    # dy -> 4 dy: wherever a model beat the sentinel N, chi2 / 16 exactly (power of two) with the same
    # row, depth and window start (the sentinel itself does not scale, so unfitted periods may now fit)
    assert fitted.sum() > 1000
    scaled, _ = _search(inp, inp.periods, dy=inp.dy * 4.0)
    np.testing.assert_array_equal(scaled[1][fitted], row[fitted])
    np.testing.assert_array_equal(scaled[3][fitted], t0[fitted])
    np.testing.assert_array_equal(scaled[2][fitted], depth[fitted])
    np.testing.assert_array_equal(scaled[0][fitted], chi2[fitted] / 16.0)
    assert np.all(scaled[0] <= n) and (scaled[0] < n).sum() > fitted.sum()

    # permuted and strided subsets reproduce the full search bit for bit
    rng = np.random.RandomState(1)
    sel = rng.permutation(P)[: max(500, P // 40)]
    sub, _ = _search(inp, inp.periods[sel])
    for a, b in zip(sub, full):
        np.testing.assert_array_equal(a, b[sel])

    # oracle spot check spread over the whole grid (kept to what it finishes in seconds)
    spot = np.linspace(0, P - 1, 96 if workload == "cfg2" else 400).astype(int)
    w = oracle.search_periods_c(inp.t, inp.y, inp.dy, inp.periods[spot], inp.templates, inp.params)
    ref = dict(y=inp.y, chi2=w[0], row=w[1], depth=w[2])
    assert_search_parity((chi2[spot], row[spot], depth[spot]), ref, rtol=1e-5, label=workload + " spot")


@pytest.mark.parametrize("workload,path", [("cfg3", "tiled"), ("cfg1", "tiled"), ("cfg2", "auto")])
def test_per_point_weights_on_the_long_curve_layouts(workload, path):
    """Unequal weights (two correlations; block size 7 in the filter layouts, 5 in the all-fp64 kernels) through the
    tiled kernel, and the widest windows of cfg-2 through whatever layout holds them, against the oracle."""
    from oracle import oracle

    inp = _inputs(workload, hetero=True)
    sel = np.linspace(0, len(inp.periods) - 1, 60 if workload != "cfg1" else 300).astype(int)
    periods = inp.periods[sel]
    got, info = _search(inp, periods, path=path)
    if path != "auto":
        assert info["path"] == path and info["block"] in (5, 7)
    w = oracle.search_periods_c(inp.t, inp.y, inp.dy, periods, inp.templates, inp.params)
    ref = dict(y=inp.y, chi2=w[0], row=w[1], depth=w[2])
    assert_search_parity(got[:3], ref, rtol=1e-5, label="%s hetero %s" % (workload, info["path"]))
