"""Batches and the iterative multi-planet search on the GPU (tls_b200/batch.py over
``tlsb_search_batch`` / ``tlsb_select_lightcurve`` / ``tlsb_final_t0_fit``).

A batch must give, curve by curve, exactly what ``.power()`` gives for that curve alone (same
kernels, same inputs), and ``.power()`` itself is held to the reference by test_gpu_power.py."""
import os
import warnings

import numpy as np
import pytest

from conftest import GOLDEN, assert_search_parity, load_search_golden

pytestmark = pytest.mark.gpu


def _curves(n_curves, hetero=False):
    from tls_b200 import workloads

    t, _, _, kw = workloads.lightcurve("small")
    rng = np.random.RandomState(42)
    ys, dys = [], []
    for c in range(n_curves):
        period = rng.uniform(2.0, 9.0)
        sigma = 10 ** rng.uniform(np.log10(50e-6), np.log10(500e-6))
        flux = workloads.inject(t, period, t[0] + rng.uniform(0, period), rp=0.03)
        ys.append(flux + rng.normal(0, sigma, len(t)))
        dys.append(sigma * rng.uniform(0.5, 2.0, len(t)) if hetero else np.full(len(t), np.std(ys[-1])))
    return t, np.array(ys), np.array(dys), kw


@pytest.mark.parametrize("hetero", [False, True])
def test_batch_equals_curve_by_curve_power(hetero):
    from tls_b200 import batch_power, transitleastsquares

    t, ys, dys, kw = _curves(5, hetero)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = batch_power(t, ys, dys if hetero else None, return_power=True, **kw)
        assert res.n_curves == 5
        for c in range(5):
            one = transitleastsquares(t, ys[c], dys[c] if hetero else None, verbose=False).power(
                show_progress_bar=False, verbose=False, **kw)
            np.testing.assert_array_equal(res.periods, one.periods)
            np.testing.assert_allclose(res.power[c], one.power, rtol=1e-12, atol=1e-12)
            for key in ("SDE", "SDE_raw", "period", "T0", "depth", "duration", "chi2_min", "chi2red_min", "rp_rs"):
                np.testing.assert_allclose(res[key][c], one[key], rtol=1e-12, err_msg="%s of curve %d" % (key, c))
            assert int(res.transit_count[c]) == one.transit_count


def test_batch_with_per_curve_time_axes_and_records():
    """tlsb_search_batch with one time axis per curve (same span): records equal the one-shot call's."""
    from tls_b200 import native, stats

    g = load_search_golden("small")
    rng = np.random.RandomState(3)
    n = len(g["t"])
    ts, ys, dys = [], [], []
    for c in range(3):
        jitter = np.zeros(n)
        jitter[1:-1] = rng.uniform(-1e-3, 1e-3, n - 2)  # inner samples move, the span does not
        ts.append(g["t"] + jitter)
        ys.append(g["y"] + rng.normal(0, 2e-5, n))
        dys.append(g["dy"])
    s = native.Searcher()
    s.set_templates(g["templates"], g["params"])
    s.set_periods(g["periods"])
    s.set_lightcurves(np.array(ts), np.array(ys), np.array(dys))
    assert s.n_curves == 3
    out = s.search_batch(stats.median_window(3))
    for c in range(3):
        chi2, row, depth, t0 = native.search_periods(ts[c], ys[c], dys[c], g["periods"], g["templates"], g["params"],
                                                     return_t0_index=True)
        np.testing.assert_array_equal(out["chi2"][c], chi2)
        np.testing.assert_array_equal(out["row"][c], row)
        np.testing.assert_array_equal(out["depth"][c], depth)
        np.testing.assert_array_equal(out["t0_index"][c], t0)
        order = np.argsort(g["periods"])
        SR, pr, pw, sde_raw, sde, amax = native.spectra(chi2[order], stats.median_window(3))
        np.testing.assert_array_equal(out["power"][c], pw)
        np.testing.assert_array_equal([out["SDE"][c], out["SDE_raw"][c]], [sde, sde_raw])
        assert np.isfinite(sde) and chi2.min() < len(g["y"])
        assert out["best_index"][c] == order[amax]
    # the device plan is reused inside a batch but a fresh search afterwards plans again
    s.select(1)
    s.search_async()
    chi2b = s.results()[0]
    np.testing.assert_array_equal(chi2b, out["chi2"][1])
    s.set_plan_mode(3)  # every period flagged, every 7th device range wrong: the batch repairs those periods
    out2 = s.search_batch(stats.median_window(3), want_power=False)
    np.testing.assert_array_equal(out2["chi2"], out["chi2"])
    np.testing.assert_array_equal(out2["row"], out["row"])
    assert s.plan_repairs >= len(g["periods"]) // 8
    s.close()


def test_batch_first_curve_of_golden_matches_reference():
    from tls_b200 import native, stats

    g = load_search_golden("cfg1_hetero")
    s = native.Searcher()
    s.set_templates(g["templates"], g["params"])
    s.set_periods(g["periods"])
    s.set_lightcurves(g["t"], np.array([g["y"], g["y"][::-1].copy()]), np.array([g["dy"], g["dy"]]))
    out = s.search_batch(stats.median_window(3), want_power=False)
    assert_search_parity((out["chi2"][0], out["row"][0], out["depth"][0]), g, rtol=1e-5, label="batch curve 0")
    s.close()


def test_multi_planet_known_answers_of_the_reference():
    """transitleastsquares/tests/test_multi_planet.py:33-49: mask the first planet of EPIC 201367065,
    search again; the reference's own numbers to 3 decimals."""
    from tls_b200 import search_planets

    z = np.load(os.path.join(GOLDEN, "power_k2_epic201367065.npz"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        found = search_planets(z["in_t"], z["in_y"], n_planets=2)
    assert len(found) == 2
    np.testing.assert_allclose(found[0].period, float(z["s_period"]), rtol=1e-5)
    np.testing.assert_allclose(found[0].SDE, float(z["s_SDE"]), rtol=1e-5)
    np.testing.assert_almost_equal(found[1].duration, 0.15061016994013998, decimal=3)
    np.testing.assert_almost_equal(found[1].SDE, 34.9911304598618, decimal=3)
    np.testing.assert_almost_equal(found[1].rp_rs, 0.025852178872027086, decimal=3)


def test_handle_pool_recycles_without_carrying_state():
    """batch_power borrows its handle from native.Searcher's per-device pool; a recycled handle (other
    curves, forced layout and plan mode left behind by the previous user) gives the same results."""
    from tls_b200 import batch_power, native

    native.Searcher.drain_pool()
    t, ys, dys, kw = _curves(4)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        first = batch_power(t, ys, **kw)
        idle = sum(len(v) for v in native.Searcher._POOL.values())
        assert idle == 1
        s = native.Searcher.acquire()  # the same handle, dirtied on purpose
        s.set_path("streaming")
        s.set_plan_mode(1)
        s.release()
        again = batch_power(t, ys[::-1].copy(), **kw)
        assert sum(len(v) for v in native.Searcher._POOL.values()) == 1
    for key in ("SDE", "period", "T0", "depth", "chi2_min"):
        np.testing.assert_array_equal(first[key], again[key][::-1], err_msg=key)
    native.Searcher.drain_pool()
    assert not native.Searcher._POOL


def _dist_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    from tls_b200 import batch_power, search_planets, workloads

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    warnings.simplefilter("ignore")
    t, y, dy, kw = workloads.lightcurve("small", planets=[4.3, 7.9])
    found = search_planets(t, y, n_planets=2, dist=dist, device=rank, **kw)
    tb, ys, dys, kwb = _curves(5)
    res = batch_power(tb, ys, dist=dist, device=rank, **kwb)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), periods=[r.period for r in found], SDE=[r.SDE for r in found],
             T0=[r.T0 for r in found], chi2=found[0].chi2, b_SDE=res.SDE, b_period=res.period, b_T0=res.T0)
    dist.destroy_process_group()


def test_power_and_batch_on_two_gpus_equal_one_gpu(tmp_path):
    """power(dist=...) / search_planets(dist=...) (periods sharded, one NCCL all-gather) and
    batch_power(dist=...) (curves sharded) on two GPUs, one process each: every rank gets exactly the
    one-GPU results."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from tls_b200 import batch_power, search_planets, workloads

    mp.spawn(_dist_worker, args=(2, 29571, str(tmp_path)), nprocs=2, join=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t, y, dy, kw = workloads.lightcurve("small", planets=[4.3, 7.9])
        one = search_planets(t, y, n_planets=2, **kw)
        tb, ys, dys, kwb = _curves(5)
        one_b = batch_power(tb, ys, **kwb)
    for rank in range(2):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        np.testing.assert_array_equal(z["periods"], [r.period for r in one])
        np.testing.assert_array_equal(z["T0"], [r.T0 for r in one])
        np.testing.assert_allclose(z["SDE"], [r.SDE for r in one], rtol=1e-12)
        np.testing.assert_array_equal(z["chi2"], one[0].chi2)
        np.testing.assert_array_equal(z["b_period"], one_b.period)
        np.testing.assert_array_equal(z["b_T0"], one_b.T0)
        np.testing.assert_allclose(z["b_SDE"], one_b.SDE, rtol=1e-12)


def _oracle_check(inp, got, n_check, label):
    """(chi2, row, depth) of the CUDA path against the C oracle on ``n_check`` periods spread over the grid."""
    from oracle import oracle

    chi2, row, depth = got[:3]
    sel = np.unique(np.linspace(0, len(inp.periods) - 1, n_check).astype(int))
    w = oracle.search_periods_c(inp.t, inp.y, inp.dy, inp.periods[sel], inp.templates, inp.params)
    np.testing.assert_array_equal(np.asarray(row)[sel], w[1], err_msg=label + ": rows")
    fin = np.isfinite(w[0])
    np.testing.assert_allclose(np.asarray(chi2)[sel][fin], w[0][fin], rtol=1e-5, atol=0, err_msg=label + ": chi2")
    np.testing.assert_allclose(np.asarray(depth)[sel], w[2], rtol=1e-5, atol=0, err_msg=label + ": depth")


def test_cfg4_batch_records_equal_the_oracle():
    """cfg-4 (BASELINE.json: a batch of K2-like 90 d curves): the records ``tlsb_search_batch`` returns for several
    curves of the batch — own planet, own noise level — against the C ORACLE (core.py:96-188 restated), not
    against this repository's own ``.power()``."""
    from tls_b200 import native, stats, transitleastsquares, workloads

    t, ys = workloads.batch_lightcurves(6)
    dys = np.repeat(np.std(ys, axis=1)[:, None], ys.shape[1], axis=1)
    model = transitleastsquares(t, ys[0], verbose=False)  # dy=None: std(y) everywhere (validate.py:39-40), as in the batch
    inputs = model.prepare(verbose=False, show_progress_bar=False)
    s = native.Searcher()
    try:
        s.set_templates(inputs.templates, inputs.params)
        s.set_periods(inputs.periods)
        s.set_lightcurves(t, ys, dys)
        out = s.search_batch(stats.median_window(model.oversampling_factor), want_power=False)
    finally:
        s.close()
    for c in (0, 2, 3, 5):
        one = transitleastsquares(t, ys[c], verbose=False).prepare(verbose=False, show_progress_bar=False)
        np.testing.assert_array_equal(one.dy, dys[c])
        np.testing.assert_array_equal(one.periods, inputs.periods)
        _oracle_check(one, (out["chi2"][c], out["row"][c], out["depth"][c]), 40, "cfg-4 curve %d" % c)


def test_cfg5_mask_and_rerun_equals_the_oracle_after_every_mask():
    """cfg-5 (BASELINE.json: mask + rerun x3; tests/test_multi_planet.py:33-40): a three-planet curve of ~13 k points
    (the tiled kernel).  After EVERY mask the search's (chi2, row, depth) must equal the C oracle on >= 100 periods of
    that run's own grid, and the three planets must come out."""
    from tls_b200 import cleaned_array, native, transit_mask, transitleastsquares, workloads

    rng = np.random.RandomState(5)
    t = np.linspace(0.0, 270.0, 12960)  # 270 d @ 30 min
    planets = (5.3, 11.7, 23.9)
    flux = np.ones(len(t))
    for per in planets:
        flux = flux * workloads.inject(t, per, 1.0 + 0.37 * per, rp=0.03)
    y = flux + rng.normal(0, 150e-6, len(t))
    found = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for run in range(3):
            model = transitleastsquares(t, y, verbose=False)
            inp = model.prepare(verbose=False, show_progress_bar=False, period_max=60.0)
            got = native.search_periods(inp.t, inp.y, inp.dy, inp.periods, inp.templates, inp.params)
            _oracle_check(inp, got, 110, "cfg-5 run %d (N = %d)" % (run, len(inp.t)))
            res = model.power(show_progress_bar=False, verbose=False, period_max=60.0)
            found.append(float(res.period))
            keep = ~transit_mask(t, res.period, 2 * res.duration, res.T0)
            t, y = cleaned_array(t[keep], y[keep])
    for per in planets:
        assert min(abs(f - per) / per for f in found) < 2e-3, (per, found)
