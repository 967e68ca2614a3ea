"""bench.py on the CPU: the reference arm's JSON line (the driver parses it), the silent exit of the other ranks,
the roofline numerator (algorithmic bytes per period), and the refusal of the GPU arm to run
without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def _bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py")] + list(args), capture_output=True, text=True,
                          env=e, cwd=REPO, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--max-periods", "300")
    assert out.returncode == 0, out.stderr[-800:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["unit"] == "periods/s" and d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("cfg1") and d["config"]["n_points"] == 4320
    cb = d["cpu_baseline"]
    # the reference's own numba path (vendored under oracle/_ref, or /root/reference in the build container) is the arm;
    # the C restatement rides along as cpu_baseline.port and takes over only where the reference package is absent
    from oracle import ref_shim

    assert cb["kind"] == ("reference" if ref_shim.available() else "port")
    assert cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "periods" in cb["sample"]
    if cb["kind"] == "reference":
        assert "search_period" in cb["sample"] and "imap_unordered" in cb["sample"]
        assert cb["serial"]["cores"] == 1 and 0 < cb["serial"]["value"] < cb["value"] * 1.5
        assert cb["port"]["kind"] == "port" and cb["port"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_silently():
    out = _bench("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                 env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_needs_a_device():
    import torch

    if torch.cuda.is_available():
        return  # the GPU box runs the real thing
    out = _bench("--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-secondary")
    assert out.returncode != 0 and out.stdout.strip() == ""
    assert "no CPU fallback" in out.stderr


def test_algorithmic_bytes_of_cfg1():
    """B(P) = 24 N + 8 (N+M) + sum_{W admissible} [8 (N+M) + 4 L_W] + 24: 1.10 MB per period at cfg-1 with 24.3
    admissible widths on average, checked against a direct evaluation for a few periods."""
    import bench

    inp = bench.build_inputs("cfg1", 3)
    assert len(inp.periods) == 9679 and len(inp.y) == 4320
    total, mean_widths, M = bench.algorithmic_bytes(inp, inp.periods)
    assert M == 518
    assert abs(mean_widths - 24.25) < 0.05
    assert abs(total / len(inp.periods) - 1.0978e6) < 1e3
    # direct evaluation for single periods, widths filtered like core.py:143-156
    from tls_b200.grid import T14

    widths = np.asarray(inp.templates["width"])
    lengths = np.asarray(inp.templates["length"])
    N, span, prm = 4320, float(np.max(inp.t) - np.min(inp.t)), inp.params
    for p in (inp.periods[0], inp.periods[4000], inp.periods[-1]):
        dmax = T14(prm["R_star_max"], prm["M_star_max"], p, small=False)
        dmin = T14(prm["R_star_min"], prm["M_star_min"], p, small=True)
        corr = (span / p + 1) / (span / p)
        lo, hi = np.floor(dmin * N), np.ceil(dmax * N * corr)
        want = 24.0 * N + 8.0 * (N + M) + 24.0
        for w in np.unique(widths):
            if lo <= w <= hi:
                want += 8.0 * (N + M) + 4.0 * lengths[int(np.argmax(widths == w))]
        got, _, _ = bench.algorithmic_bytes(inp, [p])
        assert got == want


def test_tap_work_counts_the_gate_survivors():
    """bench.tap_work: T(P) = sum_W C_W * L_W on the actual input (6.3e5 taps per period at
    cfg-1 / 50 ppm, ten times that at 500 ppm), checked against a direct loop over offsets for one period."""
    import bench

    inp = bench.build_inputs("cfg1", 3)
    taps, rate, n = bench.tap_work(inp, inp.periods)
    assert n == 24 and 5.5e5 < taps < 7.0e5 and 0.08 < rate < 0.14
    noisy = bench.build_inputs("cfg1_500ppm", 3)
    assert 8 < bench.tap_work(noisy, noisy.periods)[0] / taps < 12
    # one period, by the definition (core.py:50-58), on a short curve
    small = bench.build_inputs("cfg1", 3)
    p = float(small.periods[len(small.periods) // 2])
    got = bench.tap_work(small, [p])[0]
    uniq, L, N, span, prm, T14 = bench.admissible_ranges(small)
    M = int(uniq.max()) + int(uniq.max()) % 2
    x = small.t * (1.0 / p)
    flux = small.y[np.argsort(x - np.floor(x), kind="mergesort")]
    patched = np.concatenate([flux, flux[:M]])
    lo = np.floor(T14(prm["R_star_min"], prm["M_star_min"], p, small=True) * N)
    hi = np.ceil(T14(prm["R_star_max"], prm["M_star_max"], p, small=False) * N * ((span / p + 1) / (span / p)))
    want = 0
    for W, Lw in zip(uniq, L):
        if lo <= W <= hi:
            xth = max(1, int(W / (1 / prm["T0_fit_margin"]))) if W > prm["T0_fit_margin"] > 0 else 1
            for i in range(0, len(patched) - W + 1):
                if i % xth == 0 and 1 - np.mean(patched[i:i + W]) > prm["transit_depth_min"]:
                    want += Lw
    assert abs(got - want) <= 1e-3 * want  # window means by cumulative sums vs direct means: a handful of ties at most
