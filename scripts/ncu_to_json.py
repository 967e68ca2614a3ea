#!/usr/bin/env python
"""Key figures of one `ncu --set full` capture of the search kernel -> profiles/ncu_<workload>.json, the file bench.py
reads for roofline.traffic / roofline.bound / roofline.smem / roofline.issue_pct.  The JSON carries a hash of the
kernel sources it was captured from, so a stale profile is detected (tests/test_bench_contract.py).
usage: scripts/ncu_to_json.py gpurun_out/<tag>/prof_<workload>.ncu-rep <workload>"""
import csv, io, json, os, subprocess, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tls_b200 import build  # noqa: E402

rep, wl = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(key, default=None):
    if key not in m:
        return default
    v, u = m[key]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "nsecond": 1e-6}.get(u, 1.0)
    try:
        return float(v.replace(",", "")) * scale
    except ValueError:
        return default


wave = num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 0.0)
conf = num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 0.0)
pipes = {
    "smem_wavefronts_pct": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    "l1tex_pct": num("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    "issue_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fma_pipe_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    "fp64_pipe_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "lsu_pipe_pct": num("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    "dram_pct": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "l2_pct": num("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
}
ranked = sorted(((v, k) for k, v in pipes.items() if v is not None), reverse=True)
names = {"smem_wavefronts_pct": "smem", "l1tex_pct": "l1tex (shared-memory data stage)", "issue_pct": "issue", "fma_pipe_pct": "fp32 fma pipe",
         "fp64_pipe_pct": "fp64 pipe", "lsu_pipe_pct": "lsu", "dram_pct": "hbm", "l2_pct": "l2"}
out = {
    "workload": wl, "kernel": m.get("Kernel Name", ("?", ""))[0], "source_hash": build.source_hash(),
    "gpu_time_ms": num("gpu__time_duration.sum"),
    "dram_bytes_per_launch": (num("dram__bytes_read.sum", 0.0) or 0.0) + (num("dram__bytes_write.sum", 0.0) or 0.0),
    "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
    "registers_per_thread": num("launch__registers_per_thread"), "smem_per_block": num("launch__shared_mem_per_block_dynamic"),
    "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "inst_executed": num("smsp__inst_executed.sum"),
    "smem": {"wavefronts_per_launch": wave, "conflict_share": (conf / wave) if wave else None, "pct_of_peak": pipes["smem_wavefronts_pct"]},
    "pipes_pct_of_peak": pipes,
    "bound": names[ranked[0][1]] if ranked else None,
    "bound_ranking": ["%s %.1f%%" % (names[k], v) for v, k in ranked[:4]],
    "how": "ncu --set full --clock-control none --import-source on -k regex:tlsb_search -s 3 -c 1 python bench.py --workload %s --steps 3 --warmup 3 --no-cpu-baseline --no-secondary" % wl,
}
dst = os.path.join(REPO, "profiles", "ncu_%s.json" % wl)
with open(dst, "w") as f:
    json.dump(out, f, indent=1)
print(dst, out["bound_ranking"], "dram %.3g B" % out["dram_bytes_per_launch"])
