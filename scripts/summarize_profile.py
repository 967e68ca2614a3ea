#!/usr/bin/env python
"""Turn one gpurun_out/<tag>/ capture into the tracked files under profiles/.

usage: scripts/summarize_profile.py <tag> <round> [workload]
writes profiles/r<round>_<workload>_launches.csv   (ncu --metrics gpu__time_duration.sum launch list)
       profiles/r<round>_<workload>_ncu_summary.md (key metrics + hottest source lines of the ncu --set full capture)
       profiles/traffic_<workload>.json            (dram bytes per launch of the search kernel, read by bench.py)
"""
import csv, io, json, os, subprocess, sys

tag, rnd = sys.argv[1], sys.argv[2]
wl = sys.argv[3] if len(sys.argv) > 3 else "cfg1"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(REPO, "gpurun_out", tag)
dst = os.path.join(REPO, "profiles")
os.makedirs(dst, exist_ok=True)

# ---- launch list
rows = []
with open(os.path.join(src, "launches_%s.csv" % wl)) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(io.StringIO("".join(lines))):
    rows.append((r["ID"], r["Kernel Name"], r["Block Size"], r["Grid Size"], int(r["Metric Value"])))
with open(os.path.join(dst, "r%s_%s_launches.csv" % (rnd, wl)), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --workload %s --steps 3 --warmup 3 --no-cpu-baseline\n" % wl)
    f.write("# cold-cache, serialised launches: compare SHARES, not absolutes\n")
    f.write("id,kernel,block,grid,gpu_time_ns\n")
    for r in rows:
        f.write('%s,"%s","%s","%s",%d\n' % r)
share = {}
for _, k, _, _, ns in rows:
    import re
    mm = re.search(r"(tlsb_\w+?_kernel)", k)
    name = mm.group(1) if mm else "torch fill (L2 flush / buffers, outside the timed events)"
    share[name] = share.get(name, 0) + ns
ours = {k: v for k, v in share.items() if k.startswith("tlsb_")}
tot = float(sum(ours.values()))

# ---- full capture
rep = os.path.join(src, "prof_%s.ncu-rep" % wl)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rr[0], rr[1], rr[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]

def unit_bytes(v, u):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return float(v) * f

traffic = unit_bytes(*m["dram__bytes_read.sum"]) + unit_bytes(*m["dram__bytes_write.sum"])
with open(os.path.join(dst, "traffic_%s.json" % wl), "w") as f:
    json.dump({"kernel": m["Kernel Name"][0] if "Kernel Name" in m else "tlsb_search_kernel",
               "dram_bytes_per_launch": traffic, "source": "ncu --set full, profiles/r%s_%s_ncu_summary.md" % (rnd, wl)}, f)

lines_out = subprocess.run([sys.executable, os.path.join(REPO, "scripts", "ncu_breakdown.py"), rep], capture_output=True, text=True).stdout
bench = {}
try:
    bench = json.load(open(os.path.join(src, "bench_%s.json" % wl)))
except Exception:
    pass
with open(os.path.join(dst, "r%s_%s_ncu_summary.md" % (rnd, wl)), "w") as f:
    f.write("# Round %s — ncu summary, workload %s (capture %s)\n\n" % (rnd, wl, tag))
    f.write("Commands (scripts/gpu_check.sh, run under gpurun on one B200):\n\n```\n")
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv ... python bench.py --workload %s --steps 3 --warmup 3 --no-cpu-baseline --no-secondary\n" % wl)
    f.write("ncu --set full --clock-control none --import-source on -k regex:tlsb_search -s 3 -c 1 ... (same command)\n```\n\n")
    if bench:
        r = bench["roofline"]
        f.write("Bench line of the same build (not under a profiler): value %.0f periods/s, %.3f ms/step, e2e %.0f periods/s, "
                "search kernel %.3f ms/launch (CUDA events), roofline.frac %.3f of measured HBM peak, kernel share of step %.3f.\n\n"
                % (bench["value"], bench["ms_per_step"], bench["e2e"]["value"], r["kernel_ms_per_launch"], r["frac"], r["kernel_share_of_step"]))
    f.write("## Launch list: share of our kernels' device time (cold-cache, serialised)\n\n| kernel | ns total | share |\n|---|---|---|\n")
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1]):
        f.write("| %s | %d | %.1f %% |\n" % (k, v, 100 * v / tot))
    for k, v in share.items():
        if not k.startswith("tlsb_"):
            f.write("| %s | %d | not ours |\n" % (k, v))
    f.write("\n## Search kernel, ncu --set full (one launch)\n\n| metric | value | unit |\n|---|---|---|\n")
    f.write("| kernel | %s | |\n" % (m.get("Kernel Name", ("?", ""))[0]))
    for k in KEYS:
        if k in m:
            f.write("| %s | %s | %s |\n" % (k, m[k][0], m[k][1]))
    f.write("| dram traffic per launch (read+write) | %.0f | byte |\n" % traffic)
    f.write("\n## Hottest source lines (warp-state samples; line numbers of tls_b200/csrc/tlsb_device.cuh, tlsb_resident.cu / tlsb_tiled.cu)\n\n```\n%s```\n" % lines_out)
print("ok", share)
