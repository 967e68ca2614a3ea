#!/usr/bin/env python
"""Per-source-line summary of an ncu report (source page): samples and instructions executed.
usage: scripts/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-": data.append(r)
ix = {n: i for i, n in enumerate(hdr)}
tot_s = sum(int(r[ix["# Samples"]] or 0) for r in data); tot_i = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print("total samples %d, warp instructions %d" % (tot_s, tot_i))
data.sort(key=lambda r: -int(r[ix["# Samples"]] or 0))
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
for r in data[:top]:
    s = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[n]] or 0), n[6:]) for n in stalls), reverse=True)[:3]
    print("%5s %5.1f%% inst %5.1f%%  %-28s | %s" % (r[0], 100.0 * s / tot_s, 100.0 * int(r[ix["Instructions Executed"]] or 0) / tot_i,
          " ".join("%s:%d" % (n, v) for v, n in st if v), r[1].strip()[:90]))
