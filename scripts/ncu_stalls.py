#!/usr/bin/env python
"""Warp-stall sample totals by reason of an ncu report (source page). usage: scripts/ncu_stalls.py report.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-": data.append(r)
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = {n: sum(int(r[ix[n]] or 0) for r in data) for n in stalls}
all_ = sum(tot.values())
print(" ".join("%s %.1f%%" % (n[6:], 100.0 * v / all_) for n, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 200 > all_))
