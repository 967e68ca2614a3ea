#!/usr/bin/env python
"""Warp-state samples of an ncu report grouped by source-line ranges (phases).
usage: scripts/ncu_phases.py report.ncu-rep name:lo-hi [name:lo-hi ...]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
ranges = []
for spec in sys.argv[2:]:
    name, r = spec.split(":"); lo, hi = r.split("-"); ranges.append((name, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-": data.append(r)
ix = {n: i for i, n in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
acc = {n: 0 for n, _, _ in ranges}; other = 0
for r in data:
    ln = int(r[0]); s = int(r[ix["# Samples"]] or 0)
    for n, lo, hi in ranges:
        if lo <= ln <= hi: acc[n] += s; break
    else: other += s
for n, lo, hi in ranges: print("%-28s lines %4d-%4d  %5.1f%%" % (n, lo, hi, 100.0 * acc[n] / tot))
print("%-28s %5.1f%%" % ("other", 100.0 * other / tot))
