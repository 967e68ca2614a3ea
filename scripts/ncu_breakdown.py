#!/usr/bin/env python
"""Warp-state samples of an ncu report grouped by (source file, line range), so that phases that live in different
files (tlsb_device.cuh / tlsb_tiled.cu / tlsb_resident.cu) do not mix.
usage: scripts/ncu_breakdown.py report.ncu-rep [name=file:lo-hi ...]      (file = a substring of the path)
Without specs: the top lines per file."""
import csv, io, subprocess, sys
from collections import defaultdict

rep = sys.argv[1]
specs = []
for spec in sys.argv[2:]:
    name, rest = spec.split("=")
    f, r = rest.split(":")
    lo, hi = r.split("-")
    specs.append((name, f, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
hdr = None
lines = []  # (file, line, text, samples, inst, stall dict)
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {n: i for i, n in enumerate(hdr)}
        stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        s = int(r[ix["# Samples"]] or 0)
        lines.append((cur, int(r[0]), r[1].strip(), s, int(r[ix["Instructions Executed"]] or 0),
                      {n[6:]: int(r[ix[n]] or 0) for n in stalls}))
tot = sum(l[3] for l in lines) or 1
toti = sum(l[4] for l in lines) or 1
if specs:
    acc = defaultdict(lambda: [0, 0, defaultdict(int)])
    for f, ln, _, s, i, st in lines:
        for name, sf, lo, hi in specs:
            if sf in f and lo <= ln <= hi:
                key = name
                break
        else:
            key = "other:" + f
        acc[key][0] += s
        acc[key][1] += i
        for k, v in st.items():
            acc[key][2][k] += v
    for key, (s, i, st) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
        print("%-34s samples %5.1f%%  inst %5.1f%%   %s" % (key, 100.0 * s / tot, 100.0 * i / toti,
                                                            " ".join("%s %.0f%%" % (k, 100.0 * v / max(s, 1)) for k, v in top)))
else:
    for f, ln, text, s, i, st in sorted(lines, key=lambda l: -l[3])[:50]:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print("%-18s %5d %5.1f%% inst %5.1f%%  %-34s | %s" % (f, ln, 100.0 * s / tot, 100.0 * i / toti,
                                                               " ".join("%s:%d" % kv for kv in top if kv[1]), text[:80]))
