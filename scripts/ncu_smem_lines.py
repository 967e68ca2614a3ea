#!/usr/bin/env python
"""Shared-memory wavefronts per source line of an ncu report: total, ideal, excessive (bank conflicts).
usage: scripts/ncu_smem_lines.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hdr = None; rows = []
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ix = {n: i for i, n in enumerate(hdr)}; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":
        w = int(r[ix["L1 Wavefronts Shared"]] or 0); e = int(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
        if w: rows.append((w, e, cur, int(r[0]), r[1].strip()))
tot = sum(r[0] for r in rows); tote = sum(r[1] for r in rows)
print("shared wavefronts %d, excessive %d (%.1f%%)" % (tot, tote, 100.0 * tote / max(tot, 1)))
for w, e, f, ln, text in sorted(rows, reverse=True)[:top]:
    print("%-18s %5d  wavefronts %5.1f%%  excessive %5.1f%% of all (%.0f%% of this line) | %s" % (f, ln, 100.0 * w / tot, 100.0 * e / tot, 100.0 * e / w, text[:70]))
