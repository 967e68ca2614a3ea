#!/usr/bin/env python
"""Per-source-line shared-memory wavefronts of an ncu report (actual vs ideal). usage: scripts/ncu_smem.py report.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-": data.append(r)
ix = {n: i for i, n in enumerate(hdr)}
I = lambda x: int(x or 0) if (x or "0").isdigit() else 0
W, WI, IE = ix["L1 Wavefronts Shared"], ix["L1 Wavefronts Shared Ideal"], ix["Instructions Executed"]
tot = sum(I(r[W]) for r in data); toti = sum(I(r[WI]) for r in data)
print("shared wavefronts %.3e, ideal %.3e (excess %.0f%%)" % (tot, toti, 100.0 * (tot - toti) / max(1, tot)))
data.sort(key=lambda r: -I(r[W]))
for r in data[:top]:
    print("%5s wf %5.1f%% (%.2e) ideal %.2e  inst %.2e | %s" % (r[0], 100.0 * I(r[W]) / tot, I(r[W]), I(r[WI]), I(r[IE]), r[1].strip()[:100]))
