#!/usr/bin/env python
"""Shared-memory wavefronts (and the excess = bank-conflict part) per CUDA source line of an .ncu-rep.
usage: ncu_smem.py report.ncu-rep [n_images] [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nimg = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
lines, cur = [], ""
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Line No":
        iW, iX, iI = r.index("L1 Wavefronts Shared"), r.index("L1 Wavefronts Shared Excessive"), r.index("Instructions Executed")
    elif r and r[0] == "File Path":
        cur = r[1]
    elif r and r[0].isdigit() and len(r) > 8:
        try:
            lines.append((cur.split("/")[-1], int(r[0]), r[1], int(r[iW]), int(r[iX]), int(r[iI])))
        except ValueError:
            pass
tw, tx = sum(l[3] for l in lines), sum(l[4] for l in lines)
print("shared-memory wavefronts per image %.0f, of which excess (bank conflicts) %.0f" % (tw / nimg, tx / nimg))
for f, ln, s, w, x, ins in sorted(lines, key=lambda l: -l[3])[:top]:
    print("%-12s %4d  wf %7.0f (%4.1f%%)  excess %6.0f  instr %6.0f  %s" % (f[:12], ln, w / nimg, 100.0 * w / tw, x / nimg, ins / nimg, s.strip()[:90]))
