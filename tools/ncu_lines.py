#!/usr/bin/env python
"""Attribute an .ncu-rep's executed warp instructions and stall samples to CUDA source lines and
device functions of one file.  usage: ncu_lines.py report.ncu-rep source.cu [n_images] [top]"""
import csv
import io
import re
import subprocess
import sys

rep, srcfile = sys.argv[1], sys.argv[2]
nimg = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines = []
for i, r in enumerate(rows):
    if r and r[0] == "Line No":
        hdr = r
        iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        fname = rows[i - 2][1] if i >= 2 else ""
        cur = fname
    elif r and r[0] == "File Path":
        cur = r[1]
    elif r and r[0].isdigit() and len(r) > 8:
        try:
            lines.append((cur, int(r[0]), r[1], int(r[iI]), int(r[iS])))
        except ValueError:
            pass
tot = sum(l[3] for l in lines); ts = max(1, sum(l[4] for l in lines))
print("total warp instr %d (%.0f per image), samples %d" % (tot, tot / nimg, ts))
src = open(srcfile).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"\s*(?:template.*>\s*)?(?:__device__|__global__).*?(\w+)\(", l)
    if m:
        funcs.append((i, m.group(1)))


def fn(path, line):
    if not path.endswith(srcfile.split("/")[-1]):
        return path.split("/")[-1]
    name = "?"
    for i, n in funcs:
        if i <= line:
            name = n
    return name


agg = {}
for path, ln, s, ins, sm in lines:
    a = agg.setdefault(fn(path, ln), [0, 0]); a[0] += ins; a[1] += sm
for k, (i, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-30s %9.0f instr/img %5.1f%%  samples %5.1f%%" % (k, i / nimg, 100 * i / tot, 100 * sm / ts))
print()
for path, ln, s, ins, sm in sorted(lines, key=lambda l: -l[3])[:top]:
    print("%-14s %4d %8.0f %5.1f%% smp %5.1f%%  %s" % (path.split("/")[-1][:14], ln, ins / nimg, 100 * ins / tot, 100 * sm / ts, s.strip()[:100]))
