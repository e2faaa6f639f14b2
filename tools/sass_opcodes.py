#!/usr/bin/env python
"""SASS evidence for the Blackwell-native claim and the no-spill claim, regenerated from the built library.

    python tools/sass_opcodes.py [lib.so] > profiles/r02_sass_opcodes.txt

Per kernel of libyoloface_b200.so (cuobjdump -sass): instruction count and the counts of the opcodes that prove what the
kernel runs on -- UTCIMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), UTMALDG (TMA tensor load), UBLKCP
(cp.async.bulk), SYNCS (mbarrier), IDP (dp4a), IMAD.HI / IMAD.WIDE (requant multiply), LDL / STL (local-memory traffic =
register spills or dynamically indexed local arrays) -- next to `cuobjdump -res-usage` (registers, stack, shared)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "stm32h7-yolo_b200", "libyoloface_b200.so")
WATCH = ["UTCIMMA", "LDTM", "UTCBAR", "UTMALDG", "UBLKCP", "SYNCS", "IDP", "IMAD.HI", "IMAD.WIDE", "VIMNMX3", "LDS", "STS", "LDL", "STL", "BAR"]

sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in line:
        usage[cur] = line.strip()
        cur = None
kern, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); counts[kern] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if kern and m:
        op = m.group(1)
        counts[kern]["_total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[kern][w] += 1
print("# %s" % os.path.relpath(lib, ROOT))
print("# cuobjdump -sass / -res-usage, CUDA %s" % subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1])
demangle = subprocess.run(["c++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
for (k, c), name in zip(counts.items(), demangle):
    short = re.sub(r"\(.*", "", name)
    print("\n%s\n  %s" % (short, usage.get(k, "")))
    print("  instructions %d | " % c["_total"] + "  ".join("%s %d" % (w, c[w]) for w in WATCH if c[w]))
    if c["LDL"] or c["STL"]:
        print("  -> local-memory instructions present (LDL %d, STL %d)" % (c["LDL"], c["STL"]))
