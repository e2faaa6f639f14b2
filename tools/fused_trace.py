#!/usr/bin/env python
"""Per-phase SM-cycle trace of the fused kernel (CTA 0, first image).
usage: make -C stm32h7-yolo_b200/csrc TRACE=1 && YF_B200_LIB=stm32h7-yolo_b200/libyoloface_b200_trace.so python tools/fused_trace.py [n_images]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pkg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
yf = pkg.load()
net = yf.Network(chunk_images=max(n, 256), mode="fused")
imgs = np.load(os.path.join(ROOT, "tests", "golden", "images_56.npy"))
x = torch.from_numpy(np.concatenate([imgs] * (n // 27 + 1))[:n].copy()).cuda()
y = torch.empty((n, 7, 7, 18), dtype=torch.int8, device="cuda")
for _ in range(3):
    net.enqueue(x, y, n)
net.sync()
net.fused_trace(True)
net.enqueue(x, y, n); net.sync()
st = net.fused_trace(False, read=True)
steps = net.steps()
F = yf.fused_program(56, 56)
nph, split = len(steps), F["split"]
# phase order of CTA 0: front(image 0), front(image 1), back(pair), ... (csrc/yf_fused.cu advance_phase); the kernel
# stamps clock64() at the start of its first 80 phases and once at the end
slots = 148 * 3
grid = min(n, slots)
if int(os.environ.get("YF_B200_LAT_GRID", "0")) > 0 and n <= 148:
    grid = min(grid, int(os.environ["YF_B200_LAT_GRID"]))
my_images = (n + grid - 1) // grid
seq, p, k = [], 0, 0
while k < my_images and len(seq) < 79:
    seq.append((p, k))
    if p == split - 1 and split < nph:
        if (k & 1) or k == my_images - 1:
            p = split
        else:
            p, k = 0, k + 1
    elif p == nph - 1:
        p, k = 0, k + 1
    else:
        p += 1
print("n=%d: CTA 0 runs %d image(s); spec kernel %s; phases shown: %d" % (n, my_images, F.get("spec"), len(seq)))
tot = st[len(seq)] - st[0] if len(seq) < 80 else st[79] - st[0]
per_kind = {}
for i, (p, k) in enumerate(seq[:79]):
    if i + 1 >= 80:
        break
    d = st[i + 1] - st[i]
    s = steps[p]
    tag = "back " if p >= split else "front"
    per_kind[tag] = per_kind.get(tag, 0) + d
    print("  img %d %s %-14s kind %d rows %4d cout %2d  %7d cyc  %5.1f%%" % (k, tag, s["name"], s["kind"], F["phases"][p]["rows_out"], F["phases"][p]["cout"], d, 100.0 * d / max(tot, 1)))
print("total %d cycles; %s" % (tot, per_kind))
# inner stamps of single phases (thread 0's view), one extra launch per phase: YF_B200_TRACE_INNER=3,5,8
inner = [int(v) for v in os.environ.get("YF_B200_TRACE_INNER", "").split(",") if v]
names = {12: "unit: before tcgen05.ld", 13: "unit: accumulators in registers", 14: "unit: requantised + table", 0: "dw:start", 1: "dw:setup done", 2: "params landed", 3: "accumulators released", 4: "epilogue done", 5: "entry", 6: "before params wait",
         8: "dw:loop done", 10: "before end barrier", 11: "after end barrier"}
lead_names = {5: "entry", 1: "params landed", 2: "MMAs committed", 3: "accumulators ready", 4: "released + housekeeping", 10: "before end barrier", 11: "after end barrier"}
for p in inner:
    os.environ["YF_B200_TRACE_PHASE"] = str(p)
    net.fused_trace(True)
    net.enqueue(x, y, n); net.sync()
    st = net.fused_trace(False, read=True)
    t0 = st[p]
    evs = sorted((st[96 + i] - t0, names[i]) for i in names if st[96 + i] >= t0 and st[96 + i] <= st[p + 1] + 100)
    print("  phase %2d %-12s (%d cyc): " % (p, steps[p]["name"], st[p + 1] - st[p]) + "; ".join("%s +%d" % (nm, d) for d, nm in evs))
    evs = sorted((st[112 + i] - t0, lead_names[i]) for i in lead_names if st[112 + i] >= t0 and st[112 + i] <= st[p + 1] + 100)
    if evs:
        print("           control thread: " + "; ".join("%s +%d" % (nm, d) for d, nm in evs))
net.close()
