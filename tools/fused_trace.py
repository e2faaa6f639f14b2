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
np_ = len(steps)
for im in range(2 if n > 296 else 1):
    base = im * (np_ + 1)
    tot = st[base + np_] - st[base]
    print("n=%d image #%d of CTA 0: total cycles %d" % (n, im, tot))
    for i, s in enumerate(steps):
        d = st[base + i + 1] - st[base + i]
        print("  %-14s kind %d rows %4d cout %2d  %7d cyc  %5.1f%%" % (s["name"], s["kind"], F["phases"][i]["rows_out"], F["phases"][i]["cout"], d, 100.0 * d / tot))
inner = st[96:108]
tph = int(os.environ.get("YF_B200_TRACE_PHASE", "1"))
if F["phases"][tph]["kind"] == 1:
    print("phase %d (conv) thread 0: params there -> MMAs committed %d | housekeeping %d | -> accumulators ready %d | epilogue %d | fence %d | barrier %d" % (
        tph, inner[1] - inner[0], inner[2] - inner[1], inner[3] - inner[2], inner[4] - inner[3], inner[10] - inner[4], inner[11] - inner[10]))
elif F["phases"][tph]["kind"] == 2:
    print("phase %d (depthwise) thread 0: entry->setup %d | loop %d | fence %d | barrier %d" % (tph, inner[1] - inner[0], inner[8] - inner[1], inner[10] - inner[8], inner[11] - inner[10]))
net.close()
