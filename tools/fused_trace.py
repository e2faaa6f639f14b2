#!/usr/bin/env python
"""Per-phase SM-cycle trace of the fused kernel (CTA 0, first image). usage: fused_trace.py [n_images]"""
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
    sub = st[80 + 8 * im:80 + 8 * im + 7]
    print("  phase 14 sub-stamps: param wait %d | mma issue %d | mma wait %d | epilogue %d | fence %d | barrier %d" % tuple(sub[i + 1] - sub[i] for i in range(6)))
inner = st[96:108]
print("inner stamps of phase %s (thread 0): entry->setup %d | setup->iter0 %d | iters %s | loop end->ret %s | fence %d | barrier %d" % (
    os.environ.get("YF_B200_TRACE_PHASE", "1"), inner[1] - inner[0], inner[2] - inner[1], [inner[i + 1] - inner[i] for i in range(2, 7)],
    inner[9] - inner[8], inner[10] - inner[9], inner[11] - inner[10]))
net.close()
