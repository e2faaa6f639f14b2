#!/usr/bin/env python
"""Where the time of yf_b200_detect goes: device-resident images -> heads (fused kernel) -> decode + NMS kernel ->
detections copied to the host.  usage: python tools/detect_probe.py [images]   (run under ncu for per-kernel times)"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pkg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
yf = pkg.load()
net = yf.Network(chunk_images=8192)
imgs = np.load(os.path.join(ROOT, "tests", "golden", "images_56.npy"))
x = torch.randint(-128, 128, (n, 56, 56, 3), dtype=torch.int8)
x[::2] = torch.from_numpy(imgs)[torch.arange((n + 1) // 2) % len(imgs)]
xd = x.cuda()
for _ in range(2):
    d, c = net.detect(xd, 0.7, 0.4, max_det=16, n=n)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); d, c = net.detect(xd, 0.7, 0.4, max_det=16, n=n); ts.append(time.perf_counter() - t0)
t = float(np.median(ts))
y = torch.empty((n, 7, 7, 18), dtype=torch.int8, device="cuda")
th = []
for _ in range(5):
    t0 = time.perf_counter(); net.run(xd, y, n=n); torch.cuda.synchronize(); th.append(time.perf_counter() - t0)
print("detect(%d device images): %.2f ms = %.2f M img/s, %d detections; heads only: %.2f ms" % (n, t * 1e3, n / t / 1e6, int(c.sum()), float(np.median(th)) * 1e3))
net.close()
