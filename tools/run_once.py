#!/usr/bin/env python
"""Run the hot path a few times on device-resident data (for ncu captures).
usage: run_once.py [n_images] [mode] [reps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pkg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mode = sys.argv[2] if len(sys.argv) > 2 else "auto"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
yf = pkg.load()
net = yf.Network(chunk_images=max(n, 256), mode=mode)
imgs = np.load(os.path.join(ROOT, "tests", "golden", "images_56.npy"))
x = torch.from_numpy(np.concatenate([imgs] * (n // 27 + 1))[:n].copy()).cuda()
y = torch.empty((n, 7, 7, 18), dtype=torch.int8, device="cuda")
for _ in range(reps):
    net.enqueue(x, y, n)
net.sync()
print("ok", net.stats())
net.close()
