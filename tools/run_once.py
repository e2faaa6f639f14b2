#!/usr/bin/env python
"""Run the hot path a few times on device-resident data (for ncu captures).
usage: run_once.py [n_images] [mode] [reps] [input_size]"""
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
hw = int(sys.argv[4]) if len(sys.argv) > 4 else 56
yf = pkg.load()
os.environ.setdefault("YF_B200_GRAPH", "0")          # ncu profiles kernels one by one: plain launches, not graph replays
net = yf.Network(chunk_images=max(n, 256), mode=mode)
imgs = np.load(os.path.join(ROOT, "tests", "golden", "images_56.npy"))
if hw == 56:
    x = torch.from_numpy(np.concatenate([imgs] * (n // 27 + 1))[:n].copy()).cuda()
else:
    net.set_input_size(hw, hw)
    x = torch.randint(-128, 128, (n, hw, hw, 3), dtype=torch.int8, device="cuda")
y = torch.empty((n, hw // 8, hw // 8, 18), dtype=torch.int8, device="cuda")
for _ in range(reps):
    net.enqueue(x, y, n)
net.sync()
print("ok", net.stats())
net.close()
