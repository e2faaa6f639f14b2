#!/usr/bin/env python
"""Throughput of ONE handle spread over 1, 2, 4, 8 GPUs (yf_b200_config.device_mask): blocking yf_b200_run of 65,536
images from pinned host memory to pinned host memory; the library splits the call dynamically over its members.
usage: python tools/one_handle_probe.py [images]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pkg  # noqa: E402

yf = pkg.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
x = torch.randint(-128, 128, (n, 56, 56, 3), dtype=torch.int8).pin_memory()
y = torch.empty((n, 7, 7, 18), dtype=torch.int8).pin_memory()
ng = torch.cuda.device_count()
ref = None
for d in (1, 2, 4, 8):
    if d > ng:
        break
    net = yf.Network(devices=list(range(d)), chunk_images=1024)
    for _ in range(2):
        net.run(x, y, n=n)
    ts = []
    for _ in range(7):
        t0 = time.perf_counter(); net.run(x, y, n=n); ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    if ref is None:
        ref = y.clone()
    same = bool(torch.equal(ref, y))
    print("devices %d  yf_b200_run(%d pinned host images): %.2f M img/s (%.1f ms)  heads equal to the 1-GPU run: %s" % (d, n, n / t / 1e6, t * 1e3, same))
    net.close()
