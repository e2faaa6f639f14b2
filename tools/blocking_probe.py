#!/usr/bin/env python
"""Time the blocking host call (yf_b200_run, pinned host in/out) at a few batch sizes.  usage: blocking_probe.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pkg  # noqa: E402

yf = pkg.load()
for n in (1, 8, 16, 32, 256, 1024):
    net = yf.Network(chunk_images=max(n, 256))
    pin = os.environ.get("PROBE_PAGEABLE") is None
    x = torch.randint(-128, 128, (n, 56, 56, 3), dtype=torch.int8)
    y = torch.empty((n, 7, 7, 18), dtype=torch.int8)
    if pin:
        x, y = x.pin_memory(), y.pin_memory()
    for _ in range(20):
        net.run(x, y, n=n)
    t0 = time.perf_counter()
    reps = 300
    for _ in range(reps):
        net.run(x, y, n=n)
    dt = (time.perf_counter() - t0) / reps
    print("small<=%s pinned=%d n=%4d  %.1f us per call  %.2f M img/s" % (os.environ.get("YF_B200_SMALL", "16"), pin, n, dt * 1e6, n / dt / 1e6))
    net.close()
