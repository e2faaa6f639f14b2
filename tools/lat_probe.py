#!/usr/bin/env python
"""Device-resident launch latency of small batches (CUDA events around one enqueue) and the blocking host call.
usage: python tools/lat_probe.py [n ...]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pkg  # noqa: E402

yf = pkg.load()
net = yf.Network(chunk_images=512)
ns = [int(v) for v in sys.argv[1:]] or [1, 8, 148]
for n in ns:
    x = torch.randint(-128, 128, (n, 56, 56, 3), dtype=torch.int8, device="cuda")
    y = torch.empty((n, 7, 7, 18), dtype=torch.int8, device="cuda")
    st = torch.cuda.Stream(); net.set_stream(st.cuda_stream)
    for _ in range(5):
        net.enqueue(x, y, n)
    net.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(100):
        e0.record(st); net.enqueue(x, y, n); e1.record(st); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    net.set_stream(None)
    xh = x.cpu().numpy()
    for _ in range(5):
        net.run(xh)
    th = []
    for _ in range(200):
        t0 = time.perf_counter(); net.run(xh); th.append((time.perf_counter() - t0) * 1e6)
    print("n=%d  device-resident launch: median %.1f us (min %.1f)   blocking host call: median %.1f us (min %.1f)" %
          (n, float(np.median(ts)), min(ts), float(np.median(th)), min(th)))
net.close()
