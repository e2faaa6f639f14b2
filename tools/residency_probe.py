#!/usr/bin/env python
"""Throughput of the fused kernel at 8,192 images per launch (set YF_B200_FUSED_PAD to lower the residency)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pkg  # noqa: E402

yf = pkg.load()
n = 8192
net = yf.Network(chunk_images=n, mode="fused")
x = torch.randint(-128, 128, (n, 56, 56, 3), dtype=torch.int8, device="cuda")
y = torch.empty((n, 7, 7, 18), dtype=torch.int8, device="cuda")
s = torch.cuda.Stream(); net.set_stream(s.cuda_stream)
for _ in range(3):
    net.enqueue(x, y, n)
net.sync()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(s)
for _ in range(10):
    net.enqueue(x, y, n)
b.record(s); b.synchronize(); net.sync()
ms = a.elapsed_time(b) / 10
print("pad=%s smem=%d: %.3f ms per 8192 -> %.2f M img/s" % (os.environ.get("YF_B200_FUSED_PAD", "0"), net.stats()["fused_smem_bytes"], ms, n / ms / 1e3))
net.close()
