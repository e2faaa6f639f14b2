#!/usr/bin/env python
"""Randomised stress of the overlapping paths (kernel lanes, ring, small-batch path) against the CPU oracle.
usage: stress_lanes.py [rounds] [seed]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pkg  # noqa: E402
from oracle_lib import Oracle  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
yf = pkg.load()
o = Oracle()
imgs = np.load(os.path.join(ROOT, "tests", "golden", "images_56.npy"))
pool = rng.integers(-128, 128, (4096, 56, 56, 3), dtype=np.int8)
pool[::2] = imgs[np.arange(2048) % len(imgs)]
ref = o.run_batch(pool, threads=os.cpu_count())
net = yf.Network(chunk_images=256)
stream = torch.cuda.Stream()
bad = 0
for r in range(rounds):
    k = int(rng.integers(1, 12))
    sizes = [int(rng.choice([1, 2, 3, 8, 9, 17, 64, 147, 148, 149, 200, 256, 300, 700])) for _ in range(k)]
    starts = [int(rng.integers(0, 4096 - s)) for s in sizes]
    mode = r % 4
    if mode == 3:                                  # device-resident launches one after the other (latency shape for <= 148 images)
        net.set_stream(stream.cuda_stream)
        with torch.cuda.stream(stream):
            d_in = [torch.from_numpy(pool[a:a + s]).cuda() for a, s in zip(starts, sizes)]
            d_out = [torch.full((s, 7, 7, 18), 77, dtype=torch.int8, device="cuda") for s in sizes]
            stream.synchronize()
            for x, y, s in zip(d_in, d_out, sizes):
                net.enqueue(x, y, s)
        net.sync(); net.set_stream(None)
        got = [t.cpu().numpy() for t in d_out]
    elif mode == 0:                                  # device-resident batches over the lanes
        net.set_stream(stream.cuda_stream)
        with torch.cuda.stream(stream):
            d_in = [torch.from_numpy(pool[a:a + s]).cuda() for a, s in zip(starts, sizes)]
            d_out = [torch.full((s, 7, 7, 18), 77, dtype=torch.int8, device="cuda") for s in sizes]
            stream.synchronize()
            net.enqueue_batches(d_in, d_out, sizes)
            net.enqueue_batches(d_in[::-1], d_out[::-1], sizes[::-1])
        net.sync(); net.set_stream(None)
        got = [t.cpu().numpy() for t in d_out]
    elif mode == 1:                                # pipelined host ring
        h_in = [torch.from_numpy(pool[a:a + s].copy()).pin_memory() for a, s in zip(starts, sizes)]
        h_out = [torch.full((s, 7, 7, 18), 77, dtype=torch.int8).pin_memory() for s in sizes]
        for x, y, s in zip(h_in, h_out, sizes):
            net.submit(x, y, s)
        net.wait()
        got = [t.numpy() for t in h_out]
    else:                                          # blocking calls, pageable memory (small path, pieces)
        got = [net.run(pool[a:a + s].copy()) for a, s in zip(starts, sizes)]
    for g, a, s in zip(got, starts, sizes):
        if not np.array_equal(g, ref[a:a + s]):
            bad += 1
            print("MISMATCH round %d mode %d size %d" % (r, mode, s))
st = net.stats()
print("stress: %d rounds, %d mismatches, launches %d (%d in the latency shape)" % (rounds, bad, st["kernel_launches"], st["latency_launches"]))
net.close()
sys.exit(1 if bad else 0)
