#!/usr/bin/env python
"""Where does the end-to-end rate go when all GPUs of the box are fed at once?  

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 tools/e2e_probe.py

Every rank runs the same variant at the same time (barriers around each), ~1.5 s per variant; rank 0 prints the
aggregate images/s (or the images/s a raw copy rate could feed) per variant:
  raw_h2d_64M        pinned -> device copies of 64 MiB blocks (tools/h2d_ceiling.py: the box ceiling)
  raw_h2d_2M4        the same with 2.4 MB blocks (one 256-image batch) from a ring of 64 buffers
  raw_h2d_2M4+d2h    ... while a second stream copies 226 KB blocks device -> pinned host (the heads)
  raw_h2d_2M4_2streams   2.4 MB blocks alternating over two streams
  submit_256         yf_b200_submit per 256-image step, continuous (one yf_b200_wait at the end of the variant)
  submit_256_x20     the bench's e2e region: 20 submits + yf_b200_wait, repeated
  submit_1024        yf_b200_submit per 1,024-image step, continuous
  run_8192           blocking yf_b200_run of 8,192 images per call (the library pipelines 1,024-image chunks)
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import pkg  # noqa: E402
import h2d_ceiling  # noqa: E402

IN, OUT = 56 * 56 * 3, 7 * 7 * 18
SECONDS = float(os.environ.get("YF_PROBE_SECONDS", "1.5"))


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def total(v):
        if not dist:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t.item())

    res = {}
    _, agg = h2d_ceiling.measure(SECONDS, 64, dist)
    res["raw_h2d_64M"] = agg * 1e9 / IN
    # ---- raw copies at the batch granularity ----
    h_in = [torch.empty((256, 56, 56, 3), dtype=torch.int8).pin_memory() for _ in range(64)]
    for t in h_in:
        t.fill_(3)
    d_in = [torch.empty((256, 56, 56, 3), dtype=torch.int8, device="cuda") for _ in range(8)]
    h_out = [torch.empty((256, 7, 7, 18), dtype=torch.int8).pin_memory() for _ in range(8)]
    d_out = torch.zeros((256, 7, 7, 18), dtype=torch.int8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for with_d2h in (False, True):
        barrier()
        t0 = time.perf_counter(); k = 0
        while time.perf_counter() - t0 < SECONDS:
            for _ in range(16):
                with torch.cuda.stream(s1):
                    d_in[k % 8].copy_(h_in[k % 64], non_blocking=True)
                if with_d2h:
                    with torch.cuda.stream(s2):
                        h_out[k % 8].copy_(d_out, non_blocking=True)
                k += 1
            s1.synchronize(); s2.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        res["raw_h2d_2M4+d2h" if with_d2h else "raw_h2d_2M4"] = total(k * 256 / dt)
    # two copy streams alternating: does the set-up of one copy hide behind the transfer of the other?
    barrier()
    t0 = time.perf_counter(); k = 0
    while time.perf_counter() - t0 < SECONDS:
        for _ in range(16):
            with torch.cuda.stream(s1 if k & 1 else s2):
                d_in[k % 8].copy_(h_in[k % 64], non_blocking=True)
            k += 1
        s1.synchronize(); s2.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    res["raw_h2d_2M4_2streams"] = total(k * 256 / dt)
    # ---- the library ----
    yf = pkg.load()
    net = yf.Network(device=local, chunk_images=1024)
    for _ in range(4):
        net.submit(h_in[0], h_out[0], 256)
    net.wait()
    barrier()
    t0 = time.perf_counter(); k = 0
    while time.perf_counter() - t0 < SECONDS:
        for _ in range(16):
            net.submit(h_in[k % 64], h_out[k % 8], 256); k += 1
    net.wait()
    dt = time.perf_counter() - t0
    barrier()
    res["submit_256"] = total(k * 256 / dt)
    barrier()
    t0 = time.perf_counter(); k = 0
    while time.perf_counter() - t0 < SECONDS:
        for _ in range(20):
            net.submit(h_in[k % 64], h_out[k % 8], 256); k += 1
        net.wait()
    dt = time.perf_counter() - t0
    barrier()
    res["submit_256_x20"] = total(k * 256 / dt)
    big_in = [torch.empty((1024, 56, 56, 3), dtype=torch.int8).pin_memory() for _ in range(16)]
    big_out = [torch.empty((1024, 7, 7, 18), dtype=torch.int8).pin_memory() for _ in range(8)]
    for t in big_in:
        t.fill_(5)
    net.submit(big_in[0], big_out[0], 1024); net.wait()
    barrier()
    t0 = time.perf_counter(); k = 0
    while time.perf_counter() - t0 < SECONDS:
        for _ in range(8):
            net.submit(big_in[k % 16], big_out[k % 8], 1024); k += 1
    net.wait()
    dt = time.perf_counter() - t0
    barrier()
    res["submit_1024"] = total(k * 1024 / dt)
    x = torch.empty((8192, 56, 56, 3), dtype=torch.int8).pin_memory(); x.fill_(7)
    y = torch.empty((8192, 7, 7, 18), dtype=torch.int8).pin_memory()
    net.run(x, y, n=8192)
    barrier()
    t0 = time.perf_counter(); k = 0
    while time.perf_counter() - t0 < SECONDS:
        net.run(x, y, n=8192); k += 1
    dt = time.perf_counter() - t0
    barrier()
    res["run_8192"] = total(k * 8192 / dt)
    net.close()
    if rank == 0:
        print(json.dumps({"n_gpus": world, "images_per_s": {k: round(v) for k, v in res.items()},
                          "GBps_h2d": {k: round(v * IN / 1e9, 1) for k, v in res.items()}}))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
