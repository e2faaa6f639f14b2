#!/usr/bin/env python
"""Aggregate pinned-host -> device copy bandwidth of this box with N GPUs copying AT THE SAME TIME.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/h2d_ceiling.py
    python tools/h2d_ceiling.py                      (one GPU)

The end-to-end figure of bench.py moves 9,408 B per image from pinned host memory to every GPU; at 8 GPUs that is a
demand of 8 x ~45 GB/s on the host's memory system and PCIe root complexes, and whatever the box sustains in aggregate is
the ceiling of e2e (the 0.53 scaling efficiency at N = 8 had never been compared with a measured
ceiling).  Every rank copies `--mb` MiB blocks from its own pinned buffer for `--seconds`, all ranks between barriers;
rank 0 prints one JSON line with the per-rank and aggregate GB/s and the images/s that bandwidth could feed.
bench.py runs the same measurement inline and reports it as e2e.bounds.box_h2d_GBps."""
import argparse
import json
import os
import time

import torch


def measure(seconds=1.5, mb=64, dist=None):
    """-> (this rank's GB/s, aggregate GB/s over all ranks); all ranks copy concurrently"""
    dev = torch.cuda.current_device()
    n = mb << 20
    src = [torch.empty(n, dtype=torch.int8).pin_memory() for _ in range(4)]        # 4 x mb MiB: larger than the CPU's LLC share
    dst = torch.empty(n, dtype=torch.int8, device="cuda")
    for s in src:
        s.fill_(1)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        dst.copy_(src[0], non_blocking=True)
    st.synchronize()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter(); k = 0
    with torch.cuda.stream(st):
        while time.perf_counter() - t0 < seconds:
            for _ in range(8):
                dst.copy_(src[k % 4], non_blocking=True); k += 1
            st.synchronize()
    dt = time.perf_counter() - t0
    mine = k * n / dt / 1e9
    total = mine
    if dist:
        t = torch.tensor([mine], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        total = float(t.item())
        dist.barrier()
    del src, dst
    return mine, total


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--mb", type=int, default=64)
    a = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mine, total = measure(a.seconds, a.mb, dist)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "rank0_GBps": mine, "aggregate_GBps": total, "block_MiB": a.mb,
                          "images_per_s_this_could_feed": total * 1e9 / (56 * 56 * 3)}))
    if dist:
        dist.destroy_process_group()
