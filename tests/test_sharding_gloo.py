"""N > 1 host logic on CPU: world_size-2 `gloo` ranks shard a batch by image (no data-path
collective) and rank 0 gathers the detections in image order."""
import os
import sys

import numpy as np
import pytest

import pkg

HERE = os.path.dirname(os.path.abspath(__file__))


def test_shard_bounds_cover_the_batch_exactly():
    yf = pkg.load()
    from importlib import import_module  # noqa: F401
    sh = __import__("stm32h7_yolo_b200.sharding", fromlist=["shard_bounds"])
    for n in (0, 1, 7, 256, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [sh.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_bounds(10, 2, 2)
    assert yf is not None


def _worker(rank, world, port, tmp):
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    import pkg as _pkg
    _pkg.load()
    sh = __import__("stm32h7_yolo_b200.sharding", fromlist=["shard_bounds"])
    from oracle_lib import Oracle
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    imgs = np.load(os.path.join(HERE, "golden", "images_56.npy"))
    batch = np.concatenate([imgs, imgs[::-1]])[:37]                      # ragged: 37 images over 2 ranks
    lo, hi = sh.shard_bounds(len(batch), world, rank)
    o = Oracle()                                                          # stand-in for the GPU shard (no GPU here)
    heads = o.run_batch(batch[lo:hi], threads=2)
    dets = np.zeros((hi - lo, 16, 5), np.float32); counts = np.zeros(hi - lo, np.int32)
    for i in range(hi - lo):
        d = o.decode_nms(heads[i], 0.7, 0.4)[:16]
        dets[i, :len(d)] = d; counts[i] = len(d)
    dist.barrier()
    out = sh.gather_detections(*sh.pack_detections(dets, counts), dist=dist)
    if rank == 0:
        np.savez(os.path.join(tmp, "gathered.npz"), flat=out[0], counts=out[1])
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather(tmp_path, oracle, golden):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    imgs = golden["images"]
    batch = np.concatenate([imgs, imgs[::-1]])[:37]
    exp_counts, exp_flat = [], []
    for img in batch:
        d = oracle.decode_nms(oracle.run(img), 0.7, 0.4)[:16]
        exp_counts.append(len(d)); exp_flat.append(d)
    assert np.array_equal(got["counts"], np.array(exp_counts, np.int32))
    assert np.array_equal(got["flat"], np.concatenate(exp_flat, axis=0))
    assert got["counts"].sum() > 10
