"""Batch sharding across GPUs for the yoloface path: contiguous image ranges, one
process per GPU, NO data-path collective -- each rank runs its shard independently and only the
(tiny) detection lists are gathered on rank 0.  torch.distributed is plumbing (rendezvous, barrier,
gather of a few bytes); with the `gloo` backend this runs on CPU for tests."""
import numpy as np


def shard_bounds(n_images, world_size, rank):
    """Contiguous split: rank g owns images [lo, hi); the first n % world ranks get one extra image."""
    if world_size < 1 or not (0 <= rank < world_size) or n_images < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n_images, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_detections(dets, counts):
    """[n, max_det, 5] + [n] -> (flat [total, 5] float32, counts int32): what travels between ranks."""
    dets = np.asarray(dets, np.float32); counts = np.asarray(counts, np.int32)
    if not len(counts):
        return np.zeros((0, 5), np.float32), counts
    keep = np.arange(dets.shape[1])[None, :] < counts[:, None]      # image-major, slot order preserved
    return dets[keep].reshape(-1, 5), counts


def gather_detections(flat, counts, dist=None, dst=0, group=None):
    """Gather every rank's packed detections on `dst` in image order.  Returns (flat, counts) on dst,
    None elsewhere.  Without an initialised process group this is the identity (single GPU).
    `group`: the process group to use -- pass a HOST (gloo) group when the default group is NCCL: the detections are
    already in host memory and go from host to host, no GPU collective is involved."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return flat, counts
    world, rank = dist.get_world_size(), dist.get_rank()
    bucket = [None] * world if rank == dst else None
    dist.gather_object((flat, counts), bucket, dst=dst, group=group)
    if rank != dst:
        return None
    return (np.concatenate([b[0] for b in bucket], axis=0), np.concatenate([b[1] for b in bucket], axis=0))
