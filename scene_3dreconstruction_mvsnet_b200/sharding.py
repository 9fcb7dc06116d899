"""Reference-view sharding across ranks.

Every (scan, reference view) pair is an independent forward pass in the reference's evaluation loop
(eval.py:326, one reference view per iteration; datasets/dataloader_eval.py:41-49 builds the metas),
so the multi-GPU decomposition is a static partition of the meta list: one process per GPU, rank r
takes metas i with i % world_size == r, weights replicated.  There is no data-path collective; the
only (optional) communication is gathering the [h,w] depth / confidence maps on rank 0.
"""
import os

import torch
import torch.distributed as dist


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_indices(n_items, rank, world_size):
    """Indices of the metas owned by `rank` (round-robin, like a DistributedSampler without padding)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of size %d" % (rank, world_size))
    return list(range(rank, n_items, world_size))


def shard_metas(metas, rank, world_size):
    return [metas[i] for i in shard_indices(len(metas), rank, world_size)]


def sweep_metas(n_scans, views_per_scan):
    """The (scan, ref_view) list of a full sweep, e.g. DTU test: 22 scans x 49 views (lists/dtu/test.txt)."""
    return [(s, v) for s in range(n_scans) for v in range(views_per_scan)]


def gather_maps(local_indices, local_maps, n_items, dst=0):
    """Collect per-view result maps on rank `dst`.  local_maps: list of tensors [k,h,w] aligned with
    local_indices.  Returns a list of n_items tensors on dst (None elsewhere).  Works on gloo (CPU
    tensors) and nccl (CUDA tensors); ranks may own different numbers of views."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        out = [None] * n_items
        for i, m in zip(local_indices, local_maps):
            out[i] = m
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    payload = (list(local_indices), [m.cpu() for m in local_maps])
    gathered = [None] * world if rank == dst else None
    dist.gather_object(payload, gathered, dst=dst)
    if rank != dst:
        return None
    out = [None] * n_items
    for idxs, maps in gathered:
        for i, m in zip(idxs, maps):
            out[i] = m
    return out
