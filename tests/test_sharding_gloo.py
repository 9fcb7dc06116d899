"""N>1 host logic on CPU: view sharding + the optional final gather over gloo, world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scene_3dreconstruction_mvsnet_b200 import sharding


def test_shard_indices_partition():
    for n in (0, 1, 7, 49, 1078):
        for world in (1, 2, 3, 8):
            parts = [sharding.shard_indices(n, r, world) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    metas = sharding.sweep_metas(22, 49)  # DTU test sweep (BASELINE config 5)
    assert len(metas) == 1078
    assert [len(sharding.shard_metas(metas, r, 8)) for r in range(8)] == [135] * 6 + [134] * 2


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_views = 7  # ragged: rank 0 gets 4 views, rank 1 gets 3
        mine = sharding.shard_indices(n_views, rank, world)
        maps = [torch.full((2, 3, 4), float(i)) for i in mine]  # stand-in for stacked (depth, confidence)
        out = sharding.gather_maps(mine, maps, n_views, dst=0)
        if rank == 0:
            assert len(out) == n_views
            for i, m in enumerate(out):
                assert m is not None and float(m.mean()) == float(i)
        else:
            assert out is None
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t) == 10.0 + world - 1
    finally:
        dist.destroy_process_group()


def test_gather_maps_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port), nprocs=2, join=True)
