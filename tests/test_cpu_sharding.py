"""N > 1 host logic on CPU: world_size-2 gloo process group, shard ownership and the statistics reduction
that bench.py performs over NCCL on GPUs."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chipmunk2d_b200.sharding import shard_range, space_kind, reduce_step_stats


def test_shards_partition_the_spaces():
    for total in (1, 7, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(total, world, r)
                seen += list(range(lo, hi))
            assert seen == list(range(total))
    kinds = [space_kind(i) for i in range(6)]
    assert kinds == ["PyramidStack", "Chains"] * 3


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(10, world, rank)
    bodies = float(sum(107 if space_kind(i) == "PyramidStack" else 82 for i in range(lo, hi)))
    sums, maxes, t = reduce_step_stats(dist, torch, "cpu", [bodies, 100.0 * (rank + 1)], [0.5 + rank], 10.0 + 5.0 * rank)
    out[rank] = (sums, maxes, t)
    dist.destroy_process_group()


def test_gloo_world_size_2_reduction():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    for rank in (0, 1):
        sums, maxes, t = out[rank]
        assert sums == [5 * 107 + 5 * 82, 300.0]       # every space counted exactly once across ranks
        assert maxes == [1.5] and t == 15.0            # time = max over ranks
