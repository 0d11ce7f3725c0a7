"""Multi-GPU layout of the batched workload: independent spaces are partitioned across
ranks in contiguous blocks; no body, pair or contact ever crosses a GPU.  The only inter-rank traffic is
the reduction of timings and step statistics (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_range(n_total, world, rank):
    """Contiguous block [lo, hi) of space indices owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def space_kind(global_index):
    """Config 5 alternates PyramidStack (even) and Chains (odd) by GLOBAL space index, so every shard
    holds the same mix whatever the rank count."""
    return "PyramidStack" if global_index % 2 == 0 else "Chains"


def reduce_step_stats(dist, torch, device, sums, maxes, time_ms):
    """All-reduce (SUM of `sums`, MAX of `maxes` and of the elapsed time).  Returns python lists."""
    s = torch.tensor(list(sums), dtype=torch.float64, device=device)
    m = torch.tensor(list(maxes) + [time_ms], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    m = m.tolist()
    return s.tolist(), m[:-1], m[-1]
