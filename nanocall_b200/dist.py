"""Host-side multi-GPU plumbing: reads shard by read across ranks, no data-path collective.
torch.distributed is used only for the barrier and the max-over-ranks of the timings (bench.py)."""
import numpy as np


def shard_bounds(lengths, world):
    """Contiguous read ranges balanced by event count (same rule as nanocall-b200's --gpus sharding,
    host/main.cpp): bounds[r]..bounds[r+1] are rank r's reads."""
    lengths = np.asarray(lengths, dtype=np.int64)
    world = max(1, min(int(world), max(1, lengths.size)))
    total = int(lengths.sum())
    bounds = [0] * (world + 1)
    acc, g = 0, 1
    for i, n in enumerate(lengths):
        if g >= world:
            break
        acc += int(n)
        if acc * world >= total * g:
            bounds[g] = i + 1
            g += 1
    for k in range(g, world + 1):
        bounds[k] = lengths.size
    return bounds


def max_over_ranks(values, device="cpu"):
    """Element-wise MAX of a list of floats over all ranks (identity when not initialised)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def sum_over_ranks(values, device="cpu"):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


def whole_job_rate(events_this_rank, seconds_this_rank, device="cpu"):
    """events/s of the whole job = all ranks' events / the slowest rank's time."""
    total = sum_over_ranks([events_this_rank], device)[0]
    slowest = max_over_ranks([seconds_this_rank], device)[0]
    return total / slowest, total, slowest
