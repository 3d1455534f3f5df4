"""One process per GPU: env-index sharding and the (tiny) collectives the path needs.

Environments are independent (overcooked_mdp.py:1375-1430 reads only its own state), so the step
path has NO data-path collective: rank r owns a contiguous range of environment indices and the
layout table is replicated.  torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU
tests) carries only the run seed and the final counters.
"""
import os

import torch
import torch.distributed as dist


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init(backend=None):
    rank, ws, local = world()
    if ws > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend)
    return rank, ws, local


def shard_range(n_total, rank, world_size):
    """Contiguous shard [begin, end) of rank; sizes differ by at most one."""
    return n_total * rank // world_size, n_total * (rank + 1) // world_size


def shard_segments(n_total, n_layouts, rank, world_size):
    """For a mixed batch stored as n_layouts contiguous GLOBAL segments: the layout index of every
    env in this rank's shard (so each rank keeps its layouts contiguous too)."""
    import numpy as np

    b, e = shard_range(n_total, rank, world_size)
    bounds = [n_total * i // n_layouts for i in range(n_layouts + 1)]
    idx = np.arange(b, e)
    return (np.searchsorted(np.array(bounds[1:]), idx, side="right")).astype(np.int32)


def broadcast_seed(seed, device="cpu"):
    """Rank 0's 64-bit seed to everybody."""
    if not (dist.is_available() and dist.is_initialized()):
        return int(seed)
    t = torch.tensor([int(seed)], dtype=torch.int64, device=device)
    dist.broadcast(t, src=0)
    return int(t.item())


def reduce_counters(steps, elapsed_ms, reward_sum, device="cpu"):
    """(sum of steps, max of elapsed, sum of rewards) over ranks."""
    if not (dist.is_available() and dist.is_initialized()):
        return steps, elapsed_ms, reward_sum
    s = torch.tensor([float(steps), float(reward_sum)], dtype=torch.float64, device=device)
    m = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return s[0].item(), m[0].item(), s[1].item()


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
