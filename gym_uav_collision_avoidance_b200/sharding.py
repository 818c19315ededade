"""Multi-GPU plumbing: environments are independent, so the env axis is cut into contiguous shards, one process
(one GPU) per shard, with NO data-path collective.  The only exchange is the optional sum of the episode
counters (episodes, reach, collisions, steps — what the reference scripts read from `env.target_reach_count`,
`env.collision_count`, `env.steps`: test_sac_multi.py:164-165) over `torch.distributed` (NCCL on the GPUs,
gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch

STAT_KEYS = ("episodes", "reach", "collisions", "steps")


def shard_range(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """-> (env_index_base, num_envs) of `rank`: contiguous, sizes differ by at most one, every env owned once."""
    if not (0 <= rank < world) or total_envs < 0:
        raise ValueError("bad shard request")
    q, r = divmod(total_envs, world)
    base = rank * q + min(rank, r)
    return base, q + (1 if rank < r else 0)


def reduce_stats(stats: Dict[str, int], device=None, group=None) -> Dict[str, int]:
    """Sum the episode counters over all ranks (one all_reduce of 4 int64); identity without a process group."""
    import torch.distributed as dist

    out = {k: int(stats.get(k, 0)) for k in STAT_KEYS}
    if not (dist.is_available() and dist.is_initialized()):
        return out
    t = torch.tensor([out[k] for k in STAT_KEYS], dtype=torch.int64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(STAT_KEYS, (int(v) for v in t.tolist())))


def max_over_ranks(value: float, device=None, group=None) -> float:
    """A timing taken on every rank -> the slowest rank's (multi-GPU numbers are max over ranks)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
