"""Multi-GPU host logic: games shard by global env id, one process per GPU.

The env path has no collective.  The only exchange on the hot path is the 24-byte sum-all-reduce of the
advantage statistics (count, sum, sum of squares) so that every rank normalises with the reference's
global mean / unbiased std (RL/ppo/process_batch.py:141-142).
"""
from __future__ import annotations

from typing import Tuple

import torch


def shard_range(n_total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(first_env_id, n_envs) of `rank`: contiguous, sizes differ by at most one, union is [0, n_total)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_total, world_size)
    n = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, n


def adv_stats_from_tensor(adv: torch.Tensor) -> torch.Tensor:
    """(count, sum, sumsq) in fp64 — the host twin of catan_adv_stats for tensors that are not on a GPU."""
    a = adv.double().reshape(-1)
    return torch.stack([torch.tensor(float(a.numel()), dtype=torch.float64), a.sum(), (a * a).sum()])


def allreduce_adv_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """sum-all-reduce the three doubles in place (NCCL on GPU tensors, gloo on CPU tensors)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def mean_std_from_stats(stats: torch.Tensor) -> Tuple[float, float]:
    """mean and UNBIASED std (torch.Tensor.std default) from (count, sum, sumsq)."""
    n, s, ss = (float(x) for x in stats)
    mean = s / n
    var = max((ss - s * mean) / (n - 1.0), 0.0)
    return mean, var ** 0.5
