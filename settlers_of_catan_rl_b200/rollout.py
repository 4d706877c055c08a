"""Rollout-side PPO pieces of the hot path on the GPU (RL/ppo/process_batch.py:106-142).

``gae`` and ``normalise_advantages`` wrap the sm_100a kernels behind the C ABI; tensors are fp32 CUDA,
time-major ``[T(+1), N]`` or the reference's ``[T(+1), N, 1]``.  When envs are sharded across GPUs the
reference's *global* advantage statistics are recovered by sum-all-reducing three doubles
(count, sum, sum of squares) between the two normalisation kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib


def _p(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


def _stream(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _check_f32_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise _lib.CatanError("rollout kernels take contiguous fp32 CUDA tensors (no CPU path)")


def gae(rewards: torch.Tensor, values: torch.Tensor, masks: torch.Tensor, gamma: float = 0.999, gae_lambda: float = 0.95,
        returns: Optional[torch.Tensor] = None, advantages: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """process_batch.py:134-140.  rewards [T,N(,1)], values/masks [T+1,N(,1)] -> (returns, advantages) like rewards."""
    _check_f32_cuda(rewards, values, masks)
    T = rewards.shape[0]
    N = rewards.numel() // T
    if values.shape[0] != T + 1 or masks.shape[0] != T + 1 or values.numel() != (T + 1) * N or masks.numel() != (T + 1) * N:
        raise ValueError("values and masks must be [T+1, N]")
    returns = torch.empty_like(rewards) if returns is None else returns
    advantages = torch.empty_like(rewards) if advantages is None else advantages
    _check_f32_cuda(returns, advantages)
    lib = _lib.load()
    with torch.cuda.device(rewards.device):
        _lib.check(lib.catan_gae(_p(rewards), _p(values), _p(masks), T, N, float(gamma), float(gae_lambda), _p(returns),
                                 _p(advantages), _stream(rewards)))
    return returns, advantages


def normalise_advantages(advantages: torch.Tensor, eps: float = 1e-5, group=None,
                         stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """process_batch.py:141-142, in place: (A - mean) / (std_unbiased + eps) over ALL elements.

    ``group``: a torch.distributed process group (or True for the default group) over which envs are
    sharded; the (count, sum, sumsq) statistics are sum-all-reduced so every rank normalises with the
    reference's global mean / std."""
    _check_f32_cuda(advantages)
    lib = _lib.load()
    if stats is None:
        stats = torch.empty(3, dtype=torch.float64, device=advantages.device)
    n = advantages.numel()
    with torch.cuda.device(advantages.device):
        _lib.check(lib.catan_adv_stats(_p(advantages), n, _p(stats), _stream(advantages)))
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=None if group is True else group)
        _lib.check(lib.catan_adv_apply(_p(advantages), n, _p(stats), float(eps), _stream(advantages)))
    return advantages
