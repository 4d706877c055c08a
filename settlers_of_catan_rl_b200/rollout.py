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
        if t.data_ptr() % 16:
            raise _lib.CatanError("rollout kernels read 16-byte vectors: pass tensors whose storage offset keeps them 16-byte aligned")


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


class RolloutStorage:
    """Device-resident rollout buffer with the reference's record semantics.

    Replaces the Python lists of ``GamesAndPoliciesManager`` (RL/ppo/game_manager.py:34-150) and their stacking
    in ``BatchProcessor.process_rollouts`` (RL/ppo/process_batch.py:37-104): per env only the decisions of one
    "active" seat are recorded; rewards are summed between that seat's decisions; ``tmasks[t] == 0`` marks an
    observation that is the first of a new game.  Buffers are time-major ``[T(+1), N, ...]`` CUDA tensors written
    by ``catan_rollout_store`` (one launch per tick), so they feed ``gae`` / the minibatch gather directly.
    """

    def __init__(self, env, num_steps: int, active_pid: Optional[torch.Tensor] = None):
        from . import layout as L
        self.env, self.T, self.N = env, int(num_steps), env.n_envs
        dev = env.device
        T, N = self.T, self.N
        self.obs = torch.zeros((T + 1, N, L.OBS_STRIDE), dtype=torch.uint8, device=dev)
        self.masks = torch.zeros((T, N, L.MASK_STRIDE), dtype=torch.uint8, device=dev)
        self.actions = torch.zeros((T, N, L.ACTION_WORDS), dtype=torch.int32, device=dev)
        self.logp = torch.zeros((T, N), dtype=torch.float32, device=dev)
        self.rewards = torch.zeros((T, N), dtype=torch.float32, device=dev)
        self.tmasks = torch.ones((T + 1, N), dtype=torch.float32, device=dev)
        self.cursors = torch.zeros((N, 4), dtype=torch.int32, device=dev)
        self.acc = torch.zeros((N, 4), dtype=torch.float64, device=dev)
        self.flags = torch.zeros(N, dtype=torch.uint8, device=dev)
        self.collecting = torch.ones(N, dtype=torch.uint8, device=dev)
        if active_pid is None:   # game_manager.py:24-27: a random seat per env is the recorded one
            active_pid = torch.randint(1, 5, (N,), device=dev, dtype=torch.uint8)
        self.active_pid = active_pid.to(device=dev, dtype=torch.uint8).contiguous()
        self._c = _lib.CatanRollout(
            obs=self.obs.data_ptr(), masks=self.masks.data_ptr(), actions=self.actions.data_ptr(), logp=self.logp.data_ptr(),
            rewards=self.rewards.data_ptr(), tmasks=self.tmasks.data_ptr(), cursors=self.cursors.data_ptr(),
            acc=self.acc.data_ptr(), flags=self.flags.data_ptr(), active_pid=self.active_pid.data_ptr(),
            collecting=self.collecting.data_ptr(), T=T, N=N)

    def _call(self, actions, logp, stepped, begin, fresh):
        e = self.env
        z = C.c_void_p(0)
        with torch.cuda.device(e.device):
            _lib.check(_lib.load().catan_rollout_store(
                C.byref(self._c), _p(e.obs), _p(e.masks), _p(e.reward), _p(e.info),
                z if actions is None else _p(actions), z if logp is None else _p(logp), z if stepped is None else _p(stepped),
                int(begin), int(fresh), _stream(e.obs)))

    def begin(self, fresh: bool) -> None:
        """fresh=True right after ``env.reset()`` (manager.reset); False between rollouts (_after_rollouts)."""
        self._call(None, None, None, True, fresh)

    def record(self, actions: torch.Tensor, logp: torch.Tensor, stepped: Optional[torch.Tensor] = None) -> None:
        """call once per tick after ``env.step``; ``stepped`` = the step mask used for that tick (None = all)."""
        assert actions.dtype == torch.int32 and logp.dtype == torch.float32 and actions.is_contiguous() and logp.is_contiguous()
        self._call(actions, logp, stepped, False, False)

    def finished(self) -> bool:
        return not bool(self.collecting.any().item())

    def collect(self, policy_fn, max_ticks: int = 1_000_000) -> int:
        """Run ticks until every env holds T+1 observations.  ``policy_fn(env) -> (actions int32 [N,20], logp fp32 [N])``.
        Envs that filled their quota are frozen with the step mask (game_manager.py:78).  Returns the tick count."""
        ticks = 0
        stepped = torch.empty_like(self.collecting)
        while ticks < max_ticks:
            if ticks % 8 == 0 and self.finished():
                break
            actions, logp = policy_fn(self.env)
            stepped.copy_(self.collecting)
            self.env.step(actions, step_mask=stepped)
            self.record(actions, logp, stepped)
            ticks += 1
        return ticks

    def gather(self, indices: torch.Tensor, values: torch.Tensor, returns: torch.Tensor, advantages: torch.Tensor, out: Optional[dict] = None) -> dict:
        """One minibatch (process_batch.py:177-200): rows ``indices`` (int32 CUDA, flat ``t * N + n`` with t < T) of every
        rollout array, gathered by ONE launch into contiguous tensors.  ``values`` [T+1,N], ``returns`` / ``advantages``
        [T,N] fp32.  Keys: obs uint8 [B,1920], masks uint8 [B,336], actions int32 [B,20], logp, values, returns, tmasks,
        advantages fp32 [B]  (slice obs / masks into the policy's keys with ``VecCatanEnv.obs_views`` / ``mask_views``)."""
        from . import layout as L
        _check_f32_cuda(values, returns, advantages)
        if not (indices.is_cuda and indices.dtype == torch.int32 and indices.is_contiguous()):
            raise _lib.CatanError("indices must be a contiguous int32 CUDA tensor")
        if values.numel() != (self.T + 1) * self.N or returns.numel() != self.T * self.N or advantages.numel() != self.T * self.N:
            raise ValueError("values must be [T+1, N], returns and advantages [T, N]")
        B, dev = indices.numel(), self.env.device
        if out is None:
            out = {"obs": torch.empty((B, L.OBS_STRIDE), dtype=torch.uint8, device=dev),
                   "masks": torch.empty((B, L.MASK_STRIDE), dtype=torch.uint8, device=dev),
                   "actions": torch.empty((B, L.ACTION_WORDS), dtype=torch.int32, device=dev)}
            for k in ("logp", "values", "returns", "tmasks", "advantages"):
                out[k] = torch.empty(B, dtype=torch.float32, device=dev)
        mb = _lib.CatanMinibatch(**{k: out[k].data_ptr() for k in ("obs", "masks", "actions", "logp", "values", "returns", "tmasks", "advantages")})
        with torch.cuda.device(dev):
            _lib.check(_lib.load().catan_minibatch_gather(C.byref(self._c), _p(values), _p(returns), _p(advantages), _p(indices), B,
                                                          C.byref(mb), _stream(self.obs)))
        return out

    def minibatches(self, num_mini_batch: int, values: torch.Tensor, returns: torch.Tensor, advantages: torch.Tensor, generator=None,
                    perm: Optional[torch.Tensor] = None):
        """generator_standard (process_batch.py:169-200): a random permutation of the T*N rows cut into ``num_mini_batch``
        index lists of ``T*N // num_mini_batch`` rows (BatchSampler(SubsetRandomSampler(...), drop_last=True))."""
        batch = self.T * self.N
        size = batch // num_mini_batch
        if perm is None:
            perm = torch.randperm(batch, device=self.env.device, generator=generator)
        perm = perm.to(device=self.env.device, dtype=torch.int32)
        for k in range(num_mini_batch):
            yield self.gather(perm[k * size:(k + 1) * size].contiguous(), values, returns, advantages)

    def compute_returns(self, values: torch.Tensor, gamma: float = 0.999, gae_lambda: float = 0.95, normalise: bool = True, group=None):
        """values: [T+1, N] fp32 (denormalised).  Returns (returns, advantages) — process_batch.py:134-142."""
        returns, adv = gae(self.rewards, values.contiguous(), self.tmasks, gamma, gae_lambda)
        if normalise:
            normalise_advantages(adv, group=group)
        return returns, adv
