"""PPO self-play on the device: the reference's ``run_update`` cycle (RL/robust_train.py:95-156) — gather rollouts, recompute
values / GAE / normalised advantages every epoch, 10 epochs x 64 minibatches of clipped-surrogate updates
(RL/ppo/ppo.py:25-80, RL/ppo/process_batch.py:106-200, RL/ppo/arguments.py:4-116) — with every stage resident in HBM.

Reference                                                   Here
----------------------------------------------------------  -------------------------------------------------------------
128 worker processes x 5 envs, CPU policy at batch 1,       ``VecCatanEnv`` (N envs, one launch set per tick); ``CatanPolicy.act`` on all N
rollouts pickled over pipes (vec_gather_experience.py)      envs per tick; the WHOLE tick (policy inputs -> 12-head sampling pass -> env step
                                                            -> rollout record) is ONE CUDA graph replayed until every env holds T+1 obs
python lists -> stacked CPU tensors (process_rollouts)      ``RolloutStorage``: time-major device buffers written by ``catan_rollout_store``
values in chunks of 10 envs, GAE + advantage norm on CPU    ``get_value`` over (T+1)*N rows in large chunks; ``catan_gae`` + ``catan_adv_*``
minibatch = CPU index + .to(device) per key                 ``catan_minibatch_gather`` + ``catan_policy_inputs``: two launches per minibatch
loss / Adam                                                 the same loss; fused Adam; gradients live in ONE flat buffer
single GPU                                                  sharded envs; the flat gradient buffer (7.7 MB) is all-reduced over NCCL once per
                                                            optimiser step and the advantage statistics (24 B) once per epoch

A minibatch of T*N/64 rows (409 600 at 131 072 envs) is processed in micro-batches whose losses are weighted by their share of
the minibatch, so the accumulated gradient is the minibatch gradient of the reference's mean losses.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Optional

import torch

from . import layout as L
from .policy_io import PolicyInputs
from .policy_net import CatanPolicy
from .rollout import RolloutStorage, gae, normalise_advantages
from .vec_env import VecCatanEnv


@dataclass
class PPOConfig:
    """defaults = RL/ppo/arguments.py:4-116"""
    lr: float = 3e-4
    eps: float = 1e-5
    gamma: float = 0.999
    gae_lambda: float = 0.95
    clip_param: float = 0.2
    ppo_epoch: int = 10
    num_mini_batch: int = 64
    value_loss_coef: float = 1.0
    entropy_coef: float = 0.04
    max_grad_norm: float = 0.5
    num_steps: int = 200
    use_linear_lr_decay: bool = True
    total_updates: int = 7812                  # 1e9 env steps / (200 x 640) in the reference
    micro_batch: int = 32768                   # rows per forward / backward pass of the update
    value_chunk: int = 65536                   # rows per no-grad value pass
    dtype: torch.dtype = torch.float32         # torch.bfloat16: autocast for the policy (BASELINE config 4)
    graph: bool = True                         # replay the rollout tick as a CUDA graph


class SelfPlayTrainer:
    """all four seats of every env are played by the learner (the reference's state at update 0, where the three opponent
    snapshots are copies of the central policy: robust_train.py:62-66); only the decisions of each env's "active" seat are
    recorded (game_manager.py:26, :102-105)."""

    def __init__(self, n_envs: int, policy: Optional[CatanPolicy] = None, cfg: Optional[PPOConfig] = None, device="cuda:0", seed: int = 0,
                 first_env_id: int = 0, group=None, **env_config):
        self.cfg = cfg or PPOConfig()
        self.device = torch.device(device)
        self.group = group
        self.N, self.T = int(n_envs), self.cfg.num_steps
        self.policy = (policy or CatanPolicy()).to(self.device)
        if group is not None:                                          # every rank starts from rank 0's weights
            import torch.distributed as dist
            for p in self.policy.parameters():
                dist.broadcast(p.data, src=0, group=None if group is True else group)
        self.env = VecCatanEnv(self.N, device=self.device, seed=seed, first_env_id=first_env_id, **env_config)
        self.env.reset()
        g = torch.Generator(device=self.device).manual_seed(seed + 17 * (first_env_id + 1))
        active = torch.randint(1, 5, (self.N,), device=self.device, generator=g).to(torch.uint8)   # game_manager.py:24-27
        self.store = RolloutStorage(self.env, self.T, active)
        self.inputs = PolicyInputs(self.N, self.device, self.cfg.dtype)
        self.actions = torch.zeros((self.N, L.ACTION_WORDS), dtype=torch.int32, device=self.device)
        self.logp = torch.zeros(self.N, dtype=torch.float32, device=self.device)
        self.stepped = torch.ones(self.N, dtype=torch.uint8, device=self.device)
        self.env_steps = torch.zeros((), dtype=torch.int64, device=self.device)      # env steps actually taken (frozen envs excluded)
        self._flag = torch.zeros(1, dtype=torch.uint8).pin_memory()
        self._flag_event = torch.cuda.Event()
        self._graph = None
        self._fresh = True
        # the update's buffers
        self.values = torch.zeros((self.T + 1, self.N), dtype=torch.float32, device=self.device)
        mb = self.T * self.N // self.cfg.num_mini_batch
        self.micro = max(1, min(self.cfg.micro_batch, mb))
        self.mb_inputs = PolicyInputs(max(self.micro, min(self.cfg.value_chunk, (self.T + 1) * self.N)), self.device, self.cfg.dtype)
        params = [p for p in self.policy.parameters()]
        self.flat_grad = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=self.device)
        off = 0
        for p in params:                                               # gradients accumulate straight into the flat bucket
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.optimiser = torch.optim.Adam(params, lr=self.cfg.lr, eps=self.cfg.eps, fused=self.device.type == "cuda")
        self.update_num = 0
        self.stats = {}
        self._mb_bufs = {}

    # ------------------------------------------------------------------ rollouts (game_manager.py:69-140, vectorised)
    def _autocast(self):
        return torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.cfg.dtype == torch.bfloat16)

    def _tick(self) -> None:
        self.stepped.copy_(self.store.collecting)
        obs, masks = self.inputs(self.env.obs, self.env.masks)
        with self._autocast():
            _, rows, logp = self.policy.act(obs, masks)
        self.actions.copy_(rows)
        self.logp.copy_(logp.view(-1))
        self.env.step(self.actions, step_mask=self.stepped)
        self.store.record(self.actions, self.logp, self.stepped)
        self.env_steps += self.stepped.sum()

    def _capture(self) -> None:
        self.policy.eval()
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(3):                                         # warm-up on a side stream (cuBLAS workspaces, autotuning)
                self._tick()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._tick()

    def collect(self, max_ticks: int = 100000) -> int:
        """ticks until every env holds T+1 observations of its active seat (game_manager.py:78)"""
        self.policy.eval()
        self.store.begin(fresh=self._fresh)
        self._fresh = False
        if self.cfg.graph and self._graph is None:
            self._capture()                                            # (its warm-up ticks are part of this rollout)
        ticks, pending = 0, False
        while ticks < max_ticks:
            if self._graph is not None:
                self._graph.replay()
            else:
                self._tick()
            ticks += 1
            if ticks % 8 == 0:
                # "has every env filled its quota?" without stalling the launch queue: the flag of the PREVIOUS check is read once
                # its copy has landed; the loop overruns by at most 16 ticks in which every env is frozen
                if pending and self._flag_event.query():
                    if not self._flag[0]:
                        break
                    pending = False
                if not pending:
                    self._flag.copy_(self.store.collecting.any().view(1).to(torch.uint8), non_blocking=True)
                    self._flag_event.record()
                    pending = True
        torch.cuda.current_stream(self.device).synchronize()
        assert not bool(self.store.collecting.any()), "rollout did not finish within max_ticks"
        return ticks

    # ------------------------------------------------------------------ advantages (process_batch.py:106-142)
    @torch.no_grad()
    def compute_advantages(self):
        self.policy.eval()
        rows = self.store.obs.view(-1, L.OBS_STRIDE)
        out = self.values.view(-1)
        chunk = self.mb_inputs.capacity
        for a in range(0, rows.shape[0], chunk):
            b = min(rows.shape[0], a + chunk)
            obs, _ = self.mb_inputs(rows[a:b])
            with self._autocast():
                v = self.policy.get_value(obs)
            out[a:b] = self.policy.value_normaliser.denormalise(v.float().view(-1))
        returns, adv = gae(self.store.rewards, self.values, self.store.tmasks, self.cfg.gamma, self.cfg.gae_lambda)
        normalise_advantages(adv, group=self.group)
        return returns, adv

    def _mb_out(self, rows: int) -> dict:
        """gather buffers of a micro-batch size, allocated once"""
        hit = self._mb_bufs.get(rows)
        if hit is None:
            dev = self.device
            hit = {"obs": torch.empty((rows, L.OBS_STRIDE), dtype=torch.uint8, device=dev),
                   "masks": torch.empty((rows, L.MASK_STRIDE), dtype=torch.uint8, device=dev),
                   "actions": torch.empty((rows, L.ACTION_WORDS), dtype=torch.int32, device=dev)}
            for k in ("logp", "values", "returns", "tmasks", "advantages"):
                hit[k] = torch.empty(rows, dtype=torch.float32, device=dev)
            self._mb_bufs[rows] = hit
        return hit

    # ------------------------------------------------------------------ PPO.update (ppo.py:25-80)
    def update(self):
        cfg, pol = self.cfg, self.policy
        norm = pol.value_normaliser
        if cfg.use_linear_lr_decay:                                    # RL/ppo/utils.py update_linear_schedule
            lr = cfg.lr - cfg.lr * (self.update_num / float(cfg.total_updates))
            for g in self.optimiser.param_groups:
                g["lr"] = lr
        batch = self.T * self.N
        mb = batch // cfg.num_mini_batch
        sums = torch.zeros(3, device=self.device)
        n_steps = 0
        for _ in range(cfg.ppo_epoch):
            returns, adv = self.compute_advantages()
            perm = torch.randperm(batch, device=self.device).to(torch.int32)
            pol.train()
            for k in range(cfg.num_mini_batch):
                idx = perm[k * mb:(k + 1) * mb]
                self.flat_grad.zero_()
                for a in range(0, mb, self.micro):
                    part = idx[a:a + self.micro].contiguous()
                    m = self.store.gather(part, self.values, returns, adv, out=self._mb_out(part.numel()))
                    obs, masks = self.mb_inputs(m["obs"], m["masks"])
                    with self._autocast():
                        values, logp, entropy = pol.evaluate_actions(obs, masks, m["actions"])
                    w = part.numel() / float(mb)
                    value_preds = norm.normalise(m["values"]).view(-1, 1)
                    rets = norm.normalise(m["returns"]).view(-1, 1)
                    ratio = torch.exp(logp - m["logp"].view(-1, 1))
                    adv_t = m["advantages"].view(-1, 1)
                    action_loss = -torch.min(ratio * adv_t, torch.clamp(ratio, 1.0 - cfg.clip_param, 1.0 + cfg.clip_param) * adv_t).mean()
                    clipped = value_preds + (values - value_preds).clamp(-cfg.clip_param, cfg.clip_param)
                    value_loss = 0.5 * torch.max((values - rets).pow(2), (clipped - rets).pow(2)).mean()
                    ((value_loss * cfg.value_loss_coef + action_loss - entropy * cfg.entropy_coef) * w).backward()
                    sums += torch.stack((value_loss.detach(), action_loss.detach(), entropy.detach())) * w
                if self.group is not None:                              # one flat NCCL bucket per optimiser step
                    import torch.distributed as dist
                    dist.all_reduce(self.flat_grad, op=dist.ReduceOp.AVG, group=None if self.group is True else self.group)
                # clip_grad_norm_(parameters, max_grad_norm) on the flat bucket, without a host sync
                total = torch.linalg.vector_norm(self.flat_grad)
                self.flat_grad.mul_(torch.clamp(cfg.max_grad_norm / (total + 1e-6), max=1.0))
                self.optimiser.step()
                n_steps += 1
        self.update_num += 1
        s = (sums / max(1, n_steps)).tolist()
        return {"value_loss": s[0] * cfg.value_loss_coef, "action_loss": s[1], "entropy": s[2], "optimiser_steps": n_steps}

    def warmup_update(self, passes: int = 2) -> None:
        """forward + backward of a few micro-batches of the current rollout WITHOUT an optimiser step (cuBLAS heuristics, allocator
        pools, autograd graph caches), so that a timed ``update`` does not pay for them; weights and Adam state are untouched"""
        returns, adv = self.compute_advantages()
        self.policy.train()
        perm = torch.randperm(self.T * self.N, device=self.device).to(torch.int32)
        for k in range(passes):
            idx = perm[k * self.micro:(k + 1) * self.micro].contiguous()
            m = self.store.gather(idx, self.values, returns, adv, out=self._mb_out(idx.numel()))
            obs, masks = self.mb_inputs(m["obs"], m["masks"])
            with self._autocast():
                values, logp, entropy = self.policy.evaluate_actions(obs, masks, m["actions"])
            (values.mean() + logp.mean() + entropy).backward()
        self.flat_grad.zero_()

    # ------------------------------------------------------------------ one run_update cycle (robust_train.py:95-156)
    def run_update(self) -> dict:
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        steps0 = int(self.env_steps.item())
        ev[0].record()
        ticks = self.collect()
        ev[1].record()
        out = self.update()
        ev[2].record()
        torch.cuda.synchronize(self.device)
        steps = int(self.env_steps.item()) - steps0
        out.update(ticks=ticks, env_steps=steps, recorded_decisions=self.T * self.N, collect_ms=ev[0].elapsed_time(ev[1]),
                   update_ms=ev[1].elapsed_time(ev[2]), games_finished=int((self.store.tmasks == 0).sum().item()))
        self.stats = out
        return out
