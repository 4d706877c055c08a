"""``VecCatanEnv`` — N four-player Catan games advanced in lock-step on one B200.

The vector counterpart of the reference's ``EnvWrapper`` (env/wrapper.py:11-50): ``reset()`` /
``step(actions)`` keep their meaning, but observations, legal-action masks, rewards and step info are
written by the kernels straight into PyTorch-owned CUDA tensors (``obs`` uint8 [N, 1920], ``masks``
uint8 [N, 336], ``reward`` fp32 [N, 4], ``info`` uint8 [N, 16]); nothing crosses PCIe inside a
rollout.  PyTorch is only plumbing here (device memory + the current stream).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib, layout as L


#: launches of one step: transition, encode of the rows, encode of the masks + sampler (caller's stream) | longest-road search, encode of the
#: searched games (library stream 1) | encode of the games that ended and were reset (library stream 2)
LAUNCHES_PER_STEP = 6


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


class VecCatanEnv:
    def __init__(self, n_envs: int, device="cuda:0", seed: int = 0, first_env_id: int = 0, **config):
        if not torch.cuda.is_available():
            raise _lib.CatanError("VecCatanEnv needs a CUDA device: there is no CPU implementation of the engine")
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CatanError("VecCatanEnv runs on CUDA devices only")
        self.n_envs = int(n_envs)
        self.seed, self.first_env_id = int(seed), int(first_env_id)
        self.config = _lib.make_config(**config)
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._h = C.c_void_p()
        _lib.check(self.lib.catan_create(self.n_envs, idx, self.seed, self.first_env_id, C.byref(self.config),
                                         C.byref(self._h)))
        dev = self.device
        self.obs = torch.zeros((self.n_envs, L.OBS_STRIDE), dtype=torch.uint8, device=dev)
        self.masks = torch.zeros((self.n_envs, L.MASK_STRIDE), dtype=torch.uint8, device=dev)
        self.reward = torch.zeros((self.n_envs, 4), dtype=torch.float32, device=dev)
        self.info = torch.zeros((self.n_envs, L.INFO_STRIDE), dtype=torch.uint8, device=dev)
        _lib.check(self.lib.catan_bind(self._h, _ptr(self.obs), _ptr(self.masks), _ptr(self.reward), _ptr(self.info)))
        self.kernel_launches = 0

    # ------------------------------------------------------------------ lifecycle
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.catan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ EnvWrapper surface, vectorised
    def reset(self, reset_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """EnvWrapper.reset for all envs (or those with a non-zero byte in ``reset_mask``)."""
        if reset_mask is not None:
            assert reset_mask.dtype == torch.uint8 and reset_mask.is_cuda and reset_mask.numel() == self.n_envs
        _lib.check(self.lib.catan_reset(self._h, _ptr(reset_mask), self._stream()))
        self.kernel_launches += 1
        return self.obs

    def step(self, actions: torch.Tensor, step_mask: Optional[torch.Tensor] = None):
        """EnvWrapper.step + get_action_masks for all envs.  ``actions``: int32 CUDA tensor [N, 20].
        ``step_mask`` (uint8 [N], optional): envs with a zero byte are frozen (state and outputs untouched)."""
        assert actions.dtype == torch.int32 and actions.is_cuda and actions.is_contiguous()
        assert actions.shape == (self.n_envs, L.ACTION_WORDS)
        if step_mask is not None:
            assert step_mask.dtype == torch.uint8 and step_mask.is_cuda and step_mask.numel() == self.n_envs
        _lib.check(self.lib.catan_step_masked(self._h, _ptr(actions), _ptr(step_mask), self._stream()))
        self.kernel_launches += LAUNCHES_PER_STEP
        return self.obs, self.reward, self.info[:, L.INFO_DONE], self.info

    def sample_random(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty((self.n_envs, L.ACTION_WORDS), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.catan_sample_random(self._h, _ptr(out), self._stream()))
        self.kernel_launches += 1
        return out

    def step_sample(self, actions_io: torch.Tensor):
        """One call (six launches on three streams, replayed as one CUDA graph with set_graphs; the sampler fused into the masks launch): apply ``actions_io`` and overwrite it with the next random-legal actions."""
        assert actions_io.dtype == torch.int32 and actions_io.is_cuda and actions_io.is_contiguous()
        _lib.check(self.lib.catan_step_sample(self._h, _ptr(actions_io), self._stream()))
        self.kernel_launches += LAUNCHES_PER_STEP
        return self.obs, self.reward, self.info[:, L.INFO_DONE], self.info

    def get_action_masks(self) -> torch.Tensor:
        return self.masks

    # ------------------------------------------------------------------ host-buffer path (the e2e call)
    def step_host(self, actions: np.ndarray, obs: np.ndarray = None, masks: np.ndarray = None, reward: np.ndarray = None,
                  info: np.ndarray = None) -> None:
        def p(a):
            return C.c_void_p(0 if a is None else a.ctypes.data)
        assert actions.dtype == np.int32 and actions.flags.c_contiguous
        _lib.check(self.lib.catan_step_host(self._h, p(actions), p(obs), p(masks), p(reward), p(info), self._stream()))
        self.kernel_launches += LAUNCHES_PER_STEP

    def step_host_async(self, actions: np.ndarray, obs: np.ndarray = None, masks: np.ndarray = None, reward: np.ndarray = None,
                        info: np.ndarray = None) -> None:
        """``step_host`` without the final synchronisation: every buffer must be pinned; the results are in place once the
        current stream has been synchronised.  Two envs on two streams overlap one's copies with the other's kernels."""
        def p(a):
            return C.c_void_p(0 if a is None else a.ctypes.data)
        assert actions.dtype == np.int32 and actions.flags.c_contiguous
        _lib.check(self.lib.catan_step_host_async(self._h, p(actions), p(obs), p(masks), p(reward), p(info), self._stream()))
        self.kernel_launches += LAUNCHES_PER_STEP

    def step_sample_host_async(self, actions_io: np.ndarray, reward: np.ndarray = None, info: np.ndarray = None) -> None:
        """``step_sample`` with pinned host buffers, not synchronised: ``actions_io`` is applied and overwritten with the next
        random-legal actions; valid once the current stream has been synchronised.  int32 rows, or uint8 rows (one byte per
        word, 255 = -1: the compact transport, catan_step_sample_host_async_u8)."""
        def p(a):
            return C.c_void_p(0 if a is None else a.ctypes.data)
        assert actions_io.dtype in (np.int32, np.uint8) and actions_io.flags.c_contiguous
        fn = self.lib.catan_step_sample_host_async if actions_io.dtype == np.int32 else self.lib.catan_step_sample_host_async_u8
        _lib.check(fn(self._h, p(actions_io), p(reward), p(info), self._stream()))
        self.kernel_launches += LAUNCHES_PER_STEP + (2 if actions_io.dtype == np.uint8 else 0)

    def reset_host(self, obs: np.ndarray = None, masks: np.ndarray = None, info: np.ndarray = None) -> None:
        def p(a):
            return C.c_void_p(0 if a is None else a.ctypes.data)
        _lib.check(self.lib.catan_reset_host(self._h, p(obs), p(masks), p(info), self._stream()))
        self.kernel_launches += 1

    # ------------------------------------------------------------------ state export / import (save_state / restore_state)
    def export_state(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.n_envs - first if count is None else count
        out = np.zeros((count, L.STATE_WORDS), dtype=np.int16)
        _lib.check(self.lib.catan_export_state(self._h, first, count, C.c_void_p(out.ctypes.data)))
        return out

    def import_state(self, states: np.ndarray, first: int = 0) -> None:
        states = np.ascontiguousarray(states, dtype=np.int16).reshape(-1, L.STATE_WORDS)
        _lib.check(self.lib.catan_import_state(self._h, first, states.shape[0], C.c_void_p(states.ctypes.data)))
        self.kernel_launches += 1

    def randomise_uncertainty(self, controlling_pid, max_attempts: int = 10000) -> None:
        """Game.randomise_uncertainty (game/game.py:1207-1282) for every env: ``controlling_pid`` = a PlayerId for all envs, or a uint8
        CUDA tensor [N] (0 = leave that env alone).  Obs / mask rows are refreshed; envs without a consistent deal get error bit 9."""
        if not torch.is_tensor(controlling_pid):
            controlling_pid = torch.full((self.n_envs,), int(controlling_pid), dtype=torch.uint8, device=self.device)
        assert controlling_pid.dtype == torch.uint8 and controlling_pid.is_cuda and controlling_pid.numel() == self.n_envs
        _lib.check(self.lib.catan_randomise_uncertainty(self._h, _ptr(controlling_pid), int(max_attempts), self._stream()))
        self.kernel_launches += 2

    def rows_host(self):
        """(obs rows, mask rows) of all envs as numpy arrays (a D2H copy; used by the single-env adapter after an import)"""
        return self.obs.cpu().numpy(), self.masks.cpu().numpy()

    def route_by_policy(self, policy_map: torch.Tensor, n_policies: int, active: Optional[torch.Tensor] = None):
        """game_manager.py:21-31 / :82-93 vectorised: ``policy_map`` uint8 [N,4] = policy index playing PlayerId p+1 in env n.
        Returns (counts int32 [K], lists int32 [K,N]): ``lists[k, :counts[k]]`` are the envs (ascending) whose next decision
        belongs to policy k; ``active`` (uint8 [N], optional) leaves frozen envs out."""
        assert policy_map.dtype == torch.uint8 and policy_map.is_cuda and policy_map.is_contiguous() and policy_map.shape == (self.n_envs, 4)
        if active is not None:
            assert active.dtype == torch.uint8 and active.is_cuda and active.numel() == self.n_envs
        counts = torch.empty(n_policies, dtype=torch.int32, device=self.device)
        lists = torch.empty((n_policies, self.n_envs), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.catan_route_by_policy(_ptr(self.info), _ptr(policy_map), _ptr(active), self.n_envs, int(n_policies),
                                                  _ptr(counts), _ptr(lists), self._stream()))
        return counts, lists

    def lr_stats(self) -> np.ndarray:
        """catan_read_lr_stats: [updates, searched by a block, full enumerations, search tasks, search cycles sum, max,
        walk steps sum, max] since construction"""
        out = np.zeros(8, dtype=np.uint64)
        _lib.check(self.lib.catan_read_lr_stats(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def lr_histograms(self) -> np.ndarray:
        """catan_read_lr_histograms: [3, 24] log2 histograms (cycles per search, walk steps per search, cycles of a step's longest search)"""
        out = np.zeros((3, 24), dtype=np.uint64)
        _lib.check(self.lib.catan_read_lr_histograms(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def set_graphs(self, enable: bool) -> None:
        """replay every distinct step call as one CUDA graph (catan_set_graphs): pass the same tensors / pinned buffers every tick"""
        _lib.check(self.lib.catan_set_graphs(self._h, int(enable)))

    def set_timing(self, enable: bool) -> None:
        """CUDA events around the transition and encode kernels of every following step (catan_set_timing)"""
        _lib.check(self.lib.catan_set_timing(self._h, int(enable)))

    def read_timing(self, detail: bool = False):
        """(steps timed, average ms of transition_kernel, average ms of the two encode launches) since set_timing(True);
        ``detail``: a fourth entry, the average ms of the observation-rows launch alone"""
        out = np.zeros(9, dtype=np.float64)
        _lib.check(self.lib.catan_read_timing(self._h, C.c_void_p(out.ctypes.data)))
        n = max(1.0, out[0])
        self.stream_timing = {"search_wait": out[4] / n, "search": out[5] / n, "search_rest": out[6] / n, "reset_stream": out[7] / n,
                              "after_transition": out[8] / n}     # (average ms per step, see catan_read_timing)
        return (int(out[0]), out[1] / n, out[2] / n) + ((out[3] / n,) if detail else ())

    def err_flags(self, clear: bool = False) -> np.ndarray:
        out = np.zeros(self.n_envs, dtype=np.uint32)
        _lib.check(self.lib.catan_read_err_flags(self._h, C.c_void_p(out.ctypes.data), int(clear)))
        return out

    def set_reward_annealing_factor(self, factor: float) -> None:
        """EnvWrapper.reward_annealing_factor (RL/ppo/game_manager.py:164-166)."""
        self.config.reward_annealing_factor = float(factor)
        _lib.check(self.lib.catan_set_config(self._h, C.byref(self.config)))

    # ------------------------------------------------------------------ views for the policy
    def obs_views(self) -> Dict[str, torch.Tensor]:
        """Zero-copy uint8 views of the packed observation, keyed like the reference's obs dict."""
        out = {}
        for key, off, shape in L.OBS_NUMERIC:
            n = int(np.prod(shape))
            out[key] = self.obs[:, off:off + n].view(self.n_envs, *shape)
        for key, li in L.OBS_LISTS:
            a = L.OBS_DEV_LISTS + li * L.OBS_DEV_PAD
            out[key] = self.obs[:, a:a + L.OBS_DEV_PAD]
        out["player_id"] = self.obs[:, L.OBS_META]
        return out

    def obs_float(self, dtype=torch.float32) -> Dict[str, torch.Tensor]:
        """The observation as the float tensors ``policy.obs_to_torch`` would produce (RL/models/policy.py:168-183):
        numeric blocks cast to ``dtype`` with the two ratio features rescaled; card lists as int64 [N, 25]."""
        f = self.obs[:, :L.OBS_FEATURES].to(dtype)
        for col, div in L.OBS_RATIO_COLUMNS:
            f[:, col] *= 1.0 / div
        out = {}
        for key, off, shape in L.OBS_NUMERIC:
            n = int(np.prod(shape))
            out[key] = f[:, off:off + n].view(self.n_envs, *shape)
        for key, li in L.OBS_LISTS:
            a = L.OBS_DEV_LISTS + li * L.OBS_DEV_PAD
            out[key] = self.obs[:, a:a + L.OBS_DEV_PAD].long()
        return out

    def mask_views(self):
        """List of 12 uint8 views shaped like EnvWrapper.get_action_masks() with a leading env dim."""
        return [self.masks[:, off:off + int(np.prod(shape))].view(self.n_envs, *shape) for off, shape in L.MASK_HEADS]


class HostEnvGroups:
    """Several ``VecCatanEnv`` handles kept in flight from the host, each on its own stream with its own pinned host buffers
    (the sub-process manager's pipelining, RL/ppo/vec_gather_experience.py: while the host looks at one group's result, the other
    groups' kernels and PCIe copies run).  ``pump()`` is ONE library call per round (``catan_step_sample_host_groups``): for every
    group in turn it waits for the group's previous step, reads the ``done`` flags of its info rows, and issues its next step from
    the actions in ``actions[g]`` (pinned; overwritten with the next random-legal actions, as ``step_sample_host_async``)."""

    def __init__(self, envs, packed_actions: bool = False):
        self.envs = list(envs)
        self.packed = bool(packed_actions)      # the one-byte-per-word host transport of action rows (catan_step_sample_host_async_u8)
        G = len(self.envs)
        dev = self.envs[0].device
        self.lib = self.envs[0].lib
        self.streams = [torch.cuda.Stream(device=e.device) for e in self.envs]
        self.actions = [torch.empty((e.n_envs, L.ACTION_WORDS), dtype=torch.uint8 if self.packed else torch.int32).pin_memory() for e in self.envs]
        self.reward = [torch.empty((e.n_envs, 4), dtype=torch.float32).pin_memory() for e in self.envs]
        self.info = [torch.zeros((e.n_envs, L.INFO_STRIDE), dtype=torch.uint8).pin_memory() for e in self.envs]
        arr = C.c_void_p * G
        self._envs = arr(*[e._h.value for e in self.envs])
        self._act = arr(*[t.data_ptr() for t in self.actions])
        self._rew = arr(*[t.data_ptr() for t in self.reward])
        self._info = arr(*[t.data_ptr() for t in self.info])
        self._streams = arr(*[s.cuda_stream for s in self.streams])
        self.done_seen = C.c_longlong(0)
        self.device = dev

    def prime(self) -> None:
        """the first actions of every group: sampled on the device, brought to the host"""
        for e, s, a in zip(self.envs, self.streams, self.actions):
            with torch.cuda.stream(s):
                d = e.sample_random()
                if self.packed:
                    d = torch.where(d < 0, torch.full_like(d, 255), d).to(torch.uint8)
                a.copy_(d, non_blocking=True)

    def pump(self, rounds: int = 1) -> None:
        _lib.check(self.lib.catan_step_sample_host_groups(self._envs, len(self.envs), self._act, 1 if self.packed else 0, self._rew, self._info,
                                                          self._streams, int(rounds), C.byref(self.done_seen)))
        for e in self.envs:
            e.kernel_launches += (LAUNCHES_PER_STEP + (2 if self.packed else 0)) * int(rounds)   # (+ unpack / pack of the byte rows)

    def synchronize(self) -> None:
        for s in self.streams:
            s.synchronize()
