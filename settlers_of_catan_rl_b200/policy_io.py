"""Hand-off between the packed env rows and the reference's policy network.

The reference converts every observation dict and mask list to torch tensors on the host, env by env
(``SettlersAgentPolicy.obs_to_torch`` / ``act_masks_to_torch``, RL/models/policy.py:168-190), stacks them per key
(RL/ppo/process_batch.py:43-51, :80-84) and turns the sampled heads back into numpy for ``env.step``
(``torch_act_to_np``, policy.py:192-199).  Here one kernel launch (``catan_policy_inputs``) expands a batch of packed
uint8 rows — the env's own buffers, a routed subset, or a gathered minibatch — into exactly the tensors
``SettlersAgentPolicy.act / evaluate_actions / get_value`` take, and the sampled heads go back into the int32 [B, 20]
action rows that ``VecCatanEnv.step`` reads, without leaving the device.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, layout as L

FEATURE_STRIDE = 1792           # CATAN_POLICY_FEATURE_STRIDE
TYPE_CONDITIONAL_HEADS = (1, 6, 9)   # policy.py:188-189
_DTYPES = {torch.float32: 0, torch.bfloat16: 1}


def _p(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


class PolicyInputs:
    """Pre-allocated device buffers for batches of up to ``capacity`` rows; ``__call__`` fills them with one launch and returns
    ``(obs_dict, action_masks)`` as views: ``obs_dict[key]`` is ``[B, ...]`` of ``dtype`` for the numeric keys and int64
    ``[B, 25]`` for the five card lists (the padded-tensor form of player_modules.py:49-53); ``action_masks`` is the list of 12
    with heads 1 / 6 / 9 as ``[types, B, dim]``."""

    def __init__(self, capacity: int, device="cuda:0", dtype: torch.dtype = torch.float32):
        if dtype not in _DTYPES:
            raise ValueError("policy inputs are produced in float32 or bfloat16")
        self.lib = _lib.load()
        self.capacity, self.dtype, self.device = int(capacity), dtype, torch.device(device)
        self.features = torch.empty((self.capacity, FEATURE_STRIDE), dtype=dtype, device=self.device)
        self.lists = torch.empty(5 * self.capacity * L.OBS_DEV_PAD, dtype=torch.int64, device=self.device)
        self.head_masks = torch.empty(L.MASK_ENTRIES * self.capacity, dtype=dtype, device=self.device)
        self.kernel_launches = 0
        self._view_cache = {}

    def __call__(self, obs_rows: torch.Tensor, mask_rows: Optional[torch.Tensor] = None,
                 index: Optional[torch.Tensor] = None) -> Tuple[Dict[str, torch.Tensor], Optional[List[torch.Tensor]]]:
        """``index`` (int32 [B], optional): batch row b is row ``index[b]`` of ``obs_rows`` / ``mask_rows`` (one policy's env
        list from ``VecCatanEnv.route_by_policy``); without it the batch is all the rows."""
        R = obs_rows.shape[0]
        assert obs_rows.dtype == torch.uint8 and obs_rows.is_cuda and obs_rows.is_contiguous() and obs_rows.shape == (R, L.OBS_STRIDE)
        if mask_rows is not None:
            assert mask_rows.dtype == torch.uint8 and mask_rows.is_cuda and mask_rows.is_contiguous() and mask_rows.shape == (R, L.MASK_STRIDE)
        B = R
        if index is not None:
            assert index.dtype == torch.int32 and index.is_cuda and index.is_contiguous() and index.dim() == 1
            B = index.shape[0]
        assert 0 < B <= self.capacity
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.catan_policy_inputs(_p(obs_rows), _p(mask_rows), _p(index), B, _DTYPES[self.dtype], _p(self.features), _p(self.lists),
                                                _p(self.head_masks) if mask_rows is not None else C.c_void_p(0), stream))
        self.kernel_launches += 1
        return self._views(B, mask_rows is not None)

    def _views(self, B: int, with_masks: bool):
        """the views of a batch size are built once: a call costs one launch, not two dozen tensor constructions.  The returned
        tensors are VIEWS of this object's buffers: the next call overwrites them (clone what must outlive it)."""
        hit = self._view_cache.get((B, with_masks))
        if hit is not None:
            return hit
        f = self.features[:B]
        obs = {}
        for key, off, shape in L.OBS_NUMERIC:
            obs[key] = f[:, off:off + int(np.prod(shape))].view(B, *shape)
        lists = self.lists[:5 * B * L.OBS_DEV_PAD].view(5, B, L.OBS_DEV_PAD)
        for key, li in L.OBS_LISTS:
            obs[key] = lists[li]
        masks = None
        if with_masks:
            masks = []
            for h, (off, shape) in enumerate(L.MASK_HEADS):
                flat = self.head_masks[off * B:(off + int(np.prod(shape))) * B]
                masks.append(flat.view(shape[0], B, shape[1]) if h in TYPE_CONDITIONAL_HEADS else flat.view(B, shape[0]))
        if len(self._view_cache) >= 64:                                # (per-policy batch sizes change every tick: keep the newest)
            self._view_cache.pop(next(iter(self._view_cache)))
        self._view_cache[(B, with_masks)] = (obs, masks)
        return obs, masks


def actions_to_rows(actions: Sequence, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The 12 sampled heads of ``policy.act`` (``[B, 1]`` long tensors; heads 7 / 8 a list of four of them, or ``[B, 4]``) ->
    int32 ``[B, 20]`` action rows (catan_layout.h; policy.py:192-199 + wrapper.py:114-166 read them in this order)."""
    if isinstance(actions, torch.Tensor):                  # already action rows (CatanPolicy.act)
        assert actions.dim() == 2 and actions.shape[1] == L.ACTION_WORDS
        return actions.to(torch.int32)
    cols = []
    for h in range(12):
        a = actions[h]
        if isinstance(a, (list, tuple)):
            a = torch.cat([x.reshape(-1, 1) for x in a], dim=1)
        cols.append(a.reshape(a.shape[0], -1))
    assert [c.shape[1] for c in cols] == [1] * 7 + [4, 4] + [1] * 3
    B = cols[0].shape[0]
    if out is None:
        out = torch.zeros((B, L.ACTION_WORDS), dtype=torch.int32, device=cols[0].device)
    out[:, :18] = torch.cat(cols, dim=1)
    return out


def rows_to_actions(action_rows: torch.Tensor) -> List[torch.Tensor]:
    """int32 ``[B, 20]`` action rows -> the list of 12 long tensors ``evaluate_actions`` takes (``[B, 1]``; heads 7 / 8
    ``[B, 4]``: process_batch.py:67-75)."""
    a = action_rows.long()
    return [a[:, h:h + 1] for h in range(7)] + [a[:, L.A_GIVE:L.A_GIVE + 4], a[:, L.A_RECV:L.A_RECV + 4]] + \
           [a[:, L.A_RES_A:L.A_RES_A + 1], a[:, L.A_RES_B:L.A_RES_B + 1], a[:, L.A_DISCARD:L.A_DISCARD + 1]]


def masked_categorical(logits: torch.Tensor, mask: Optional[torch.Tensor] = None, actions: Optional[torch.Tensor] = None,
                       deterministic: bool = False, uniforms: Optional[torch.Tensor] = None,
                       generator: Optional[torch.Generator] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The tail of a reference action head in one launch (RL/distributions.py:11-40 ``Categorical.forward`` +
    ``FixedCategorical.sample / mode / log_probs / entropy``): ``logits`` fp32 ``[B, D]``, ``mask`` fp32 ``[B, D]`` (0 = illegal) or
    None.  Returns ``(actions [B, 1] int64, log_probs [B, 1], entropy [B])`` for the given ``actions`` (evaluate), the mode
    (``deterministic``), or a sample drawn by inverse CDF from ``uniforms`` (default: ``torch.rand(B)`` of ``generator``).
    Forward only (rollouts run under ``no_grad``); the PPO update keeps the torch distribution for its gradients."""
    B, D = logits.shape
    logits = logits.detach()
    assert logits.dtype == torch.float32 and logits.is_cuda and logits.is_contiguous()
    if mask is not None:
        mask = mask.detach()
        assert mask.dtype == torch.float32 and mask.is_contiguous() and mask.shape == (B, D) and mask.device == logits.device
    dev = logits.device
    logp = torch.empty((B, 1), dtype=torch.float32, device=dev)
    entropy = torch.empty(B, dtype=torch.float32, device=dev)
    given = out = None
    if actions is not None:
        given = actions.reshape(B).to(torch.int64).contiguous()
    else:
        out = torch.empty((B, 1), dtype=torch.int64, device=dev)
        if not deterministic and uniforms is None:
            uniforms = torch.rand(B, device=dev, generator=generator)
        if deterministic:
            uniforms = None
    if uniforms is not None:
        assert uniforms.dtype == torch.float32 and uniforms.is_contiguous() and uniforms.numel() == B
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(_lib.load().catan_masked_categorical(_p(logits), _p(mask), _p(given), _p(uniforms), B, D, _p(out), _p(logp), _p(entropy), stream))
    return (given.view(B, 1) if out is None else out), logp, entropy


class FusedCategorical:
    """``FixedCategorical``'s surface (RL/distributions.py:11-23: ``sample / mode / log_probs / entropy``) on the fused kernel, for
    the rollout path (no gradients): ``Categorical.forward`` can return ``FusedCategorical(x, mask)`` instead of
    ``FixedCategorical(logits=x + torch.log(mask))`` when ``not torch.is_grad_enabled()``.  The first ``sample()`` / ``mode()``
    computes action, log-prob and entropy in one launch; ``log_probs`` of that same action and ``entropy`` are then free."""

    def __init__(self, logits: torch.Tensor, mask: Optional[torch.Tensor] = None, generator: Optional[torch.Generator] = None):
        self.logits = logits.detach().float().contiguous()
        self.mask = None if mask is None else mask.detach().float().expand_as(self.logits).contiguous()
        self.generator = generator
        self._action = self._logp = self._entropy = None

    def _run(self, **kw):
        self._action, self._logp, self._entropy = masked_categorical(self.logits, self.mask, generator=self.generator, **kw)
        return self._action

    def sample(self) -> torch.Tensor:
        return self._run()

    def mode(self) -> torch.Tensor:
        return self._run(deterministic=True)

    def log_probs(self, actions: torch.Tensor) -> torch.Tensor:
        if self._action is not None and actions is self._action:      # (identity only: a value comparison would be a host sync)
            return self._logp
        return masked_categorical(self.logits, self.mask, actions=actions)[1]

    def entropy(self) -> torch.Tensor:
        if self._entropy is None:
            self._run(deterministic=True)
        return self._entropy
