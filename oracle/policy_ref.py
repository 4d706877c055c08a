"""TEST INFRASTRUCTURE — plain torch fp32 restatement of what the reference does to an observation dict and a mask list
before its policy network sees them (RL/models/policy.py:168-190 obs_to_torch / act_masks_to_torch, batched as
RL/ppo/process_batch.py:43-51, :80-84), applied to the packed rows.  Works on any device; the product kernel
(catan_policy_inputs) is compared with it bit for bit, and tests/test_policy_io_vs_reference.py pins it against the real
reference policy network fed with the reference's own conversions.  Never imported by the package.
"""
from __future__ import annotations

import numpy as np
import torch

from settlers_of_catan_rl_b200 import layout as L


def rows_to_policy_inputs(obs_rows: torch.Tensor, mask_rows: torch.Tensor = None, dtype=torch.float32):
    B = obs_rows.shape[0]
    f = obs_rows[:, :L.OBS_FEATURES].to(torch.float32)
    for col, div in L.OBS_RATIO_COLUMNS:                     # wrapper.py:613-627: len / 8.0, knights / 4.0
        f[:, col] = f[:, col] / div
    f = f.to(dtype)
    obs = {}
    for key, off, shape in L.OBS_NUMERIC:                    # policy.py:169-173, :180-183: float32, leading batch dim
        obs[key] = f[:, off:off + int(np.prod(shape))].reshape(B, *shape)
    for key, li in L.OBS_LISTS:                              # process_batch.py:43-47: pad_sequence(batch_first) of long lists
        a = L.OBS_DEV_LISTS + li * L.OBS_DEV_PAD
        obs[key] = obs_rows[:, a:a + L.OBS_DEV_PAD].long()
    if mask_rows is None:
        return obs, None
    masks = []
    for h, (off, shape) in enumerate(L.MASK_HEADS):
        m = mask_rows[:, off:off + int(np.prod(shape))].to(torch.float32).to(dtype).reshape(B, *shape)
        if h in (1, 6, 9):                                   # policy.py:188-189: [1, types, dim] -> [types, 1, dim]; cat on dim 1
            m = m.transpose(0, 1).contiguous()
        masks.append(m)
    return obs, masks
