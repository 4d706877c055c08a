"""catan_policy_inputs (one launch: packed rows -> the tensors SettlersAgentPolicy.act / evaluate_actions take) against the
torch statement of the reference's conversions (oracle/policy_ref.py, pinned against the reference's policy network in
tests/test_policy_io_vs_reference.py) — bit-exact, fp32 and bf16 — and the action hand-off back into the env."""
import numpy as np
import pytest
import torch

from oracle.policy_ref import rows_to_policy_inputs
from tests.common import load_golden
from settlers_of_catan_rl_b200 import layout as L

pytestmark = pytest.mark.gpu


def _same(got_obs, got_masks, want_obs, want_masks):
    assert set(got_obs) == set(want_obs)
    for k in want_obs:
        assert got_obs[k].shape == want_obs[k].shape and got_obs[k].dtype == want_obs[k].dtype, k
        assert torch.equal(got_obs[k], want_obs[k]), k
    if want_masks is None:
        assert got_masks is None
        return
    assert len(got_masks) == 12
    for h in range(12):
        assert got_masks[h].shape == want_masks[h].shape and got_masks[h].dtype == want_masks[h].dtype, h
        assert got_masks[h].is_contiguous() and torch.equal(got_masks[h], want_masks[h]), h


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_policy_inputs_match_the_reference_conversions(dtype):
    from settlers_of_catan_rl_b200 import VecCatanEnv
    from settlers_of_catan_rl_b200.policy_io import PolicyInputs
    n = 3001
    env = VecCatanEnv(n, seed=11)
    env.reset()
    acts = env.sample_random()
    pin = PolicyInputs(n, env.device, dtype)
    for tick in range(400):
        env.step_sample(acts)
        if tick % 57 == 0 or tick == 399:
            obs, masks = pin(env.obs, env.masks)
            _same(obs, masks, *rows_to_policy_inputs(env.obs, env.masks, dtype))
    # ragged batches out of the same buffers, and the value-only form
    for b in (1, 2, 333):
        rows = env.obs[n - b:].contiguous()
        obs, masks = pin(rows, env.masks[n - b:].contiguous())
        _same(obs, masks, *rows_to_policy_inputs(rows, env.masks[n - b:], dtype))
        obs, masks = pin(rows)
        _same(obs, masks, *rows_to_policy_inputs(rows, None, dtype))
    assert float(pin.features[:, L.OBS_FEATURES:].abs().sum()) == 0.0
    assert pin.kernel_launches > 0


def test_policy_inputs_of_a_golden_row_and_action_round_trip():
    from settlers_of_catan_rl_b200.policy_io import PolicyInputs, actions_to_rows, rows_to_actions
    g = load_golden("default_s1")
    rows = torch.from_numpy(g["obs"][::13].copy()).cuda()
    mrows = torch.from_numpy(g["masks"][::13].copy()).cuda()
    obs, masks = PolicyInputs(rows.shape[0])(rows, mrows)
    want = g["obs"][::13].astype(np.float64)
    for col, div in L.OBS_RATIO_COLUMNS:
        want[:, col] /= div
    for key, off, shape in L.OBS_NUMERIC:
        k = int(np.prod(shape))
        assert np.array_equal(obs[key].reshape(rows.shape[0], -1).double().cpu().numpy(), want[:, off:off + k]), key
    flat = torch.cat([(m.transpose(0, 1) if h in (1, 6, 9) else m).reshape(rows.shape[0], -1) for h, m in enumerate(masks)], dim=1)
    assert np.array_equal(flat.cpu().numpy(), g["masks"][::13][:, :L.MASK_ENTRIES].astype(np.float32))
    a = torch.from_numpy(g["actions"][:500].copy()).cuda()
    heads = rows_to_actions(a)
    assert [tuple(h.shape) for h in heads] == [(500, 1)] * 7 + [(500, 4)] * 2 + [(500, 1)] * 3 and heads[0].dtype == torch.int64
    assert torch.equal(actions_to_rows(heads), a)
    # policy.act returns heads 7 / 8 as lists of four [B, 1] tensors (action_heads_module.py RecurrentResourceActionHead)
    heads[7] = [heads[7][:, k:k + 1] for k in range(4)]
    heads[8] = [heads[8][:, k:k + 1] for k in range(4)]
    assert torch.equal(actions_to_rows(heads), a)
