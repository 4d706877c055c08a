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


def _reference_head(logits, mask):
    """RL/distributions.py:11-40, restated: Categorical.forward + FixedCategorical"""
    d = torch.distributions.Categorical(logits=logits if mask is None else logits + torch.log(mask))
    p = d.probs.masked_fill(d.probs <= 0, 1)
    return d, -1 * p.mul(p.log()).sum(-1)


@pytest.mark.parametrize("D", [2, 5, 13, 54, 73, 200])
def test_masked_categorical_matches_the_reference_distribution(D):
    from settlers_of_catan_rl_b200.policy_io import masked_categorical
    g = torch.Generator(device="cuda").manual_seed(D)
    B = 4099
    logits = (torch.randn(B, D, device="cuda", generator=g) * 3).contiguous()
    mask = (torch.rand(B, D, device="cuda", generator=g) < 0.4).float()
    mask[torch.arange(B), torch.randint(0, D, (B,), device="cuda", generator=g)] = 1.0     # at least one legal entry
    for mk in (mask, None):
        d, ent = _reference_head(logits, mk)
        # mode (deterministic=True: FixedCategorical.mode)
        a, lp, en = masked_categorical(logits, mk, deterministic=True)
        assert torch.equal(a, d.probs.argmax(dim=-1, keepdim=True))
        torch.testing.assert_close(lp.view(-1), d.log_prob(a.view(-1)), rtol=1e-5, atol=1e-5)     # tolerance: fp32 log-softmax
        torch.testing.assert_close(en, ent, rtol=1e-5, atol=1e-5)
        # evaluate: log-probs of given legal actions
        given = torch.multinomial(d.probs, 1)
        a2, lp2, en2 = masked_categorical(logits, mk, actions=given)
        assert torch.equal(a2, given) and torch.equal(en2, en)
        torch.testing.assert_close(lp2.view(-1), d.log_prob(given.view(-1)), rtol=1e-5, atol=1e-5)
        # sample: always legal, log-prob of what was drawn, and u -> action is the inverse CDF
        u = torch.rand(B, device="cuda", generator=g)
        a3, lp3, _ = masked_categorical(logits, mk, uniforms=u)
        if mk is not None:
            assert bool((mk.gather(1, a3) == 1).all())
        torch.testing.assert_close(lp3.view(-1), d.log_prob(a3.view(-1)), rtol=1e-5, atol=1e-5)
        cdf = d.probs.double().cumsum(-1)
        lo = torch.where(a3 > 0, cdf.gather(1, (a3 - 1).clamp(min=0)), torch.zeros_like(cdf[:, :1])).view(-1)
        hi = cdf.gather(1, a3).view(-1)
        assert bool(((u.double() >= lo - 1e-5) & (u.double() <= hi + 1e-5)).all())
        assert bool((masked_categorical(logits, mk, uniforms=torch.zeros(B, device="cuda"))[0].view(-1) ==
                     (d.probs > 0).float().argmax(dim=-1)).all())                              # u = 0 -> the first legal entry
        top = masked_categorical(logits, mk, uniforms=torch.full((B,), 1.0 - 2 ** -24, device="cuda"))[0]
        assert bool((d.probs.gather(1, top) > 0).all())


def test_masked_categorical_sampling_frequencies():
    from settlers_of_catan_rl_b200.policy_io import masked_categorical
    D, B = 13, 400_000
    logits = torch.tensor([0.3, -1.0, 2.0, 0.0, 0.5, -0.5, 1.0, 0.0, 0.0, -2.0, 0.7, 0.1, 0.2], device="cuda").repeat(B, 1).contiguous()
    mask = torch.tensor([1, 0, 1, 1, 0, 1, 1, 0, 0, 1, 1, 0, 1], device="cuda", dtype=torch.float32).repeat(B, 1).contiguous()
    a, _, _ = masked_categorical(logits, mask, generator=torch.Generator(device="cuda").manual_seed(1))
    freq = torch.bincount(a.view(-1), minlength=D).double() / B
    p = torch.softmax(logits[0] + torch.log(mask[0]), -1).double()
    assert bool(((freq - p).abs() <= 5 * (p * (1 - p) / B).sqrt() + 1e-9).all())


def test_fused_categorical_has_the_fixed_categorical_surface():
    from settlers_of_catan_rl_b200.policy_io import FusedCategorical
    g = torch.Generator(device="cuda").manual_seed(9)
    logits = torch.randn(777, 54, device="cuda", generator=g)
    mask = (torch.rand(777, 54, device="cuda", generator=g) < 0.5).float()
    mask[:, 7] = 1.0
    d, ent = _reference_head(logits, mask)
    f = FusedCategorical(logits, mask, generator=g)
    a = f.sample()
    assert a.shape == (777, 1) and a.dtype == torch.int64 and bool((mask.gather(1, a) == 1).all())
    torch.testing.assert_close(f.log_probs(a).view(-1), d.log_prob(a.view(-1)), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(f.entropy(), ent, rtol=1e-5, atol=1e-5)
    other = torch.multinomial(d.probs, 1)
    torch.testing.assert_close(f.log_probs(other).view(-1), d.log_prob(other.view(-1)), rtol=1e-5, atol=1e-5)
    assert torch.equal(FusedCategorical(logits, mask).mode(), d.probs.argmax(dim=-1, keepdim=True))
