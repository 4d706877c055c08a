"""GAE scan + advantage normalisation kernels against the reference's own lines run with torch on CPU
(RL/ppo/process_batch.py:134-142).  Tolerance: returns / raw advantages bit-exact (same IEEE fp32 op order);
normalised advantages rtol 1e-5 (north_star), because torch reduces mean/std in a different order."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def reference_gae(rewards, values, masks, gamma, lam):
    """verbatim semantics of process_batch.py:134-142 on CPU fp32 tensors [T(+1), N, 1]"""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    gae = 0
    for step in reversed(range(T)):
        delta = rewards[step] + gamma * values[step + 1] * masks[step + 1] - values[step]
        gae = delta + gamma * lam * masks[step + 1] * gae
        returns[step] = gae + values[step]
    advantages = returns - values[:-1]
    normed = (advantages - advantages.mean()) / (advantages.std() + 1e-5)
    return returns, advantages, normed


def _inputs(T, N, seed):
    g = torch.Generator().manual_seed(seed)
    rewards = torch.zeros(T, N, 1)
    rewards[torch.rand(T, N, 1, generator=g) < 0.01] = 500.0
    rewards += (torch.rand(T, N, 1, generator=g) < 0.05).float() * torch.randn(T, N, 1, generator=g) * 5
    values = 150.0 + 150.0 * torch.tanh(torch.randn(T + 1, N, 1, generator=g))
    masks = (torch.rand(T + 1, N, 1, generator=g) > 0.02).float()
    return rewards, values, masks


@pytest.mark.parametrize("T,N", [(200, 640), (200, 4099), (7, 1), (1, 33)])
def test_gae_matches_reference_lines(T, N):
    from settlers_of_catan_rl_b200 import gae, normalise_advantages
    rewards, values, masks = _inputs(T, N, T * 1000 + N)
    want_ret, want_adv, want_norm = reference_gae(rewards, values, masks, 0.999, 0.95)
    ret, adv = gae(rewards.cuda(), values.cuda(), masks.cuda(), 0.999, 0.95)
    assert torch.equal(ret.cpu(), want_ret), (ret.cpu() - want_ret).abs().max()
    assert torch.equal(adv.cpu(), want_adv)
    if T * N > 1:
        normalise_advantages(adv)
        torch.testing.assert_close(adv.cpu(), want_norm, rtol=1e-5, atol=1e-5)


def test_gae_full_size_properties():
    """BASELINE config 4 size (T=200, N=131 072): lambda=1, gamma=1, no terminals => returns are suffix sums + bootstrap;
    normalised advantages have mean 0 / unbiased std 1."""
    from settlers_of_catan_rl_b200 import gae, normalise_advantages
    T, N = 200, 131072
    g = torch.Generator(device="cuda").manual_seed(5)
    rewards = torch.randint(0, 3, (T, N), generator=g, device="cuda").float()
    values = torch.randint(0, 8, (T + 1, N), generator=g, device="cuda").float()
    masks = torch.ones(T + 1, N, device="cuda")
    ret, adv = gae(rewards, values, masks, 1.0, 1.0)
    want = torch.flip(torch.cumsum(torch.flip(rewards, [0]), 0), [0]) + values[-1]
    assert torch.equal(ret, want)          # small integers: exact in fp32
    assert torch.equal(adv, ret - values[:-1])
    normalise_advantages(adv)
    assert abs(adv.double().mean().item()) < 1e-6
    assert abs(adv.double().std().item() - 1.0) < 1e-4
