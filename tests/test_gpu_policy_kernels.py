"""The two policy-side sm_100a kernels (csrc/policy_kernels.cu) and their gradients against the torch ops they replace."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B", [1, 2, 3, 7, 1000])
def test_tile_attention_matches_torch_forward_and_backward(B):
    from settlers_of_catan_rl_b200.policy_ops import tile_attention
    g = torch.Generator(device="cuda").manual_seed(B)
    qkv = torch.randn(B, 19, 192, device="cuda", generator=g, requires_grad=True)
    ref_in = qkv.detach().clone().requires_grad_(True)
    y = tile_attention(qkv)
    q, k, v = (t.transpose(1, 2) for t in ref_in.view(B, 19, 3, 4, 16).unbind(2))        # [B, 4, 19, 16]
    att = torch.softmax(q @ k.transpose(-1, -2) / 4.0, dim=-1) @ v
    want = att.transpose(1, 2).reshape(B, 19, 64)
    torch.testing.assert_close(y, want, rtol=1e-5, atol=1e-5)
    dy = torch.randn(B, 19, 64, device="cuda", generator=g)
    y.backward(dy)
    want.backward(dy)
    torch.testing.assert_close(qkv.grad, ref_in.grad, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("rows,dim", [(1, 16), (5, 16), (7777, 16), (3, 25), (311296, 25), (2, 64), (100003, 64), (40, 33)])
def test_layer_norm_small_matches_torch_forward_and_backward(rows, dim):
    from settlers_of_catan_rl_b200.policy_ops import layer_norm_small
    g = torch.Generator(device="cuda").manual_seed(dim)
    x = (torch.randn(rows, dim, device="cuda", generator=g) * 3 + 1).requires_grad_(True)
    w = torch.randn(dim, device="cuda", generator=g).requires_grad_(True)
    b = torch.randn(dim, device="cuda", generator=g).requires_grad_(True)
    x2, w2, b2 = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    y = layer_norm_small(x, w, b, 1e-5)
    want = F.layer_norm(x2, (dim,), w2, b2, 1e-5)
    torch.testing.assert_close(y, want, rtol=1e-5, atol=1e-5)
    dy = torch.randn(rows, dim, device="cuda", generator=g)
    y.backward(dy)
    want.backward(dy)
    torch.testing.assert_close(x.grad, x2.grad, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(w.grad, w2.grad, rtol=2e-4, atol=2e-3 * max(1.0, rows ** 0.5 / 30))
    torch.testing.assert_close(b.grad, b2.grad, rtol=2e-4, atol=2e-3 * max(1.0, rows ** 0.5 / 30))
    # 3-D input, no grad
    with torch.no_grad():
        z = layer_norm_small(x.detach().view(1, rows, dim), w.detach(), b.detach())
    torch.testing.assert_close(z.view(rows, dim), want.detach(), rtol=1e-5, atol=1e-5)


def test_policy_on_cuda_matches_policy_on_cpu():
    """the CUDA path of CatanPolicy (fused attention / LayerNorm / categorical kernels) == its plain torch path, which is the one
    pinned against the reference network on CPU (tests/test_policy_net_vs_reference.py)"""
    from settlers_of_catan_rl_b200 import VecCatanEnv, PolicyInputs, CatanPolicy
    torch.manual_seed(0)
    env = VecCatanEnv(512, seed=2)
    env.reset()
    a = env.sample_random()
    for _ in range(300):
        env.step_sample(a)
    obs, masks = PolicyInputs(512)(env.obs, env.masks)
    pol = CatanPolicy().cuda().eval()
    cpu = CatanPolicy().eval()
    cpu.load_state_dict(pol.state_dict())
    with torch.no_grad():
        v, rows, lp = pol.act(obs, masks)
        v2, lp2, ent2 = pol.evaluate_actions(obs, masks, rows)
        obs_c = {k: t.cpu() for k, t in obs.items()}
        masks_c = [m.cpu() for m in masks]
        v3, lp3, ent3 = cpu.evaluate_actions(obs_c, masks_c, rows.cpu())
    torch.testing.assert_close(lp, lp2, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(lp2.cpu(), lp3, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(v2.cpu(), v3, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ent2.cpu(), ent3, rtol=1e-4, atol=1e-4)
    # and the gradients of a PPO-like loss agree
    pol.train(); cpu.train()
    vg, lpg, eg = pol.evaluate_actions(obs, masks, rows)
    (vg.mean() + lpg.mean() + eg).backward()
    vc, lpc, ec = cpu.evaluate_actions(obs_c, masks_c, rows.cpu())
    (vc.mean() + lpc.mean() + ec).backward()
    checked = 0
    for (n, p), (_, q) in zip(pol.named_parameters(), cpu.named_parameters()):
        assert (p.grad is None) == (q.grad is None), n               # (the value normaliser's constants take no gradient)
        if p.grad is not None:
            torch.testing.assert_close(p.grad.cpu(), q.grad, rtol=2e-3, atol=2e-5, msg=lambda m, n=n: n + ": " + m)
            checked += 1
    assert checked > 150
