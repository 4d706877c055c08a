"""PPO self-play on the device (BASELINE configs 3-5 in miniature): the CUDA-graphed tick, the recorded rollout against a
re-evaluation of the same rows, the advantage pass against the reference's lines, and whole ``run_update`` cycles."""
import numpy as np
import pytest
import torch

from settlers_of_catan_rl_b200 import layout as L

pytestmark = pytest.mark.gpu


def _trainer(graph, dtype=torch.float32, n=192, T=12, **kw):
    from settlers_of_catan_rl_b200 import SelfPlayTrainer, PPOConfig, CatanPolicy
    torch.manual_seed(0)
    cfg = PPOConfig(num_steps=T, ppo_epoch=2, num_mini_batch=4, micro_batch=300, value_chunk=1000, dtype=dtype, graph=graph, **kw)
    return SelfPlayTrainer(n, CatanPolicy(), cfg, seed=3, first_env_id=900)


@pytest.mark.parametrize("graph", [False, True])
def test_rollout_rows_re_evaluate_to_the_recorded_log_probs(graph):
    """what the graphed tick recorded (obs, masks, actions, log-probs of the fused sampling pass) is consistent: evaluating the
    stored rows again with ``evaluate_actions`` (torch ops, the PPO update's path) reproduces the stored log-probs, i.e. the
    PPO ratio is 1 before the first optimiser step; every recorded action was accepted by the env."""
    tr = _trainer(graph)
    ticks = tr.collect()
    assert ticks >= tr.T and int(tr.env.err_flags().any()) == 0
    assert tr.store.cursors[:, 0].min().item() == tr.T + 1
    returns, adv = tr.compute_advantages()
    idx = torch.arange(tr.T * tr.N, device="cuda", dtype=torch.int32)
    m = tr.store.gather(idx, tr.values, returns, adv)
    obs, masks = tr.mb_inputs(m["obs"], m["masks"]) if tr.mb_inputs.capacity >= idx.numel() else (None, None)
    if obs is None:
        from settlers_of_catan_rl_b200 import PolicyInputs
        obs, masks = PolicyInputs(idx.numel())(m["obs"], m["masks"])
    tr.policy.eval()
    with torch.no_grad():
        values, logp, ent = tr.policy.evaluate_actions(obs, masks, m["actions"])
    assert torch.isfinite(logp).all()
    torch.testing.assert_close(logp.view(-1), m["logp"], rtol=1e-4, atol=1e-4)
    # the advantage pass == process_batch.py:131-142 on the same values
    v = tr.values.cpu()
    r, tm = tr.store.rewards.cpu(), tr.store.tmasks.cpu()
    ret = torch.zeros_like(r)
    g = 0
    for step in reversed(range(tr.T)):
        delta = r[step] + 0.999 * v[step + 1] * tm[step + 1] - v[step]
        g = delta + 0.999 * 0.95 * tm[step + 1] * g
        ret[step] = g + v[step]
    a = ret - v[:-1]
    a = (a - a.mean()) / (a.std() + 1e-5)
    assert torch.equal(returns.cpu(), ret)
    torch.testing.assert_close(adv.cpu(), a, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_run_update_cycles(dtype):
    tr = _trainer(True, dtype)
    before = torch.cat([p.detach().view(-1).clone() for p in tr.policy.parameters()])
    for k in range(2):
        out = tr.run_update()
        assert out["optimiser_steps"] == 2 * 4 and out["env_steps"] >= tr.T * tr.N and out["ticks"] >= tr.T
        assert all(np.isfinite(out[x]) for x in ("value_loss", "action_loss", "entropy"))
    after = torch.cat([p.detach().view(-1) for p in tr.policy.parameters()])
    assert torch.isfinite(after).all() and float((after - before).abs().max()) > 0
    assert int(tr.env.err_flags().any()) == 0
    # the gradients live in the flat bucket
    p0 = next(tr.policy.parameters())
    assert p0.grad.data_ptr() == tr.flat_grad.data_ptr()


def test_league_seats_with_catan_policies():
    """SeatPolicies (game_manager.py:21-31) with four CatanPolicy networks: routed batches, legal actions, collector records"""
    from settlers_of_catan_rl_b200 import VecCatanEnv, SeatPolicies, RolloutStorage, CatanPolicy
    torch.manual_seed(1)
    pols = [CatanPolicy().cuda().eval() for _ in range(4)]

    def fn(p):
        def f(obs, masks):
            _, rows, logp = p.act(obs, masks)
            return rows, logp
        return f
    env = VecCatanEnv(256, seed=5)
    env.reset()
    sp = SeatPolicies(env, [fn(p) for p in pols])
    store = RolloutStorage(env, 6, sp.active_pid)
    store.begin(fresh=True)
    sp.active = store.collecting
    ticks = store.collect(lambda e: sp.act(e))
    assert ticks >= 6 and int(env.err_flags().any()) == 0
    assert store.cursors[:, 0].min().item() == 7
