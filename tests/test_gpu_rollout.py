"""catan_rollout_store against a line-by-line restatement of GamesAndPoliciesManager.gather_rollouts
(RL/ppo/game_manager.py:69-140): two consecutive rollouts (fresh, then carried over) of 96 envs under the
random-legal policy with short games forced by a small VP... (games end naturally; T kept small)."""
import numpy as np
import pytest
import torch

from oracle.rollout_ref import RefCollector
from settlers_of_catan_rl_b200 import layout as L

pytestmark = pytest.mark.gpu


def test_rollout_store_matches_reference_collector():
    from settlers_of_catan_rl_b200 import VecCatanEnv, RolloutStorage
    N, T = 96, 40
    env = VecCatanEnv(N, seed=77, first_env_id=4000, win_reward=500.0)
    env.reset()
    # start late in the games so that terminal steps (done / auto-reset) fall inside the recorded window
    acts = env.sample_random()
    for _ in range(900):
        env.step_sample(acts)
    rng = np.random.default_rng(0)
    active = rng.integers(1, 5, size=N).astype(np.uint8)
    store = RolloutStorage(env, T, torch.from_numpy(active))
    refs = [RefCollector(T, int(active[e])) for e in range(N)]

    def snapshot():
        return env.obs.cpu().numpy().copy(), env.masks.cpu().numpy().copy(), env.info.cpu().numpy().copy(), env.reward.cpu().numpy().copy()

    total_done = 0
    for rollout in range(2):
        obs, masks, info, _ = snapshot()
        if rollout == 0:
            store.begin(fresh=True)
            for e in range(N):
                refs[e].reset(int(info[e, L.INFO_ACTOR]), obs[e])
        else:
            store.begin(fresh=False)
            for e in range(N):
                refs[e].after_rollouts()
        ticks = 0
        while not store.finished():
            ticks += 1
            assert ticks < 5000
            stepped = store.collecting.clone()
            st = stepped.cpu().numpy().astype(bool)
            assert [r.collecting() for r in refs] == st.tolist()
            pre_obs, pre_masks, _, _ = snapshot()
            acts = env.sample_random()
            logp = (torch.arange(N, device=env.device, dtype=torch.float32) * 0.001 - ticks)
            env.step(acts, step_mask=stepped)
            store.record(acts, logp, stepped)
            post_obs, post_masks, info, reward = snapshot()
            a_host, lp_host = acts.cpu().numpy(), logp.cpu().numpy()
            for e in np.nonzero(st)[0]:
                done = bool(info[e, L.INFO_DONE])
                total_done += done
                refs[e].tick(int(info[e, L.INFO_ACTED]), pre_masks[e], a_host[e], float(lp_host[e]), reward[e], done,
                             int(info[e, L.INFO_ACTOR_PRE]), int(info[e, L.INFO_ACTOR]), post_obs[e], post_obs[e])
            # frozen envs are untouched by the masked step
            assert np.array_equal(post_obs[~st], pre_obs[~st]) and np.array_equal(post_masks[~st], pre_masks[~st])
        cur = store.cursors.cpu().numpy()
        g_obs, g_masks, g_act = store.obs.cpu().numpy(), store.masks.cpu().numpy(), store.actions.cpu().numpy()
        g_logp, g_rew, g_tm = store.logp.cpu().numpy(), store.rewards.cpu().numpy(), store.tmasks.cpu().numpy()
        for e in range(N):
            r = refs[e]
            assert cur[e].tolist() == [len(r.observations), len(r.actions), len(r.rewards), len(r.terminal_masks)], (rollout, e)
            assert len(r.observations) == T + 1
            for t, o in enumerate(r.observations):
                assert np.array_equal(g_obs[t, e], o), (rollout, e, t)
            for t in range(min(len(r.actions), T)):
                assert np.array_equal(g_act[t, e], r.actions[t]) and g_logp[t, e] == np.float32(r.action_log_probs[t])
                assert np.array_equal(g_masks[t, e], r.action_masks[t]), (rollout, e, t)
            for t in range(min(len(r.rewards), T)):
                assert g_rew[t, e] == np.float32(r.rewards[t]), (rollout, e, t)
            for t in range(min(len(r.terminal_masks), T + 1)):
                assert g_tm[t, e] == np.float32(r.terminal_masks[t]), (rollout, e, t)
    assert total_done > 0, "no game ended inside the recorded window: the done branches were not exercised"
    # the buffers feed the GAE kernel directly
    values = torch.rand(T + 1, N, device=env.device) * 300
    ret, adv = store.compute_returns(values)
    assert ret.shape == (T, N) and torch.isfinite(adv).all()
