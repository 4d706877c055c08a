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


def test_minibatch_gather_matches_the_reference_indexing():
    """generator_standard (process_batch.py:169-200): view(-1, ...)[indices] of every buffer, for a drop_last permutation"""
    from settlers_of_catan_rl_b200 import VecCatanEnv, RolloutStorage, layout as L
    T, N = 12, 96
    env = VecCatanEnv(N, seed=4)
    env.reset()
    st = RolloutStorage(env, T)
    g = torch.Generator(device="cuda").manual_seed(1)
    st.obs.copy_(torch.randint(0, 255, st.obs.shape, device="cuda", dtype=torch.uint8, generator=g))
    st.masks.copy_(torch.randint(0, 2, st.masks.shape, device="cuda", dtype=torch.uint8, generator=g))
    st.actions.copy_(torch.randint(0, 54, st.actions.shape, device="cuda", dtype=torch.int32, generator=g))
    st.logp.copy_(torch.randn(st.logp.shape, device="cuda", generator=g))
    st.tmasks.copy_((torch.rand(st.tmasks.shape, device="cuda", generator=g) > 0.1).float())
    values = torch.randn(T + 1, N, device="cuda", generator=g)
    returns = torch.randn(T, N, device="cuda", generator=g)
    adv = torch.randn(T, N, device="cuda", generator=g)
    perm = torch.randperm(T * N, device="cuda", generator=g)
    seen = list(st.minibatches(5, values, returns, adv, perm=perm))   # 1152 rows -> 5 x 230, 2 dropped
    assert len(seen) == 5 and seen[0]["obs"].shape == (T * N // 5, L.OBS_STRIDE)
    assert len(list(st.minibatches(4, values, returns, adv))) == 4    # (its own permutation)
    # against torch indexing of the flattened buffers (the reference's own lines)
    size = T * N // 5
    for k, mb in enumerate(seen):
        idx = perm[k * size:(k + 1) * size]
        assert torch.equal(mb["obs"], st.obs[:-1].reshape(-1, L.OBS_STRIDE)[idx])
        assert torch.equal(mb["masks"], st.masks.reshape(-1, L.MASK_STRIDE)[idx])
        assert torch.equal(mb["actions"], st.actions.reshape(-1, L.ACTION_WORDS)[idx])
        assert torch.equal(mb["logp"], st.logp.reshape(-1)[idx])
        assert torch.equal(mb["values"], values[:-1].reshape(-1)[idx])
        assert torch.equal(mb["returns"], returns.reshape(-1)[idx])
        assert torch.equal(mb["tmasks"], st.tmasks[:-1].reshape(-1)[idx])
        assert torch.equal(mb["advantages"], adv.reshape(-1)[idx])
    # empty and single-row batches
    one = st.gather(torch.tensor([T * N - 1], device="cuda", dtype=torch.int32), values, returns, adv)
    assert torch.equal(one["obs"][0], st.obs[T - 1, N - 1]) and float(one["values"][0]) == float(values[T - 1, N - 1])
    assert st.gather(torch.empty(0, device="cuda", dtype=torch.int32), values, returns, adv)["obs"].shape[0] == 0


def test_policy_routing_lists_match_the_reference_maps():
    """game_manager.py:21-31, :82-93: env n is played by policy_map[n][players turn]; here as per-policy env lists"""
    from settlers_of_catan_rl_b200 import VecCatanEnv, layout as L
    n, K = 3000, 4
    env = VecCatanEnv(n, seed=6)
    env.reset()
    acts = env.sample_random()
    g = torch.Generator(device="cuda").manual_seed(3)
    # a random seat -> policy assignment per env, like random.shuffle(order) in initialise()
    pmap = torch.stack([torch.randperm(K, device="cuda", generator=g) for _ in range(n)]).to(torch.uint8).contiguous()
    active = (torch.rand(n, device="cuda", generator=g) > 0.2).to(torch.uint8)
    for tick in range(40):
        env.step_sample(acts)
        for act in (None, active):
            counts, lists = env.route_by_policy(pmap, K, act)
            actor = env.info[:, L.INFO_ACTOR].long()
            pol = pmap.long().gather(1, (actor - 1).view(-1, 1)).view(-1)
            total = 0
            for k in range(K):
                sel = pol == k
                if act is not None:
                    sel = sel & (act != 0)
                want = torch.nonzero(sel).view(-1).to(torch.int32)
                c = int(counts[k])
                assert c == want.numel() and torch.equal(lists[k, :c], want), (tick, k)
                total += c
            assert total == (n if act is None else int(active.sum()))


def test_rollout_store_replays_the_reference_managers_tape():
    """tests/golden/rollout_manager_tape.npz = the UNCHANGED reference ``GamesAndPoliciesManager`` (game_manager.py:34-150) over
    the reference env with dense rewards, nine consecutive rollouts, recorded by oracle/make_rollout_golden.py.  The CUDA env
    plays the same games from the recorded actions (shared Philox stream), ``catan_rollout_store`` collects, and every
    buffer must equal what ``process_rollouts`` (process_batch.py:37-104) would stack from the manager's lists: obs / masks /
    actions / log-probs / terminal masks bit for bit, rewards to fp32 rounding of the reference's double sums."""
    import os
    from settlers_of_catan_rl_b200 import VecCatanEnv, RolloutStorage
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout_manager_tape.npz")))
    T, N, R = int(g["T"]), g["active_pid"].shape[0], g["obs"].shape[0]
    env = VecCatanEnv(N, seed=int(g["seed"]), first_env_id=int(g["first_env_id"]), dense_reward=1)
    env.reset()
    store = RolloutStorage(env, T, torch.from_numpy(g["active_pid"]))
    tape_a = torch.from_numpy(g["tape_actions"]).to(env.device)           # [N, S, 20]
    tape_lp = torch.from_numpy(g["tape_logp"]).to(env.device)
    k = torch.zeros(N, dtype=torch.int64, device=env.device)                # next decision of every env
    rows = torch.arange(N, device=env.device)
    for r in range(R):
        store.begin(fresh=(r == 0))
        ticks = 0
        while not store.finished():
            ticks += 1
            assert ticks < 20000
            stepped = store.collecting.clone()
            kk = k.clamp(max=tape_a.shape[1] - 1)                       # (an env that has consumed its tape is frozen from here on)
            acts = tape_a[rows, kk].contiguous()
            logp = tape_lp[rows, kk].contiguous()
            env.step(acts, step_mask=stepped)
            store.record(acts, logp, stepped)
            k += stepped.long()
        assert int(env.err_flags().any()) == 0
        cur = store.cursors.cpu().numpy()
        assert np.array_equal(cur, g["lengths"][r]), (r, cur, g["lengths"][r])
        assert np.array_equal(store.obs.cpu().numpy(), g["obs"][r]), r
        got_masks, got_act, got_lp = store.masks.cpu().numpy(), store.actions.cpu().numpy(), store.logp.cpu().numpy()
        got_rew, got_tm = store.rewards.cpu().numpy(), store.tmasks.cpu().numpy()
        for e in range(N):
            n_obs, n_act, n_rew, n_tm = (int(x) for x in g["lengths"][r, e])
            assert np.array_equal(got_masks[:min(n_act, T), e], g["masks"][r, :min(n_act, T), e]), (r, e)
            assert np.array_equal(got_act[:min(n_act, T), e], g["actions"][r, :min(n_act, T), e]), (r, e)
            assert np.array_equal(got_lp[:min(n_act, T), e], g["logp"][r, :min(n_act, T), e]), (r, e)
            assert np.allclose(got_rew[:min(n_rew, T), e], g["rewards"][r, :min(n_rew, T), e], rtol=1e-6, atol=1e-6), (r, e)
            assert np.array_equal(got_tm[:min(n_tm, T + 1), e], g["tmasks"][r, :min(n_tm, T + 1), e]), (r, e)
    assert np.array_equal(k.cpu().numpy(), g["tape_len"]), "every recorded decision was consumed, and no more"
