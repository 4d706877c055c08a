"""Pins the rollout-collector oracle against the REAL reference, and runs the reference's unchanged
manager over this repo's ``EnvWrapper`` adapter (the drop-in claim of INTEGRATION.md §A).

1. ``oracle/rollout_ref.RefCollector`` == ``GamesAndPoliciesManager.gather_rollouts`` (RL/ppo/game_manager.py:69-140, with
   ``reset`` :34-56 and ``_after_rollouts`` :142-150), list for list, over consecutive rollouts with games ending inside the
   window, dense rewards on.  The manager runs unchanged; only its policies' ``act`` is the pinned random-legal sampler.
2. the committed tape ``tests/golden/rollout_manager_tape.npz`` (what the GPU test replays against ``catan_rollout_store``) is
   what ``oracle/make_rollout_golden.py`` produces from the reference today (first rollouts re-generated and compared).
3. drop-in: the same unchanged manager constructed over ``settlers_of_catan_rl_b200.EnvWrapper`` (backed here by the host
   emulation of the product's game logic, since this container has no GPU) returns exactly the rollouts it returns over the
   reference's own ``EnvWrapper`` for the same games.
"""
import numpy as np
import pytest

from oracle import ref_harness as H

pytestmark = pytest.mark.skipif(not H.reference_available(), reason="reference tree not present")


def _lists_of(c):
    return dict(obs=c.observations, masks=c.action_masks, actions=c.actions, logp=c.action_log_probs, rewards=c.rewards, tmasks=c.terminal_masks)


def test_ref_collector_equals_the_reference_manager():
    from oracle import manager_harness as MH
    from oracle.rollout_ref import RefCollector
    N, T, R = 4, 90, 8
    mgr = MH.make_manager(N, T, seed=5, first_env_id=100, env_kwargs=dict(dense_reward=True), shuffle_seed=1)
    refs = [RefCollector(T, int(mgr.active_player_ids[e])) for e in range(N)]
    pos = [0] * N
    game_ends = 0
    for r in range(R):
        res = MH.rollout_lists(mgr.gather_rollouts())
        for e in range(N):
            pos[e] = MH.replay_tape_through(refs[e], mgr.envs[e].tape, pos[e], fresh=(r == 0))
            got = _lists_of(refs[e])
            for k in got:
                assert list(got[k]) == res[e][k], (r, e, k)
            assert pos[e] == len(mgr.envs[e].tape.events), "the manager took env steps the restatement did not account for"
            assert len(res[e]["obs"]) == T + 1
            game_ends += res[e]["tmasks"][1:].count(0.0)
        mgr._after_rollouts()
    # the branches where restatements drift must have been exercised: games ending on the active seat's own step, on an
    # opponent's step (done_since_prev_turn), and rewards pushed by the `done` branch
    kinds = set()
    for e in range(N):
        active = int(mgr.active_player_ids[e])
        for ev in mgr.envs[e].tape.events:
            if ev[0] == "step" and ev[1]["done"]:
                kinds.add(("own" if ev[1]["actor"] == active else "opp", "next_active" if ev[1]["n_actor"] == active else "next_other"))
    assert game_ends >= 4 and len(kinds) >= 2, (game_ends, kinds)


def test_committed_rollout_tape_is_what_the_reference_produces():
    import os
    from oracle import make_rollout_golden as G
    g = dict(np.load(G.OUT))
    assert int(g["T"]) == G.T and g["obs"].shape[:3] == (G.R, G.T + 1, G.N)
    fresh = G.generate(r=2)                                          # two rollouts are enough to show the recipe is the file's
    for k in ("obs", "masks", "actions", "logp", "rewards", "tmasks", "lengths"):
        assert np.array_equal(fresh[k], g[k][:2]), k
    assert np.array_equal(fresh["active_pid"], g["active_pid"])
    n = fresh["tape_len"]
    for e in range(G.N):
        assert np.array_equal(fresh["tape_actions"][e, :n[e]], g["tape_actions"][e, :n[e]])
    assert (g["tmasks"] == 0).sum() >= 4, "the tape must contain game ends"


def test_reference_manager_runs_unchanged_over_the_adapter():
    from oracle import manager_harness as MH
    from settlers_of_catan_rl_b200 import EnvWrapper as Adapter
    from tests.host_emu.emu_engine import EmuEngine
    N, T, R, seed, first = 2, 70, 6, 11, 40
    ref = MH.make_manager(N, T, seed=seed, first_env_id=first, shuffle_seed=3)
    ours = MH.make_manager(N, T, seed=seed, first_env_id=first, shuffle_seed=3,
                           env_factory=lambda env_id: Adapter(seed=seed, env_id=env_id, _engine=EmuEngine))
    assert [int(p) for p in ref.active_player_ids] == [int(p) for p in ours.active_player_ids]
    ends = 0
    for r in range(R):
        a, b = ref.gather_rollouts(), ours.gather_rollouts()
        la, lb = MH.rollout_lists(a), MH.rollout_lists(b)
        for e in range(N):
            ta, tb = ref.envs[e].tape, ours.envs[e].tape
            assert len(la[e]["obs"]) == len(lb[e]["obs"]) == T + 1
            for sa, sb in zip(la[e]["obs"], lb[e]["obs"]):
                assert np.array_equal(ta.obs_rows[sa], tb.obs_rows[sb]), (r, e)
            for sa, sb in zip(la[e]["masks"], lb[e]["masks"]):
                assert np.array_equal(ta.mask_rows[sa], tb.mask_rows[sb]), (r, e)
            assert [ta.actions[s].tolist() for s in la[e]["actions"]] == [tb.actions[s].tolist() for s in lb[e]["actions"]]
            assert la[e]["logp"] == lb[e]["logp"] and la[e]["rewards"] == lb[e]["rewards"] and la[e]["tmasks"] == lb[e]["tmasks"]
            ends += la[e]["tmasks"][1:].count(0.0)
        # what the manager stored is what the policy network consumes: same tensors, key by key
        obs_a, obs_b = a[0], b[0]
        for e in range(N):
            for oa, ob in zip(obs_a[e], obs_b[e]):
                for key in oa:
                    if key in ("__serial", "player_id"):
                        continue
                    va, vb = oa[key], ob[key]
                    if isinstance(va, list):
                        assert all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(va, vb)), key
                    else:
                        assert np.array_equal(np.asarray(va), np.asarray(vb)), key
        ref._after_rollouts()
        ours._after_rollouts()
        ref._update_annealing_factor(0.5)                            # game_manager.py:164-166 reaches env.reward_annealing_factor
        ours._update_annealing_factor(0.5)
    assert ends >= 1
