"""TEST INFRASTRUCTURE — generates ``tests/golden/rollout_manager_tape.npz`` by running the reference's UNCHANGED
``GamesAndPoliciesManager.gather_rollouts`` (RL/ppo/game_manager.py:69-140, + ``reset`` :34-56, ``_after_rollouts`` :142-150)
over the reference's ``EnvWrapper(dense_reward=True)`` under the shared Philox stream, with the pinned random-legal sampler as
the four policies (``oracle/manager_harness.py``).  The file holds, per env, the action / log-prob sequence the stub policy
took (the GPU test feeds exactly these to ``catan_step``) and, per rollout, what the manager returned, stacked the way
``BatchProcessor.process_rollouts`` stacks it (RL/ppo/process_batch.py:37-104): obs ``[R, T+1, N, 1920]`` (packed rows),
masks / actions / log-probs / rewards ``[R, T, N, ...]``, terminal masks ``[R, T+1, N]`` and the list lengths.

    python oracle/make_rollout_golden.py        # needs /root/reference (or oracle/_ref)
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from settlers_of_catan_rl_b200 import layout as L  # noqa: E402
from oracle import manager_harness as MH  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "rollout_manager_tape.npz")
N, T, R, SEED, FIRST = 4, 96, 9, 31, 700


def generate(n=N, t=T, r=R, seed=SEED, first=FIRST):
    mgr = MH.make_manager(n, t, seed=seed, first_env_id=first, env_kwargs=dict(dense_reward=True), shuffle_seed=seed)
    obs = np.zeros((r, t + 1, n, L.OBS_STRIDE), np.uint8)
    masks = np.zeros((r, t, n, L.MASK_STRIDE), np.uint8)
    actions = np.zeros((r, t, n, L.ACTION_WORDS), np.int32)
    logp = np.zeros((r, t, n), np.float32)
    rewards = np.zeros((r, t, n), np.float32)
    tmasks = np.ones((r, t + 1, n), np.float32)
    lengths = np.zeros((r, n, 4), np.int32)                 # len(observations), len(actions), len(rewards), len(terminal_masks)
    for k in range(r):
        res = MH.rollout_lists(mgr.gather_rollouts())
        for e in range(n):
            tape, x = mgr.envs[e].tape, res[e]
            lengths[k, e] = [len(x["obs"]), len(x["actions"]), len(x["rewards"]), len(x["tmasks"])]
            for i, s in enumerate(x["obs"][:t + 1]):
                obs[k, i, e] = tape.obs_rows[s]
            for i, s in enumerate(x["masks"][:t]):
                masks[k, i, e] = tape.mask_rows[s]
            for i, s in enumerate(x["actions"][:t]):
                actions[k, i, e] = tape.actions[s]
            logp[k, :min(t, len(x["logp"])), e] = x["logp"][:t]
            rewards[k, :min(t, len(x["rewards"])), e] = x["rewards"][:t]
            tmasks[k, :min(t + 1, len(x["tmasks"])), e] = x["tmasks"][:t + 1]
        mgr._after_rollouts()
    n_dec = max(len(env.tape.actions) for env in mgr.envs)
    tape_actions = np.zeros((n, n_dec, L.ACTION_WORDS), np.int32)
    tape_logp = np.zeros((n, n_dec), np.float32)
    tape_len = np.zeros(n, np.int32)
    for e, env in enumerate(mgr.envs):
        k = len(env.tape.actions)
        tape_len[e] = k
        tape_actions[e, :k] = np.asarray(env.tape.actions)
        tape_logp[e, :k] = env.tape.logps
    return dict(seed=np.int64(seed), first_env_id=np.int64(first), T=np.int64(t), active_pid=np.asarray([int(p) for p in mgr.active_player_ids], np.uint8),
                tape_actions=tape_actions, tape_logp=tape_logp, tape_len=tape_len, obs=obs, masks=masks, actions=actions, logp=logp,
                rewards=rewards, tmasks=tmasks, lengths=lengths)


if __name__ == "__main__":
    g = generate()
    np.savez_compressed(OUT, **g)
    print(OUT, os.path.getsize(OUT), "bytes; game ends recorded:", int((g["tmasks"] == 0).sum()), "decisions per env:", g["tape_len"].tolist())
