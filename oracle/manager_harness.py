"""TEST INFRASTRUCTURE — drives the reference's UNCHANGED rollout collector ``GamesAndPoliciesManager``
(RL/ppo/game_manager.py:11-166) with deterministic stub policies, and records everything that happens at its env
boundary as an event tape.

Used to pin ``oracle/rollout_ref.py`` (the restatement the GPU test of ``catan_rollout_store`` checks against), to generate
the committed tape ``tests/golden/rollout_manager_tape.npz`` (``oracle/make_rollout_golden.py``) and for the drop-in test
(the same manager over this repo's ``EnvWrapper`` adapter).  Never on the product path.

How the manager is made deterministic without touching its code:
  * its four policies keep their real ``obs_to_torch / act_masks_to_torch / torch_act_to_np`` (RL/models/policy.py:168-199);
    only ``act`` is replaced by the pinned random-legal sampler (``ref_harness.sample_action``, one Philox block per
    decision of the env's sampler stream), returning the reference's output format (12 heads of ``[1, 1]`` long tensors,
    heads 7 / 8 lists of four) and a log-prob that is unique per decision;
  * every env is wrapped in ``RecordingEnv``: it forwards ``reset / step / get_action_masks / game / reward_annealing_factor``
    to the real env, routes the global RNG entry points to THAT env's game stream for the duration of the call
    (``ref_harness.patched_rng``) and logs the call.  Observations, mask lists and action lists carry a serial number, so the
    manager's output lists can be compared entry by entry with what ``RefCollector`` (or the CUDA collector) produces.
"""
from __future__ import annotations

import copy
import random as _py_random
import sys
import types

import numpy as np

from settlers_of_catan_rl_b200 import layout as L
from oracle import ref_harness as H


def import_manager():
    """the reference's ``RL.ppo.game_manager`` module (headless import recipe of SURVEY 8c)"""
    H.import_reference()
    import RL.ppo.game_manager as gm  # type: ignore
    return gm


class TaggedList(list):
    """a list that can carry a serial number through the manager's in-place conversions"""
    serial = -1


class Tape:
    """what happened at the env boundary of ONE env, in order"""

    def __init__(self):
        self.events = []          # ("reset", obs_serial, actor) | ("masks", mask_serial) | ("step", {...})
        self.obs_rows = {}        # serial -> packed uint8 obs row
        self.mask_rows = {}       # serial -> packed uint8 mask row
        self.actions = []         # int32[20] per decision
        self.logps = []           # float per decision


class _Ctx:
    current = None                # the RecordingEnv whose decision is being taken (set by get_action_masks)
    serial = 0

    @classmethod
    def next_serial(cls):
        cls.serial += 1
        return cls.serial


class RecordingEnv:
    def __init__(self, env, seed: int, env_id: int, philox: bool = True):
        object.__setattr__(self, "_env", env)
        self._game_rng = H.PhiloxStream(seed, env_id, 0) if philox else None
        self._samp = H.PhiloxStream(seed, env_id, 1)
        self._decision = 0
        self.tape = Tape()
        self._last_obs_row = None
        self._last_mask_row = None

    # ---- everything the managers reach through the env
    @property
    def game(self):
        return self._env.game

    @property
    def reward_annealing_factor(self):
        return self._env.reward_annealing_factor

    @reward_annealing_factor.setter
    def reward_annealing_factor(self, v):
        self._env.reward_annealing_factor = v

    @property
    def winner(self):
        return self._env.winner

    @property
    def curr_vps(self):
        return self._env.curr_vps

    def _rng(self):
        return H.patched_rng(self._game_rng) if self._game_rng is not None else _Null()

    def _tag_obs(self, obs):
        s = _Ctx.next_serial()
        obs["__serial"] = s
        row = H.obs_to_packed(obs)
        self.tape.obs_rows[s] = row
        self._last_obs_row = row
        return s

    def reset(self):
        with self._rng():
            obs = self._env.reset()
        s = self._tag_obs(obs)
        self.tape.events.append(("reset", s, int(H.current_actor(self._env))))
        return obs

    def get_action_masks(self):
        masks = TaggedList(self._env.get_action_masks())
        masks.serial = _Ctx.next_serial()
        row = H.masks_to_packed(masks)
        self.tape.mask_rows[masks.serial] = row
        self._last_mask_row = row
        self.tape.events.append(("masks", masks.serial))
        _Ctx.current = self
        return masks

    def step(self, action):
        actor = int(H.current_actor(self._env))
        with self._rng():
            obs, reward, done, info = self._env.step(action)
        s = self._tag_obs(obs)
        PlayerId = H.import_reference()["PlayerId"]
        self.tape.events.append(("step", dict(
            actor=actor, action_serial=getattr(action, "serial", -1), obs_serial=s, done=bool(done),
            reward=[float(reward[PlayerId(p + 1)]) for p in range(4)], n_actor=int(H.current_actor(self._env)))))
        return obs, reward, done, info

    # ---- the stub policy's decision for this env
    def decide(self):
        a = H.sample_action(self._last_mask_row, self._last_obs_row, self._samp.block(self._decision))
        self._decision += 1
        logp = -float(len(self.tape.actions) + 1) / 1024.0          # unique and exactly representable in fp32
        self.tape.actions.append(a.copy())
        self.tape.logps.append(logp)
        return a, logp


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def stub_act(obs_dict, hidden_states, terminal_mask, action_masks, **kw):
    """``SettlersAgentPolicy.act`` (policy.py:71-92) replaced by the pinned random-legal sampler; same return format"""
    import torch
    env = _Ctx.current
    a, logp = env.decide()
    heads = TaggedList()
    for h in range(7):
        heads.append(torch.tensor([[int(a[h])]], dtype=torch.long))
    heads.append([torch.tensor([[int(a[L.A_GIVE + k])]], dtype=torch.long) for k in range(4)])
    heads.append([torch.tensor([[int(a[L.A_RECV + k])]], dtype=torch.long) for k in range(4)])
    for col in (L.A_RES_A, L.A_RES_B, L.A_DISCARD):
        heads.append(torch.tensor([[int(a[col])]], dtype=torch.long))
    heads.serial = len(env.tape.actions) - 1                         # index into tape.actions
    return None, heads, torch.tensor([[logp]], dtype=torch.float32), hidden_states


def make_manager(num_envs: int, num_steps: int, seed: int, first_env_id: int, env_factory=None, env_kwargs=None, shuffle_seed: int = 0,
                 philox: bool = True):
    """an unchanged ``GamesAndPoliciesManager`` whose envs are ``RecordingEnv`` proxies (over the reference's EnvWrapper, or
    over whatever ``env_factory(env_id)`` returns) and whose policies act through ``stub_act``"""
    gm = import_manager()
    RefEnvWrapper = H.import_reference()["EnvWrapper"]
    ids = iter(range(first_env_id, first_env_id + num_envs))

    def env_ctor():                                                          # what the manager's `EnvWrapper()` call returns
        env_id = next(ids)
        inner = env_factory(env_id) if env_factory is not None else RefEnvWrapper(**(env_kwargs or {}))
        return RecordingEnv(inner, seed, env_id, philox=philox and env_factory is None)

    saved = gm.EnvWrapper
    gm.EnvWrapper = env_ctor
    try:
        _py_random.seed(shuffle_seed)                                        # initialise() shuffles the seat order with `random`
        mgr = gm.GamesAndPoliciesManager(num_envs=num_envs, num_steps=num_steps)   # the reference's constructor, unchanged
    finally:
        gm.EnvWrapper = saved
    for p in mgr.policies:
        p.act = stub_act
    return mgr


def rollout_lists(result):
    """``gather_rollouts()`` output -> per env: serials / values of every list (for comparisons)"""
    observations, _hidden, rewards, actions, action_masks, action_log_probs, terminal_masks = result
    out = []
    for e in range(len(observations)):
        out.append(dict(
            obs=[int(o["__serial"]) for o in observations[e]],
            masks=[int(m.serial) for m in action_masks[e]],
            actions=[int(a.serial) for a in actions[e]],
            logp=[float(x.reshape(-1)[0]) for x in action_log_probs[e]],
            rewards=[float(r) for r in rewards[e]],
            tmasks=[float(t) for t in terminal_masks[e]],
        ))
    return out


def replay_tape_through(collector, tape: Tape, start_event: int = 0, fresh: bool = True):
    """feed the recorded env-boundary events of one env to a ``RefCollector``-shaped object, one manager loop iteration
    (masks -> step [-> reset]) at a time, for as long as it is collecting.  Returns the index of the next unread event."""
    ev = tape.events
    i = start_event
    if fresh:
        kind, s, actor = ev[i]
        assert kind == "reset"
        collector.reset(actor, s)
        i += 1
    else:
        collector.after_rollouts()
    while collector.collecting():
        assert ev[i][0] == "masks", ev[i]
        mask_serial = ev[i][1]
        st = ev[i + 1][1]
        i += 2
        obs_after_reset, actor_after_reset = st["obs_serial"], st["n_actor"]
        if st["done"]:
            kind, s, actor = ev[i]
            assert kind == "reset"
            obs_after_reset, actor_after_reset = s, actor
            i += 1
        collector.tick(st["actor"], mask_serial, st["action_serial"], tape.logps[st["action_serial"]], st["reward"], st["done"],
                       st["n_actor"], actor_after_reset, st["obs_serial"], obs_after_reset)
    return i
