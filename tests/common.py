"""Shared helpers for the parity tests."""
from __future__ import annotations

import glob
import os

import numpy as np

from settlers_of_catan_rl_b200 import layout as L

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    """the env trajectory fixtures (oracle/make_golden.py); rollout_*.npz are collector tapes (oracle/make_rollout_golden.py)"""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if not n.startswith("rollout_")]


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    g["cfg"] = {str(k): int(v) for k, v in zip(g["cfg_keys"], g["cfg_vals"])}
    return g


def state_diff(a, b, ignore=L.STATE_NON_REFERENCE_FIELDS):
    """field-wise differences between two canonical int16 state vectors"""
    sa = np.ascontiguousarray(a, dtype=np.int16).view(L.STATE_DTYPE)[0]
    sb = np.ascontiguousarray(b, dtype=np.int16).view(L.STATE_DTYPE)[0]
    return [(n, sa[n].tolist(), sb[n].tolist()) for n in L.STATE_DTYPE.names if n not in ignore and not np.array_equal(sa[n], sb[n])]


def row_diff(a, b, limit=12):
    w = np.nonzero(np.asarray(a) != np.asarray(b))[0]
    return [(int(i), int(a[i]), int(b[i])) for i in w[:limit]]


def replay_golden(env, g, check_every_state=True):
    """Replay a golden trajectory on `env`, an object with reset()/step(a)/state()/obs()/masks() returning
    numpy rows (EnvWrapper.step semantics: no auto-reset).  Raises AssertionError with the first mismatch."""
    env.reset()
    resets = list(g["reset_at"])
    ri = 0

    def check(tag, t, st, ob, mk):
        assert not state_diff(st, env.state()), (tag, t, state_diff(st, env.state())[:6])
        assert np.array_equal(ob, env.obs()), (tag, t, "obs", row_diff(ob, env.obs()))
        assert np.array_equal(mk, env.masks()), (tag, t, "masks", row_diff(mk, env.masks()))

    check("reset", -1, g["state"][0], g["obs"][0], g["masks"][0])
    for t in range(len(g["actions"])):
        err, reward, info = env.step(g["actions"][t])
        assert err == 0, ("step rejected", t, err, g["actions"][t])
        assert np.array_equal(reward, g["reward"][t]), ("reward", t, reward, g["reward"][t])
        assert int(info[L.INFO_DONE]) == int(g["done"][t]), ("done", t)
        check("step", t, g["state"][t + 1], g["obs"][t + 1], g["masks"][t + 1])
        if ri < len(resets) and resets[ri] == t + 1:
            env.reset()
            check("re-reset", t, g["reset_state"][ri], g["reset_obs"][ri], g["reset_masks"][ri])
            ri += 1
    return len(g["actions"])
