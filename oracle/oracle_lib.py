"""TEST INFRASTRUCTURE — ctypes binding of the CPU oracle (``oracle/libcatan_oracle.so``).

Imported only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs.  The product package never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from settlers_of_catan_rl_b200 import layout as L  # noqa: E402

SO = os.path.join(_HERE, "libcatan_oracle.so")


class Config(C.Structure):
    _fields_ = [
        ("max_actions_per_turn", C.c_int32),
        ("max_proposed_trades_per_turn", C.c_int32),
        ("validate_actions", C.c_int32),
        ("dense_reward", C.c_int32),
        ("auto_reset", C.c_int32),
        ("win_reward", C.c_float),
        ("reward_annealing_factor", C.c_float),
    ]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("catan_oracle.c", "catan_oracle.h")] + [
        os.path.join(_HERE, "..", "include", f) for f in ("catan_layout.h", "catan_topology.h")
    ]
    stale = force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcatan_oracle.so"], stdout=subprocess.DEVNULL)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    l = C.CDLL(SO)
    u8p, i32p, f32p, i16p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_int32, C.c_float, C.c_int16, C.c_uint64))
    l.catan_oracle_default_config.argtypes = [C.POINTER(Config)]
    l.catan_oracle_reset.argtypes = [i16p, C.c_uint64, C.c_uint64]
    l.catan_oracle_step.argtypes = [i16p, C.POINTER(Config), i32p, C.c_uint64, C.c_uint64, f32p, u8p]
    l.catan_oracle_step.restype = C.c_int
    l.catan_oracle_actor.argtypes = [i16p]
    l.catan_oracle_actor.restype = C.c_int
    l.catan_oracle_masks.argtypes = [i16p, C.POINTER(Config), u8p]
    l.catan_oracle_obs.argtypes = [i16p, u8p]
    l.catan_oracle_longest_path.argtypes = [i16p, C.c_int]
    l.catan_oracle_longest_path.restype = C.c_int
    l.catan_oracle_sample.argtypes = [u8p, u8p, C.c_uint64, C.c_uint64, C.c_uint64, i32p]
    l.catan_oracle_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
    l.catan_oracle_rollout.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(Config), i16p, u8p,
                                       u8p, u64p, f32p, i32p, u8p, i32p, C.c_int]
    l.catan_oracle_rollout.restype = C.c_int
    l.catan_oracle_gae.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, C.c_double, C.c_double, f32p, f32p]
    l.catan_oracle_state_words.restype = C.c_int
    assert l.catan_oracle_state_words() == L.STATE_WORDS, (l.catan_oracle_state_words(), L.STATE_WORDS)
    _lib = l
    return l


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def make_config(**kw) -> Config:
    c = Config()
    lib().catan_oracle_default_config(C.byref(c))
    for k, v in kw.items():
        if not hasattr(c, k):
            raise KeyError(k)
        setattr(c, k, v)
    return c


class OracleEnv:
    """One game on the C oracle, EnvWrapper-shaped but on packed arrays."""

    def __init__(self, seed: int = 0, env_id: int = 0, **cfg):
        self.l = lib()
        self.cfg = make_config(**cfg)
        self.seed, self.env_id = seed, env_id
        self.state = np.zeros(L.STATE_WORDS, dtype=np.int16)

    def reset(self):
        self.l.catan_oracle_reset(_p(self.state, C.c_int16), self.seed, self.env_id)
        return self.obs()

    def obs(self):
        o = np.zeros(L.OBS_STRIDE, dtype=np.uint8)
        self.l.catan_oracle_obs(_p(self.state, C.c_int16), _p(o, C.c_uint8))
        return o

    def masks(self):
        m = np.zeros(L.MASK_STRIDE, dtype=np.uint8)
        self.l.catan_oracle_masks(_p(self.state, C.c_int16), C.byref(self.cfg), _p(m, C.c_uint8))
        return m

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.int32)
        r = np.zeros(4, dtype=np.float32)
        info = np.zeros(L.INFO_STRIDE, dtype=np.uint8)
        err = self.l.catan_oracle_step(_p(self.state, C.c_int16), C.byref(self.cfg), _p(a, C.c_int32), self.seed,
                                       self.env_id, _p(r, C.c_float), _p(info, C.c_uint8))
        return err, r, info

    def sample(self, masks, obs, decision):
        a = np.zeros(L.ACTION_WORDS, dtype=np.int32)
        self.l.catan_oracle_sample(_p(masks, C.c_uint8), _p(obs, C.c_uint8), self.seed, self.env_id, decision,
                                   _p(a, C.c_int32))
        return a

    def structured(self):
        return self.state.view(L.STATE_DTYPE)[0]


class OracleVec:
    """catan_oracle_rollout driver: n_envs games advanced in lock-step blocks on host threads."""

    def __init__(self, n_envs: int, seed: int = 0, first_env_id: int = 0, n_threads: int = 0, **cfg):
        self.l = lib()
        cfg.setdefault("auto_reset", 1)
        self.cfg = make_config(**cfg)
        self.n, self.seed, self.first, self.n_threads = n_envs, seed, first_env_id, n_threads
        self.states = np.zeros((n_envs, L.STATE_WORDS), dtype=np.int16)
        self.obs = np.zeros((n_envs, L.OBS_STRIDE), dtype=np.uint8)
        self.masks = np.zeros((n_envs, L.MASK_STRIDE), dtype=np.uint8)
        self.decisions = np.zeros(n_envs, dtype=np.uint64)
        self.reward_sum = np.zeros((n_envs, 4), dtype=np.float32)
        self.games_done = np.zeros(n_envs, dtype=np.int32)
        self.info = np.zeros((n_envs, L.INFO_STRIDE), dtype=np.uint8)
        self.actions = np.zeros((n_envs, L.ACTION_WORDS), dtype=np.int32)
        self.fresh = True
        self.threads_used = 0

    def run(self, n_steps: int) -> int:
        self.threads_used = self.l.catan_oracle_rollout(
            self.n, self.seed, self.first, n_steps, int(self.fresh), C.byref(self.cfg), _p(self.states, C.c_int16),
            _p(self.obs, C.c_uint8), _p(self.masks, C.c_uint8), _p(self.decisions, C.c_uint64),
            _p(self.reward_sum, C.c_float), _p(self.games_done, C.c_int32), _p(self.info, C.c_uint8),
            _p(self.actions, C.c_int32), self.n_threads)
        self.fresh = False
        return self.threads_used


def gae(rewards, values, masks, gamma, lam):
    """C restatement of process_batch.py:134-140 -> (returns, advantages)."""
    T, N = rewards.shape[0], rewards.shape[1]
    r = np.ascontiguousarray(rewards, dtype=np.float32).reshape(T, N)
    v = np.ascontiguousarray(values, dtype=np.float32).reshape(T + 1, N)
    m = np.ascontiguousarray(masks, dtype=np.float32).reshape(T + 1, N)
    ret = np.zeros((T, N), dtype=np.float32)
    adv = np.zeros((T, N), dtype=np.float32)
    lib().catan_oracle_gae(_p(r, C.c_float), _p(v, C.c_float), _p(m, C.c_float), T, N, gamma, lam, _p(ret, C.c_float),
                           _p(adv, C.c_float))
    return ret, adv


def philox(c0, c1, c2, c3, k0, k1):
    out = (C.c_uint32 * 4)()
    lib().catan_oracle_philox(c0, c1, c2, c3, k0, k1, out)
    return tuple(out)
