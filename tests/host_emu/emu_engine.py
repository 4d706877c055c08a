"""TEST INFRASTRUCTURE — the single-env engine interface ``settlers_of_catan_rl_b200.EnvWrapper`` talks to
(``reset_host / step_host / export_state / import_state / rows_host / set_reward_annealing_factor / config``), backed by
the host emulation of the PRODUCT game logic (``tests/host_emu/emu.cpp`` = ``csrc/catan_game.cuh`` compiled by g++).
It lets the reference's unchanged managers run over the adapter in the CPU-only container (``EnvWrapper(_engine=EmuEngine)``).
Never part of the package; the product's engine is the CUDA library."""
from __future__ import annotations

import ctypes as C

import numpy as np

from settlers_of_catan_rl_b200 import layout as L
from tests.host_emu.emu_lib import EmuEnv


class EmuEngine:
    def __init__(self, seed=0, env_id=0, **cfg):
        self.e = EmuEnv(seed=seed, env_id=env_id, **cfg)
        self.config = self.e.cfg

    def reset_host(self, obs, masks, info):
        self.e.reset()
        obs[0], masks[0] = self.e.obs(), self.e.masks()
        if info is not None:
            info[0] = 0
            info[0, L.INFO_ACTOR] = self._actor()

    def _actor(self):
        st = self.e.state().view(L.STATE_DTYPE)[0]
        if st["need_discard"]:
            return int(st["discard_queue"][0])
        if st["must_respond"]:
            return int(st["trade_target"])
        return int(st["players_go"])

    def step_host(self, actions, obs, masks, reward, info):
        err, r, inf = self.e.step(actions[0])
        obs[0], masks[0], reward[0], info[0] = self.e.obs(), self.e.masks(), r, inf

    def export_state(self):
        return self.e.state()[None, :]

    def import_state(self, states):
        self.e.import_state(np.asarray(states).reshape(-1, L.STATE_WORDS)[0])

    def rows_host(self):
        return self.e.obs()[None, :], self.e.masks()[None, :]

    def randomise_uncertainty(self, pid, max_attempts=10000):
        self.e.randomise_uncertainty(int(pid), int(max_attempts))

    def set_reward_annealing_factor(self, factor):
        self.config.reward_annealing_factor = float(factor)
        self.e.l.emu_set_config(self.e.h, C.byref(self.config))
