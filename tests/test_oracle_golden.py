"""The C oracle against the golden fixtures generated from the real reference (CPU)."""
import numpy as np
import pytest

from oracle import oracle_lib as O
from tests.common import golden_cases, load_golden, replay_golden
from settlers_of_catan_rl_b200 import layout as L


class _OracleAdapter:
    def __init__(self, g):
        self.e = O.OracleEnv(int(g["seed"]), int(g["env_id"]), **g["cfg"])

    def reset(self):
        self.e.reset()

    def step(self, a):
        return self.e.step(a)

    def state(self):
        return self.e.state

    def obs(self):
        return self.e.obs()

    def masks(self):
        return self.e.masks()


@pytest.mark.parametrize("case", golden_cases())
def test_oracle_replays_golden(case):
    g = load_golden(case)
    assert replay_golden(_OracleAdapter(g), g) == len(g["actions"])


@pytest.mark.parametrize("case", golden_cases()[:3])
def test_oracle_sampler_reproduces_golden_actions(case):
    """the pinned random-legal sampler (C twin) regenerates the recorded actions from masks + obs"""
    g = load_golden(case)
    e = O.OracleEnv(int(g["seed"]), int(g["env_id"]))
    resets = {int(t): i for i, t in enumerate(g["reset_at"])}
    for t in range(len(g["actions"])):
        m, o = (g["reset_masks"][resets[t]], g["reset_obs"][resets[t]]) if t in resets else (g["masks"][t], g["obs"][t])
        assert np.array_equal(e.sample(m, o, t), g["actions"][t]), t


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors): zero counter/key and the pi digits one."""
    assert O.philox(0, 0, 0, 0, 0, 0) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert O.philox(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF) == (
        0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert O.philox(0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344, 0xA4093822, 0x299F31D0) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_resource_conservation_and_vp_invariants():
    """game/utils.py:18-26 with the correct total (19 per resource) + the VP identity of SURVEY.md §4."""
    v = O.OracleVec(64, seed=123, first_env_id=0)
    for _ in range(40):
        v.run(50)
        st = v.states.view(L.STATE_DTYPE).reshape(-1)
        assert np.all(st["bank"] + st["res"].sum(axis=1) == 19)
        settlements = np.stack([((st["corner_type"] == 1) & (st["corner_owner"] == p + 1)).sum(axis=1) for p in range(4)], 1)
        cities = np.stack([((st["corner_type"] == 2) & (st["corner_owner"] == p + 1)).sum(axis=1) for p in range(4)], 1)
        played_vp = np.stack([((st["played"][:, p, :] == 1) & (np.arange(25)[None, :] < st["n_played"][:, p, None])).sum(axis=1)
                              for p in range(4)], 1)
        lr = np.stack([(st["lr_holder"] == p + 1) for p in range(4)], 1)
        la = np.stack([(st["la_holder"] == p + 1) for p in range(4)], 1)
        assert np.array_equal(st["vp"], settlements + 2 * cities + played_vp + 2 * lr + 2 * la)
        assert np.array_equal(st["settlements_left"], 5 - settlements)
        assert np.array_equal(st["cities_left"], 4 - cities)
    assert v.games_done.sum() > 0
