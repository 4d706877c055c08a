"""The PRODUCT game core (catan_core.cuh) compiled for the host with one lane per warp, against the golden
fixtures and against the oracle on long random games (CPU).  Catches logic bugs before any GPU time."""
import numpy as np
import pytest

from oracle import oracle_lib as O
from tests.common import golden_cases, load_golden, replay_golden, state_diff, row_diff
from tests.host_emu.emu_lib import EmuEnv


class _EmuAdapter:
    def __init__(self, g):
        self.e = EmuEnv(int(g["seed"]), int(g["env_id"]), **g["cfg"])

    def reset(self):
        self.e.reset()

    def step(self, a):
        return self.e.step(a)

    def state(self):
        return self.e.state()

    def obs(self):
        return self.e.obs()

    def masks(self):
        return self.e.masks()


@pytest.mark.parametrize("case", golden_cases())
def test_core_replays_golden(case):
    g = load_golden(case)
    assert replay_golden(_EmuAdapter(g), g) == len(g["actions"])


@pytest.mark.parametrize("seed,cfg", [
    (21, {}),
    (22, dict(dense_reward=1, reward_annealing_factor=0.37)),
    (23, dict(max_proposed_trades_per_turn=-1)),
    (24, dict(max_actions_per_turn=3, max_proposed_trades_per_turn=1)),
    (25, dict(auto_reset=1)),
])
def test_core_matches_oracle_on_random_games(seed, cfg):
    o = O.OracleEnv(seed, 500 + seed, **{k: v for k, v in cfg.items() if k != "auto_reset"})
    e = EmuEnv(seed, 500 + seed, **cfg)
    o.reset()
    e.reset()
    games = 0
    for t in range(6000):
        om, oo = o.masks(), o.obs()
        assert not state_diff(o.state, e.state(), ignore=()), (t, state_diff(o.state, e.state(), ignore=())[:5])
        assert np.array_equal(om, e.masks()), (t, row_diff(om, e.masks()))
        assert np.array_equal(oo, e.obs()), (t, row_diff(oo, e.obs()))
        a = o.sample(om, oo, t)
        assert np.array_equal(a, e.sample()), t
        err, r, info = o.step(a)
        err2, r2, info2 = e.step(a)
        assert err == 0 and err2 == 0
        assert np.array_equal(r, r2)
        if info[0]:
            games += 1
            o.reset()
            if cfg.get("auto_reset"):
                assert info2[11] == 1
                info[11] = 1
                info[6] = info2[6]
            else:
                e.reset()
        assert np.array_equal(info, info2), (t, info, info2)
    assert games >= 1


@pytest.mark.parametrize("seed", [31, 32, 33])
def test_search_through_the_new_road_matches_oracle(seed):
    """every road placement settled by the pooled search of the paths through the new road (the device's lr_slow_kernel
    path) instead of the one-thread incremental rule: states (cur_longest_path, holder, VP) stay bit-equal to the oracle"""
    o = O.OracleEnv(seed, 700 + seed)
    e = EmuEnv(seed, 700 + seed, flavor="search")
    o.reset()
    e.reset()
    games = 0
    for t in range(5000):
        om, oo = o.masks(), o.obs()
        assert not state_diff(o.state, e.state(), ignore=()), (t, state_diff(o.state, e.state(), ignore=())[:5])
        a = o.sample(om, oo, t)
        err, r, info = o.step(a)
        err2, r2, info2 = e.step(a)
        assert err == 0 and err2 == 0 and np.array_equal(r, r2)
        if info[0]:
            games += 1
            o.reset()
            e.reset()
    assert games >= 1


def test_core_rejects_illegal_actions_like_the_reference():
    """wrapper.py:38-41: an invalid action raises; here: error code, state untouched."""
    e = EmuEnv(1, 1)
    o = O.OracleEnv(1, 1)
    e.reset()
    o.reset()
    before = e.state().copy()
    for a0 in ([9] + [0] * 19, [10] + [0] * 19, [1, 0, 72] + [0] * 17, [13] + [0] * 19, [0, 54] + [0] * 18, [12] + [0] * 19):
        err, _, info = e.step(np.array(a0, dtype=np.int32))
        err_o, _, _ = o.step(np.array(a0, dtype=np.int32))
        assert err != 0 and err == err_o and info[10] == err
        assert np.array_equal(before, e.state())


def test_longest_path_matches_oracle_on_dense_road_networks():
    rng = np.random.default_rng(0)
    from settlers_of_catan_rl_b200 import layout as L
    e = EmuEnv(3, 3)
    e.reset()
    o = O.OracleEnv(3, 3)
    import ctypes as C
    for trial in range(200):
        st = e.state().view(L.STATE_DTYPE)[0].copy()
        st["edge_owner"][:] = rng.choice([0, 1, 2], size=72, p=[0.3, 0.5, 0.2])
        st["corner_type"][:] = 0
        st["corner_owner"][:] = 0
        for c in rng.choice(54, size=6, replace=False):
            st["corner_type"][c] = 1
            st["corner_owner"][c] = rng.integers(1, 3)
        vec = np.frombuffer(st.tobytes(), dtype=np.int16).copy()
        e.import_state(vec)
        o.state[:] = vec
        for pid in (1, 2):
            want = o.l.catan_oracle_longest_path(o.state.ctypes.data_as(C.POINTER(C.c_int16)), pid)
            assert e.longest_path(pid) == want
