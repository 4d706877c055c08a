"""The EnvWrapper-shaped adapter (reference surface, env/wrapper.py:11-50) on the CUDA engine: formats, attributes the
reference's managers use, and bit-parity of the decoded observation with the golden fixtures."""
import numpy as np
import pytest

from tests.common import load_golden
from settlers_of_catan_rl_b200 import layout as L

pytestmark = pytest.mark.gpu


def _unpack_action(a):
    out = [np.array(int(a[i])) for i in range(7)]
    out.append([np.array(int(a[L.A_GIVE + k])) for k in range(4)])
    out.append([np.array(int(a[L.A_RECV + k])) for k in range(4)])
    out += [np.array(int(a[L.A_RES_A])), np.array(int(a[L.A_RES_B])), np.array(int(a[L.A_DISCARD]))]
    return out


def _players_turn(env):
    """verbatim logic of RL/ppo/game_manager.py:152-159"""
    if env.game.players_need_to_discard:
        return env.game.players_to_discard[0]
    if env.game.must_respond_to_trade:
        return env.game.proposed_trade["target_player"]
    return env.game.players_go


def test_envwrapper_adapter_matches_reference_formats_and_values():
    from settlers_of_catan_rl_b200 import EnvWrapper
    from settlers_of_catan_rl_b200.enums import PlayerId
    g = load_golden("default_s0")
    env = EnvWrapper(seed=int(g["seed"]), env_id=int(g["env_id"]))
    obs = env.reset()
    assert set(obs.keys()) == {"proposed_trade", "current_resources", "player_id", "tile_representations", "current_player_main",
                               "current_player_played_dev", "current_player_hidden_dev", "next_player_main",
                               "next_player_played_dev", "next_next_player_main", "next_next_player_played_dev",
                               "next_next_next_player_main", "next_next_next_player_played_dev"}
    assert isinstance(obs["tile_representations"], list) and len(obs["tile_representations"]) == 19
    assert obs["tile_representations"][0].shape == (60,) and obs["current_player_main"].shape == (152,)
    assert obs["next_player_main"].shape == (159,) and isinstance(obs["player_id"], PlayerId)
    masks = env.get_action_masks()
    assert [m.shape for m in masks] == [(13,), (3, 54), (73,), (19,), (5,), (2,), (3, 3), (6,), (6,), (4, 5), (5,), (5,)]
    ratio = dict(L.OBS_RATIO_COLUMNS)

    def flat(o):
        row = np.concatenate([np.asarray(o[k], dtype=np.float64).reshape(-1) for k, _, _ in L.OBS_NUMERIC])
        return row

    for t in range(len(g["actions"])):
        want = g["obs"][t][:L.OBS_FEATURES].astype(np.float64)
        for col, div in ratio.items():
            want[col] /= div
        assert np.array_equal(flat(obs), want), t
        assert int(obs["player_id"]) == int(g["obs"][t][L.OBS_META]) == int(_players_turn(env))
        mk = np.concatenate([m.reshape(-1) for m in env.get_action_masks()])
        assert np.array_equal(mk, g["masks"][t][:L.MASK_ENTRIES].astype(np.float64))
        obs, reward, done, info = env.step(_unpack_action(g["actions"][t]))
        assert [reward[PlayerId(p + 1)] for p in range(4)] == [float(x) for x in g["reward"][t]]
        assert done == bool(g["done"][t]) and "log" in info
        if t % 97 == 0:                                    # save_state / restore_state round trip
            st = env.save_state()
            env.restore_state(st)
            assert np.array_equal(env.save_state()["state"], st["state"])
    assert done and env.winner is not None and env.curr_vps[env.winner.id] >= 10
    played = obs["current_player_played_dev"]
    assert played.dtype.kind == "i" and len(played) >= 1


def test_envwrapper_raises_like_the_reference_on_invalid_actions():
    from settlers_of_catan_rl_b200 import EnvWrapper
    env = EnvWrapper(seed=5, env_id=5)
    env.reset()
    bad = [np.array(9)] + [np.array(0)] * 6 + [[np.array(0)] * 4, [np.array(0)] * 4] + [np.array(0)] * 3   # RollDice in the initial phase
    with pytest.raises(RuntimeError):
        env.step(bad)
    env.reward_annealing_factor = 0.25
    assert env.reward_annealing_factor == 0.25
