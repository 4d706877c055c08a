"""``Game.randomise_uncertainty`` (game/game.py:1207-1282, the forward-search hook) — the product's game logic
(host emulation of csrc/catan_game.cuh) against the REAL reference under the shared Philox stream: same re-dealt hidden cards,
deck and opponents' hands, same number of draws, and the games keep agreeing when played on from the randomised state."""
import numpy as np
import pytest

from oracle import ref_harness as H
from settlers_of_catan_rl_b200 import layout as L
from tests.common import state_diff
from tests.host_emu.emu_lib import EmuEnv

pytestmark = pytest.mark.skipif(not H.reference_available(), reason="reference tree not present")


def _with_ctr(vec, ctr):
    v = vec.copy()
    st = v.view(L.STATE_DTYPE)[0]
    st["rng_ctr_lo"] = np.int16(np.uint16(ctr & 0xFFFF))
    st["rng_ctr_hi"] = np.int16(np.uint16((ctr >> 16) & 0xFFFF))
    return v


@pytest.mark.parametrize("seed,env_id", [(21, 3), (22, 9)])
def test_randomise_uncertainty_matches_the_reference(seed, env_id):
    R = H.import_reference()
    PlayerId = R["PlayerId"]
    game_rng, samp = H.PhiloxStream(seed, env_id, 0), H.PhiloxStream(seed, env_id, 1)
    env = R["EnvWrapper"]()
    emu = EmuEnv(seed=seed, env_id=env_id, auto_reset=0)
    checked = attempts_seen = skipped = 0
    with H.patched_rng(game_rng):
        obs = env.reset()
        emu.reset()
        for t in range(900):
            if t % 40 == 39 and not env.game.initial_placement_phase:
                c = H.current_actor(env)
                emu.import_state(_with_ctr(H.state_to_vec(env), game_rng.ctr))       # (also aligns the draw counter)
                before = H.state_to_vec(env)
                n = emu.randomise_uncertainty(c, 400)
                if n == 0:
                    # c's beliefs admit no deal that adds up (they are heuristic bounds): the reference would spin in its
                    # `while consistent_distribution_reached == False` loop (game.py:1243) forever; nothing to compare
                    emu.import_state(_with_ctr(before, game_rng.ctr))
                    skipped += 1
                    masks = env.get_action_masks()
                    o_row, m_row = H.obs_to_packed(obs), H.masks_to_packed(masks)
                    a = H.sample_action(m_row, o_row, samp.block(t))
                    obs, _, done, _ = env.step(H.action_to_reference(a))
                    err, _, _ = emu.step(a)
                    assert err == 0
                    if done:
                        break
                    continue
                env.game.randomise_uncertainty(PlayerId(c))
                attempts_seen = max(attempts_seen, n)
                want, got = H.state_to_vec(env), emu.state()
                assert not state_diff(want, got), (t, c, state_diff(want, got)[:5])
                ctr = (int(np.uint16(got.view(L.STATE_DTYPE)[0]["rng_ctr_hi"])) << 16) | int(np.uint16(got.view(L.STATE_DTYPE)[0]["rng_ctr_lo"]))
                assert ctr == game_rng.ctr, "the two consumed a different number of draws"
                # conservation (the reference's closing assert, game.py:1276-1282) and what must not change
                st, b4 = got.view(L.STATE_DTYPE)[0], before.view(L.STATE_DTYPE)[0]
                assert np.array_equal(st["res"].sum(axis=0) + st["bank"], np.full(5, 19))
                assert np.array_equal(st["res"][c - 1], b4["res"][c - 1]) and np.array_equal(st["n_hidden"], b4["n_hidden"])
                assert np.array_equal(st["hidden"][c - 1], b4["hidden"][c - 1]) and st["deck_n"] == b4["deck_n"]
                obs = env._get_obs()
                assert np.array_equal(H.obs_to_packed(obs), emu.obs())
                checked += 1
            masks = env.get_action_masks()
            o_row, m_row = H.obs_to_packed(obs), H.masks_to_packed(masks)
            a = H.sample_action(m_row, o_row, samp.block(t))
            obs, _, done, _ = env.step(H.action_to_reference(a))
            err, _, _ = emu.step(a)
            assert err == 0
            if done:
                break
            assert np.array_equal(H.obs_to_packed(obs), emu.obs()), t
    assert checked >= 8, (checked, skipped)
