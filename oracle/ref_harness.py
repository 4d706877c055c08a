"""TEST INFRASTRUCTURE — drives the *real* reference (``/root/reference``) to pin the oracle.

Only ``tests/`` and the fixture generator (``oracle/make_golden.py``) import this module.  It is
never on the product path.  It needs the reference tree: ``/root/reference`` in this container, or the
git-ignored copy ``oracle/_ref/`` that ``oracle/build_ref.py`` makes and that travels to the GPU box; the GPU
parity tests additionally replay the committed fixtures under ``tests/golden/``.

What it does:
  * imports ``env.wrapper.EnvWrapper`` headless by stubbing ``ui.display`` / pygame / tkinter
    (needed because of ``game/game.py:13``);
  * "shared-Philox mode": replaces the reference's five RNG entry points
    (``np.random.shuffle`` board.py:72,79,81,84 game.py:42,77; ``np.random.randint`` game.py:139-140;
    ``random.choice`` game.py:643) with the build's pinned Philox4x32-10 stream
    (``include/catan_layout.h``), so the reference, the C oracle and the CUDA kernels generate the
    same games from ``(seed, env_id)``;
  * converts the reference's ``save_current_state()`` / obs dict / mask list into the canonical
    int16 state vector and the packed uint8 obs / mask rows;
  * records full trajectories under the pinned random-legal sampler.
"""
from __future__ import annotations

import os
import random as _py_random
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from settlers_of_catan_rl_b200 import layout as L  # noqa: E402

from oracle import build_ref as _build_ref  # noqa: E402

#: the mounted reference where it exists (this container), else the copy oracle/build_ref.py made (the GPU box)
REFERENCE_ROOT = os.environ.get("CATAN_REFERENCE_ROOT") or _build_ref.root() or "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "game"))


_ref = {}


def import_reference():
    """Import the reference's EnvWrapper headless (SURVEY.md 8c import recipe)."""
    if _ref:
        return _ref
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for n in ("pygame", "tkinter"):
        sys.modules.setdefault(n, types.ModuleType(n))
    d = types.ModuleType("ui.display")
    d.Display = object
    u = types.ModuleType("ui")
    u.display = d
    sys.modules["ui"] = u
    sys.modules["ui.display"] = d
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from env.wrapper import EnvWrapper  # type: ignore
    from game.enums import PlayerId, Resource, DevelopmentCard, BuildingType, ActionTypes  # type: ignore

    _ref.update(
        EnvWrapper=EnvWrapper,
        PlayerId=PlayerId,
        Resource=Resource,
        DevelopmentCard=DevelopmentCard,
        BuildingType=BuildingType,
        ActionTypes=ActionTypes,
    )
    return _ref


# ---------------------------------------------------------------------------------------------
# Pinned Philox4x32-10 (pure Python ints; the C twin is oracle/catan_oracle.c:philox4x32)
# ---------------------------------------------------------------------------------------------
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32(counter, key):
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> 32, p0 & _MASK
        hi1, lo1 = p1 >> 32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


class PhiloxStream:
    """draw d -> word (d & 3) of philox(counter=(d >> 2, stream, env_lo, env_hi), key=(seed_lo, seed_hi))."""

    def __init__(self, seed: int, env_id: int, stream: int, ctr: int = 0):
        self.key = (seed & _MASK, (seed >> 32) & _MASK)
        self.env = (env_id & _MASK, (env_id >> 32) & _MASK)
        self.stream = stream
        self.ctr = ctr

    def block(self, idx: int):
        return philox4x32((idx & _MASK, self.stream, self.env[0], self.env[1]), self.key)

    def next_u32(self) -> int:
        d = self.ctr
        self.ctr += 1
        return self.block(d >> 2)[d & 3]

    def bounded(self, n: int) -> int:
        return (self.next_u32() * n) >> 32

    def shuffle(self, a) -> None:
        for i in range(len(a) - 1, 0, -1):
            j = self.bounded(i + 1)
            a[i], a[j] = a[j], a[i]


class patched_rng:
    """Context manager: route the reference's RNG entry points to a PhiloxStream (game stream)."""

    def __init__(self, stream: PhiloxStream):
        self.stream = stream

    def __enter__(self):
        self._saved = (np.random.shuffle, np.random.randint, _py_random.choice, _py_random.shuffle)
        s = self.stream

        def _shuffle(x):
            s.shuffle(x)

        def _randint(low, high=None, size=None):
            assert size is None and high is not None
            return low + s.bounded(high - low)

        def _choice(seq):
            return seq[s.bounded(len(seq))]

        np.random.shuffle = _shuffle
        np.random.randint = _randint
        _py_random.choice = _choice
        _py_random.shuffle = _shuffle            # game.py:1250,1255 (randomise_uncertainty): the same Fisher-Yates
        return self

    def __exit__(self, *exc):
        np.random.shuffle, np.random.randint, _py_random.choice, _py_random.shuffle = self._saved
        return False


# ---------------------------------------------------------------------------------------------
# reference objects -> canonical arrays
# ---------------------------------------------------------------------------------------------
def state_to_vec(env) -> np.ndarray:
    """``EnvWrapper.save_state()`` content (game.py:1013-1091, wrapper.py:711-716) -> int16[STATE_WORDS]."""
    R = import_reference()
    PlayerId, Resource = R["PlayerId"], R["Resource"]
    g = env.game
    st = np.zeros((), dtype=L.STATE_DTYPE)
    for i, t in enumerate(g.board.tiles):
        st["tile_res"][i] = int(t.resource)
        assert int(t.resource) == int(t.terrain)
        st["tile_val"][i] = int(t.value)
        if t.contains_robber:
            st["robber_tile"] = i
    assert g.board.robber_tile.id == int(st["robber_tile"])
    for i, c in enumerate(g.board.corners):
        if c.building is not None:
            st["corner_type"][i] = int(c.building.type) + 1
            st["corner_owner"][i] = int(c.building.owner)
    for i, e in enumerate(g.board.edges):
        if e.road is not None:
            st["edge_owner"][i] = int(e.road)
    st["harbour_perm"][:] = [h.id for h in g.board.harbours]
    st["player_order"][:] = [int(p) for p in g.player_order]
    st["player_order_id"] = g.player_order_id
    st["players_go"] = int(g.players_go)
    labels = ["next", "next_next", "next_next_next"]
    for pid in PlayerId:
        p = int(pid) - 1
        pl = g.players[pid]
        for res in Resource:
            if res == Resource.Empty:
                continue
            r = int(res) - 1
            st["res"][p, r] = pl.resources[res]
            st["vis"][p, r] = pl.visible_resources[res]
            for li, lab in enumerate(labels):
                st["est_min"][p, li, r] = pl.opponent_min_res[lab][res]
                st["est_max"][p, li, r] = pl.opponent_max_res[lab][res]
        st["vp"][p] = pl.victory_points
        hb = 0
        for key in pl.harbours.keys():
            hb |= 1 if key is None else (1 << int(key))
        st["harbours"][p] = hb
        st["n_hidden"][p] = len(pl.hidden_cards)
        st["hidden"][p, : len(pl.hidden_cards)] = [int(c) for c in pl.hidden_cards]
        st["n_played"][p] = len(pl.visible_cards)
        st["played"][p, : len(pl.visible_cards)] = [int(c) for c in pl.visible_cards]
        st["settlements_left"][p] = g.building_bank["settlements"][pid]
        st["cities_left"][p] = g.building_bank["cities"][pid]
        st["init_settlements"][p] = g.initial_settlements_placed[pid]
        st["init_roads"][p] = g.initial_roads_placed[pid]
        sc = g.initial_second_settlement_corners[pid]
        st["second_corner"][p] = -1 if sc is None else int(sc)
        if pid in g.current_longest_path:
            st["has_path_key"][p] = 1
            st["cur_longest_path"][p] = g.current_longest_path[pid]
        if pid in g.current_army_size:
            st["cur_army"][p] = g.current_army_size[pid]
        # consistency of redundant reference containers
        n_roads = sum(1 for e in g.board.edges if e.road == pid)
        assert n_roads == len(pl.roads)
    for res in Resource:
        if res != Resource.Empty:
            st["bank"][int(res) - 1] = g.resource_bank[res]
    pile = list(g.development_cards_pile)
    st["deck_n"] = len(pile)
    st["deck"][: len(pile)] = [int(c) for c in pile]
    if g.longest_road is not None:
        st["lr_holder"] = int(g.longest_road["player"])
        st["lr_count"] = g.longest_road["count"]
    if g.largest_army is not None:
        st["la_holder"] = int(g.largest_army["player"])
        st["la_count"] = g.largest_army["count"]
    st["initial_phase"] = int(g.initial_placement_phase)
    st["dice_rolled"] = int(g.dice_rolled_this_turn)
    st["played_dev"] = int(g.played_development_card_this_turn)
    st["must_use_dev"] = int(g.must_use_development_card_ability)
    st["rb_active"] = int(g.road_building_active[0])
    st["rb_count"] = int(g.road_building_active[1])
    st["can_move_robber"] = int(g.can_move_robber)
    st["just_moved_robber"] = int(g.just_moved_robber)
    st["must_respond"] = int(g.must_respond_to_trade)
    st["need_discard"] = int(g.players_need_to_discard)
    st["n_discard"] = len(g.players_to_discard)
    st["discard_queue"][: len(g.players_to_discard)] = [int(x) for x in g.players_to_discard]
    if g.proposed_trade is not None:
        tr = g.proposed_trade
        st["trade_proposer"] = int(tr["player_proposing"])
        st["trade_target"] = int(tr["target_player"])
        st["n_give"] = len(tr["player_proposing_res"])
        st["give"][: len(tr["player_proposing_res"])] = [int(x) for x in tr["player_proposing_res"]]
        st["n_recv"] = len(tr["target_player_res"])
        st["recv"][: len(tr["target_player_res"])] = [int(x) for x in tr["target_player_res"]]
    st["die1"] = 0 if g.die_1 is None else int(g.die_1)
    st["die2"] = 0 if g.die_2 is None else int(g.die_2)
    st["trades_this_turn"] = g.trades_proposed_this_turn
    st["actions_this_turn"] = g.actions_this_turn
    st["turn"] = g.turn
    for c in g.development_cards_bought_this_turn:
        st["bought"][int(c)] += 1
    for pid in PlayerId:
        st["curr_vps"][int(pid) - 1] = env.curr_vps[pid]
    st["winner"] = 0 if env.winner is None else int(env.winner.id)
    return np.frombuffer(st.tobytes(), dtype="<i2").copy()


def obs_to_packed(obs) -> np.ndarray:
    """reference obs dict (wrapper.py:60-83) -> uint8[OBS_STRIDE]; asserts every value is exactly representable."""
    row = np.zeros(L.OBS_STRIDE, dtype=np.uint8)
    ratio = dict(L.OBS_RATIO_COLUMNS)
    for key, off, shape in L.OBS_NUMERIC:
        v = np.asarray(obs[key], dtype=np.float64).reshape(-1)
        assert v.size == int(np.prod(shape)), key
        for i, x in enumerate(v):
            x = x * ratio.get(off + i, 1.0)
            assert x == int(x) and 0 <= x <= 255, (key, i, x)
            row[off + i] = int(x)
    for key, li in L.OBS_LISTS:
        v = np.asarray(obs[key]).reshape(-1)
        n = len(v)
        if n == 1 and v[0] == 0:
            n = 0  # the reference encodes an empty list as [0] (wrapper.py:642-655)
        assert n <= L.OBS_DEV_PAD
        row[L.OBS_DEV_LISTS + li * L.OBS_DEV_PAD: L.OBS_DEV_LISTS + li * L.OBS_DEV_PAD + n] = v[:n]
        row[L.OBS_META + 1 + li] = n
    row[L.OBS_META] = int(obs["player_id"])
    return row


def masks_to_packed(masks) -> np.ndarray:
    """reference mask list (wrapper.py:172-185) -> uint8[MASK_STRIDE]."""
    row = np.zeros(L.MASK_STRIDE, dtype=np.uint8)
    assert len(masks) == 12
    for (off, shape), m in zip(L.MASK_HEADS, masks):
        m = np.asarray(m, dtype=np.float64)
        assert m.shape == tuple(shape), (m.shape, shape)
        assert np.all((m == 0) | (m == 1))
        row[off: off + m.size] = m.reshape(-1).astype(np.uint8)
    return row


def current_actor(env) -> int:
    """game_manager.py:152-159"""
    g = env.game
    if g.players_need_to_discard:
        return int(g.players_to_discard[0])
    if g.must_respond_to_trade:
        return int(g.proposed_trade["target_player"])
    return int(g.players_go)


# ---------------------------------------------------------------------------------------------
# pinned random-legal sampler (the C twin is catan_oracle.c:catan_oracle_sample, the CUDA twin
# is sample_random_kernel).  One Philox block (stream 1, counter = decision index) per decision.
# ---------------------------------------------------------------------------------------------
def _pick(bits, w: int) -> int:
    idx = [i for i, b in enumerate(bits) if b]
    if not idx:
        return 0
    return idx[(w * len(idx)) >> 32]


def sample_action(mask_row: np.ndarray, obs_row: np.ndarray, words) -> np.ndarray:
    """(packed masks, packed obs, 4 philox words) -> int32[ACTION_WORDS]; BASELINE.md §3 sampler."""
    a = np.zeros(L.ACTION_WORDS, dtype=np.int32)
    w0, w1, w2, w3 = words

    def head(i, row=None):
        off, shape = L.MASK_HEADS[i]
        if row is None:
            return mask_row[off: off + int(np.prod(shape))]
        return mask_row[off + row * shape[1]: off + (row + 1) * shape[1]]

    t = _pick(head(0), w0)
    a[L.A_TYPE] = t
    if t == 0:
        a[L.A_CORNER] = _pick(head(1, 0), w1)
    elif t == 2:
        a[L.A_CORNER] = _pick(head(1, 1), w1)
    elif t == 1:
        a[L.A_EDGE] = _pick(head(2), w1)
    elif t == 8:
        a[L.A_TILE] = _pick(head(3), w1)
    elif t == 4:
        card = _pick(head(4), w1)
        a[L.A_CARD] = card
        if card == 4:
            a[L.A_RES_A] = _pick(head(9, 2), w2)
        elif card == 2:
            a[L.A_RES_A] = _pick(head(9, 3), w2)
            a[L.A_RES_B] = _pick(head(10), w3)
    elif t == 5:
        a[L.A_RES_A] = _pick(head(9, 0), w1)
        a[L.A_RES_B] = _pick(head(10), w2)
    elif t == 6:
        a[L.A_PLAYER] = _pick(head(6, 0), w1)
        hand = obs_row[L.OBS_CURRENT_RES + 1: L.OBS_CURRENT_RES + 6] > 0
        a[L.A_GIVE] = 1 + _pick(hand, w2)
        a[L.A_RECV] = 1 + ((w3 * 5) >> 32)
    elif t == 7:
        a[L.A_ACCEPT] = _pick(head(5), w1)
    elif t == 11:
        a[L.A_PLAYER] = _pick(head(6, 1), w1)
    elif t == 12:
        a[L.A_DISCARD] = _pick(head(11), w1)
    return a


def action_to_reference(a: np.ndarray):
    """int32[20] -> the list-of-12 the reference's ``step`` takes (policy.py:192-199 output format)."""
    out = [np.array(int(a[i])) for i in range(7)]
    out.append([np.array(int(a[L.A_GIVE + k])) for k in range(4)])
    out.append([np.array(int(a[L.A_RECV + k])) for k in range(4)])
    out.append(np.array(int(a[L.A_RES_A])))
    out.append(np.array(int(a[L.A_RES_B])))
    out.append(np.array(int(a[L.A_DISCARD])))
    return out


def record_game(seed: int, env_id: int, max_steps: int = 6000, env_kwargs=None, n_games: int = 1):
    """Play ``n_games`` back to back (reset on done) on the reference under shared-Philox RNG and the
    pinned sampler.  Index 0 of state/obs/masks is the situation after the first reset; index t+1 is
    after step t.  When step t ends a game, index t+1 holds the TERMINAL situation (EnvWrapper.step
    semantics) and ``reset_state/obs/masks[k]`` hold the situation after the following reset."""
    R = import_reference()
    game_rng = PhiloxStream(seed, env_id, 0)
    samp = PhiloxStream(seed, env_id, 1)
    env = R["EnvWrapper"](**(env_kwargs or {}))
    PlayerId = R["PlayerId"]
    rec = dict(actions=[], state=[], obs=[], masks=[], reward=[], done=[], rng_ctr=[],
               reset_at=[], reset_state=[], reset_obs=[], reset_masks=[], reset_rng_ctr=[])
    with patched_rng(game_rng):
        obs = env.reset()
        cur_obs = obs_to_packed(obs)
        cur_masks = masks_to_packed(env.get_action_masks())
        rec["state"].append(state_to_vec(env))
        rec["obs"].append(cur_obs)
        rec["masks"].append(cur_masks)
        rec["rng_ctr"].append(game_rng.ctr)
        games = 0
        for t in range(max_steps):
            a = sample_action(cur_masks, cur_obs, samp.block(t))
            obs, reward, done, _ = env.step(action_to_reference(a))
            cur_obs = obs_to_packed(obs)
            cur_masks = masks_to_packed(env.get_action_masks())
            rec["actions"].append(a)
            rec["state"].append(state_to_vec(env))
            rec["obs"].append(cur_obs)
            rec["masks"].append(cur_masks)
            rec["reward"].append([float(reward[PlayerId(p + 1)]) for p in range(4)])
            rec["done"].append(int(done))
            rec["rng_ctr"].append(game_rng.ctr)
            if done:
                games += 1
                if games >= n_games:
                    break
                obs = env.reset()
                cur_obs = obs_to_packed(obs)
                cur_masks = masks_to_packed(env.get_action_masks())
                rec["reset_at"].append(t + 1)
                rec["reset_state"].append(state_to_vec(env))
                rec["reset_obs"].append(cur_obs)
                rec["reset_masks"].append(cur_masks)
                rec["reset_rng_ctr"].append(game_rng.ctr)
    out = dict(
        seed=np.int64(seed), env_id=np.int64(env_id),
        actions=np.asarray(rec["actions"], dtype=np.int32).reshape(-1, L.ACTION_WORDS),
        state=np.asarray(rec["state"], dtype=np.int16),
        obs=np.asarray(rec["obs"], dtype=np.uint8),
        masks=np.asarray(rec["masks"], dtype=np.uint8),
        reward=np.asarray(rec["reward"], dtype=np.float32).reshape(-1, 4),
        done=np.asarray(rec["done"], dtype=np.uint8),
        rng_ctr=np.asarray(rec["rng_ctr"], dtype=np.int64),
        reset_at=np.asarray(rec["reset_at"], dtype=np.int64),
        reset_state=np.asarray(rec["reset_state"], dtype=np.int16).reshape(-1, L.STATE_WORDS),
        reset_obs=np.asarray(rec["reset_obs"], dtype=np.uint8).reshape(-1, L.OBS_STRIDE),
        reset_masks=np.asarray(rec["reset_masks"], dtype=np.uint8).reshape(-1, L.MASK_STRIDE),
        reset_rng_ctr=np.asarray(rec["reset_rng_ctr"], dtype=np.int64),
    )
    return out
