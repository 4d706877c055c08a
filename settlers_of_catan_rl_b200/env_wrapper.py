"""``EnvWrapper`` — single-env adapter with the reference's exact surface (env/wrapper.py:11-50, :168, :711-721).

It is an N=1 view of the same CUDA engine, called through the host-buffer entry points of the C ABI
(``catan_reset_host`` / ``catan_step_host``), so the reference's managers
(``RL/ppo/game_manager.py``, ``RL/ppo/evaluation_manager.py``, ``evaluation/evaluation_manager.py``) can
construct and drive it unchanged: same constructor kwargs, ``reset() -> obs dict``,
``step(action) -> (obs, reward dict, done, info)``, ``get_action_masks() -> list of 12 arrays``,
``save_state()/restore_state()``, and the attributes they reach through the env (``env.game.players_go`` …,
``env.winner.id``, ``env.curr_vps``, ``env.reward_annealing_factor``).  Throughput is not the point of this
class (one game per launch); ``VecCatanEnv`` is the fast path.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np

from . import layout as L
from .enums import ActionTypes, PlayerId, Resource
from .vec_env import VecCatanEnv

_ERR_TEXT = {
    1: "Unknown action type.",
    2: "This action is not allowed in the current phase of the turn.",
    3: "You cannot afford this.",
    4: "You cannot place that here!",
    5: "You cannot play / buy this development card.",
    6: "You do not have these resources / the bank has none left.",
    7: "Cannot steal from a player who doesn't have a building on tile with robber.",
    8: "Bad action head value.",
}

_next_env_id = [0]


class _GameView:
    """the handful of ``env.game.*`` attributes the reference's managers read (game_manager.py:152-159,
    evaluation/evaluation_manager.py:86-88)"""

    def __init__(self, st, owner=None):
        self._owner = owner
        self.players_need_to_discard = bool(st["need_discard"])
        self.players_to_discard = [PlayerId(int(x)) for x in st["discard_queue"][: int(st["n_discard"])]]
        self.must_respond_to_trade = bool(st["must_respond"])
        if int(st["trade_proposer"]):
            self.proposed_trade = {
                "player_proposing": PlayerId(int(st["trade_proposer"])),
                "target_player": PlayerId(int(st["trade_target"])),
                "player_proposing_res": [Resource(int(x)) for x in st["give"][: int(st["n_give"])]],
                "target_player_res": [Resource(int(x)) for x in st["recv"][: int(st["n_recv"])]],
            }
        else:
            self.proposed_trade = None
        self.players_go = PlayerId(int(st["players_go"]))
        self.player_order = [PlayerId(int(x)) for x in st["player_order"]]
        self.initial_placement_phase = bool(st["initial_phase"])
        self.initial_settlements_placed = {PlayerId(p + 1): int(st["init_settlements"][p]) for p in range(4)}
        self.initial_roads_placed = {PlayerId(p + 1): int(st["init_roads"][p]) for p in range(4)}
        self.dice_rolled_this_turn = bool(st["dice_rolled"])
        self.turn = int(st["turn"])
        self.die_1, self.die_2 = int(st["die1"]) or None, int(st["die2"]) or None
        self.victory_points = {PlayerId(p + 1): int(st["vp"][p]) for p in range(4)}

    def randomise_uncertainty(self, controlling_player_id):
        """game/game.py:1207-1282, as the forward-search worker calls it (RL/forward_search_policy/worker.py:44)"""
        self._owner._randomise_uncertainty(int(controlling_player_id))


class EnvWrapper(object):
    def __init__(self, interactive=False, max_actions_per_turn=None, max_proposed_trades_per_turn=4,
                 validate_actions=True, debug_mode=False, win_reward=500, dense_reward=False, policies=None,
                 device="cuda:0", seed=0, env_id=None, _engine=None):
        if interactive:
            raise NotImplementedError("the pygame UI is outside the hot path (SURVEY.md §2 row 16)")
        if env_id is None:
            env_id = _next_env_id[0]
            _next_env_id[0] += 1
        self.max_actions_per_turn = np.inf if max_actions_per_turn is None else max_actions_per_turn
        self.max_proposed_trades_per_turn = max_proposed_trades_per_turn
        self.validate_actions = validate_actions
        self.win_reward = win_reward
        self.dense_reward = dense_reward
        cfg = dict(
            auto_reset=0,
            max_actions_per_turn=-1 if max_actions_per_turn is None else int(max_actions_per_turn),
            max_proposed_trades_per_turn=-1 if max_proposed_trades_per_turn is None else int(max_proposed_trades_per_turn),
            validate_actions=int(bool(validate_actions)), dense_reward=int(bool(dense_reward)), win_reward=float(win_reward))
        # `_engine` is a TEST hook: tests/host_emu compiles the product's game logic (csrc/catan_game.cuh) with g++ so that the
        # reference's unchanged managers can be run over this adapter in the CPU-only container.  The product always runs the
        # CUDA engine; there is no CPU path in the package.
        self._vec = _engine(seed=seed, env_id=env_id, **cfg) if _engine is not None else VecCatanEnv(
            1, device=device, seed=seed, first_env_id=env_id, **cfg)
        self._obs = np.zeros((1, L.OBS_STRIDE), np.uint8)
        self._masks = np.zeros((1, L.MASK_STRIDE), np.uint8)
        self._reward = np.zeros((1, 4), np.float32)
        self._info = np.zeros((1, L.INFO_STRIDE), np.uint8)
        self._game = None
        self.winner = None
        self.curr_vps = {PlayerId.White: 0, PlayerId.Red: 0, PlayerId.Blue: 0, PlayerId.Orange: 0}
        # The reference's constructor already holds a freshly reset Game (game.py:24).  Here the first game is dealt by
        # the first reset() so that construction does not consume the (seed, env_id) stream; anything that needs a
        # game before that triggers it.
        self._started = False

    def _ensure_started(self):
        if not self._started:
            self.reset()

    # ---- attributes the managers touch
    @property
    def reward_annealing_factor(self):
        return float(self._vec.config.reward_annealing_factor)

    @reward_annealing_factor.setter
    def reward_annealing_factor(self, value):
        self._vec.set_reward_annealing_factor(value)

    @property
    def game(self):
        self._ensure_started()
        if self._game is None:
            self._game = _GameView(self._vec.export_state().view(L.STATE_DTYPE)[0, 0], self)
        return self._game

    # ---- EnvWrapper API
    def reset(self):
        self._started = True
        self._vec.reset_host(self._obs, self._masks, self._info)
        self._game = None
        self.winner = None
        self.curr_vps = {PlayerId.White: 0, PlayerId.Red: 0, PlayerId.Blue: 0, PlayerId.Orange: 0}
        return self._obs_dict()

    def step(self, action):
        self._ensure_started()
        a = self.pack_action(action)
        self._vec.step_host(a, self._obs, self._masks, self._reward, self._info)
        self._game = None
        info = self._info[0]
        if info[L.INFO_ERR]:
            raise RuntimeError(_ERR_TEXT.get(int(info[L.INFO_ERR]), "invalid action"))   # wrapper.py:38-41
        done = bool(info[L.INFO_DONE])
        self.curr_vps = {PlayerId(p + 1): int(info[L.INFO_FINAL_VP + p]) for p in range(4)}
        if info[L.INFO_WINNER]:
            self.winner = SimpleNamespace(id=PlayerId(int(info[L.INFO_WINNER])))
        reward = {PlayerId(p + 1): float(self._reward[0, p]) for p in range(4)}
        log = {"player_id": PlayerId(int(info[L.INFO_ACTED])), "text": ActionTypes(int(info[L.INFO_ACT_TYPE])).name}
        return self._obs_dict(), reward, done, {"log": log}

    def get_action_masks(self):
        self._ensure_started()
        row = self._masks[0]
        return [row[off:off + int(np.prod(shape))].reshape(shape).astype(np.float64) for off, shape in L.MASK_HEADS]

    def save_state(self):
        self._ensure_started()
        return {"state": self._vec.export_state()[0].copy(), "vps": dict(self.curr_vps), "winner": self.winner}

    def restore_state(self, state):
        self._started = True
        self._vec.import_state(state["state"][None, :])
        self._obs[:], self._masks[:] = self._vec.rows_host()
        self._game = None
        self.curr_vps = dict(state["vps"])
        self.winner = state["winner"]

    def _randomise_uncertainty(self, pid: int) -> None:
        self._vec.randomise_uncertainty(pid)
        self._obs[:], self._masks[:] = self._vec.rows_host()
        self._game = None

    def render(self):
        raise NotImplementedError("rendering is outside the hot path (SURVEY.md §2 row 16)")

    # ---- format conversion
    @staticmethod
    def pack_action(action) -> np.ndarray:
        """list-of-12 composite action (policy.py:192-199 format) -> int32[1, 20]"""
        a = np.zeros((1, L.ACTION_WORDS), np.int32)
        for i in range(7):
            a[0, i] = int(action[i])
        for k in range(4):
            a[0, L.A_GIVE + k] = int(action[7][k])
            a[0, L.A_RECV + k] = int(action[8][k])
        a[0, L.A_RES_A], a[0, L.A_RES_B], a[0, L.A_DISCARD] = int(action[9]), int(action[10]), int(action[11])
        return a

    def _obs_dict(self):
        """packed row -> the reference's obs dict (wrapper.py:60-83): float64 arrays, list of 19 tile arrays,
        int arrays for the card lists ([0] when empty), PlayerId for player_id"""
        row = self._obs[0]
        f = row[:L.OBS_FEATURES].astype(np.float64)
        for col, div in L.OBS_RATIO_COLUMNS:
            f[col] /= div
        obs = {}
        for key, off, shape in L.OBS_NUMERIC:
            obs[key] = f[off:off + int(np.prod(shape))].reshape(shape).copy()
        obs["tile_representations"] = [t for t in obs["tile_representations"]]
        for key, li in L.OBS_LISTS:
            n = int(row[L.OBS_META + 1 + li])
            a = L.OBS_DEV_LISTS + li * L.OBS_DEV_PAD
            obs[key] = row[a:a + max(n, 1)].astype(int)
        obs["player_id"] = PlayerId(int(row[L.OBS_META]))
        return obs
