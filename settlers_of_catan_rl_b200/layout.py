"""Python mirror of ``include/catan_layout.h``: packed observation / mask / action / state layouts.

The numbers here are the contract between the CUDA kernels and the PyTorch side; a CPU test
(`tests/test_abi.py`) checks every constant against the values the built library reports.

Reference citations: observation pieces ``env/wrapper.py:52-83, :491-709``; mask heads
``env/wrapper.py:168-185``; action heads ``env/wrapper.py:114-166``; state ``game/game.py:1013-1091``.
"""
from __future__ import annotations

import numpy as np

N_TILES, N_CORNERS, N_EDGES, N_HARBOURS, N_PLAYERS, N_RES, N_DEV, DECK = 19, 54, 72, 9, 4, 5, 5, 25
N_ACTION_TYPES = 13

ACTION_WORDS = 20
A_TYPE, A_CORNER, A_EDGE, A_TILE, A_CARD, A_ACCEPT, A_PLAYER = 0, 1, 2, 3, 4, 5, 6
A_GIVE, A_RECV, A_RES_A, A_RES_B, A_DISCARD = 7, 11, 15, 16, 17

# ---- packed observation row (uint8)
OBS_PROPOSED_TRADE = 0
OBS_CURRENT_RES = 12
OBS_TILES = 18
OBS_TILE_DIM = 60
OBS_CUR_MAIN = 1158
OBS_CUR_MAIN_DIM = 152
OBS_OTHER_MAIN = 1310
OBS_OTHER_MAIN_DIM = 159
OBS_DEV_LISTS = 1787
OBS_DEV_PAD = 25
OBS_META = 1912
OBS_FEATURES = 1787
OBS_STRIDE = 1920
OBS_CUR_LR_LEN, OBS_CUR_ARMY_LEN = 101, 103
OBS_OTH_LR_LEN, OBS_OTH_ARMY_LEN = 141, 143

#: (key, offset, shape) of the numeric observation entries, in the reference's key order
#: (RL/ppo/process_batch.py:10-13).
OBS_NUMERIC = (
    ("proposed_trade", OBS_PROPOSED_TRADE, (12,)),
    ("current_resources", OBS_CURRENT_RES, (6,)),
    ("tile_representations", OBS_TILES, (N_TILES, OBS_TILE_DIM)),
    ("current_player_main", OBS_CUR_MAIN, (OBS_CUR_MAIN_DIM,)),
    ("next_player_main", OBS_OTHER_MAIN, (OBS_OTHER_MAIN_DIM,)),
    ("next_next_player_main", OBS_OTHER_MAIN + OBS_OTHER_MAIN_DIM, (OBS_OTHER_MAIN_DIM,)),
    ("next_next_next_player_main", OBS_OTHER_MAIN + 2 * OBS_OTHER_MAIN_DIM, (OBS_OTHER_MAIN_DIM,)),
)
#: (key, index of the list inside the dev-list block)
OBS_LISTS = (
    ("current_player_played_dev", 0),
    ("current_player_hidden_dev", 1),
    ("next_player_played_dev", 2),
    ("next_next_player_played_dev", 3),
    ("next_next_next_player_played_dev", 4),
)
#: absolute byte offsets whose stored value is a raw count that the reference divides (by 8 / by 4)
OBS_RATIO_COLUMNS = (
    (OBS_CUR_MAIN + OBS_CUR_LR_LEN, 8.0),
    (OBS_CUR_MAIN + OBS_CUR_ARMY_LEN, 4.0),
) + tuple(
    (OBS_OTHER_MAIN + k * OBS_OTHER_MAIN_DIM + off, div)
    for k in range(3)
    for off, div in ((OBS_OTH_LR_LEN, 8.0), (OBS_OTH_ARMY_LEN, 4.0))
)

# ---- packed mask row (uint8)
MASK_HEADS = (
    (0, (13,)),
    (13, (3, N_CORNERS)),
    (175, (N_EDGES + 1,)),
    (248, (N_TILES,)),
    (267, (5,)),
    (272, (2,)),
    (274, (3, 3)),
    (283, (6,)),
    (289, (6,)),
    (295, (4, 5)),
    (315, (5,)),
    (320, (5,)),
)
MASK_ENTRIES = 325
MASK_STRIDE = 336

# ---- per-step info row (uint8)
INFO_DONE, INFO_WINNER, INFO_FINAL_VP, INFO_ACTOR, INFO_ACTED = 0, 1, 2, 6, 7
INFO_ACT_TYPE, INFO_ROLL, INFO_ERR, INFO_RESET, INFO_ACTOR_PRE = 8, 9, 10, 11, 12
INFO_STRIDE = 16

# ---- canonical state (int16 fields, C order) — mirrors ``catan_state_t``
STATE_DTYPE = np.dtype(
    [
        ("tile_res", "<i2", (N_TILES,)),
        ("tile_val", "<i2", (N_TILES,)),
        ("robber_tile", "<i2"),
        ("corner_type", "<i2", (N_CORNERS,)),
        ("corner_owner", "<i2", (N_CORNERS,)),
        ("edge_owner", "<i2", (N_EDGES,)),
        ("harbour_perm", "<i2", (N_HARBOURS,)),
        ("player_order", "<i2", (4,)),
        ("player_order_id", "<i2"),
        ("players_go", "<i2"),
        ("res", "<i2", (4, N_RES)),
        ("vis", "<i2", (4, N_RES)),
        ("est_min", "<i2", (4, 3, N_RES)),
        ("est_max", "<i2", (4, 3, N_RES)),
        ("vp", "<i2", (4,)),
        ("harbours", "<i2", (4,)),
        ("n_hidden", "<i2", (4,)),
        ("hidden", "<i2", (4, DECK)),
        ("n_played", "<i2", (4,)),
        ("played", "<i2", (4, DECK)),
        ("settlements_left", "<i2", (4,)),
        ("cities_left", "<i2", (4,)),
        ("init_settlements", "<i2", (4,)),
        ("init_roads", "<i2", (4,)),
        ("second_corner", "<i2", (4,)),
        ("cur_longest_path", "<i2", (4,)),
        ("has_path_key", "<i2", (4,)),
        ("cur_army", "<i2", (4,)),
        ("bank", "<i2", (N_RES,)),
        ("deck_n", "<i2"),
        ("deck", "<i2", (DECK,)),
        ("lr_holder", "<i2"),
        ("lr_count", "<i2"),
        ("la_holder", "<i2"),
        ("la_count", "<i2"),
        ("initial_phase", "<i2"),
        ("dice_rolled", "<i2"),
        ("played_dev", "<i2"),
        ("must_use_dev", "<i2"),
        ("rb_active", "<i2"),
        ("rb_count", "<i2"),
        ("can_move_robber", "<i2"),
        ("just_moved_robber", "<i2"),
        ("must_respond", "<i2"),
        ("need_discard", "<i2"),
        ("n_discard", "<i2"),
        ("discard_queue", "<i2", (4,)),
        ("trade_proposer", "<i2"),
        ("trade_target", "<i2"),
        ("n_give", "<i2"),
        ("give", "<i2", (4,)),
        ("n_recv", "<i2"),
        ("recv", "<i2", (4,)),
        ("die1", "<i2"),
        ("die2", "<i2"),
        ("trades_this_turn", "<i2"),
        ("actions_this_turn", "<i2"),
        ("turn", "<i2"),
        ("bought", "<i2", (N_DEV,)),
        ("curr_vps", "<i2", (4,)),
        ("winner", "<i2"),
        ("rng_ctr_lo", "<i2"),
        ("rng_ctr_hi", "<i2"),
    ]
)
STATE_WORDS = STATE_DTYPE.itemsize // 2
#: fields that are not part of the reference's state (excluded from parity comparisons)
STATE_NON_REFERENCE_FIELDS = ("rng_ctr_lo", "rng_ctr_hi")

# the reference's default EnvWrapper kwargs (env/wrapper.py:12-13)
DEFAULT_CONFIG = dict(
    max_actions_per_turn=-1,
    max_proposed_trades_per_turn=4,
    validate_actions=1,
    dense_reward=0,
    auto_reset=1,
    win_reward=500.0,
    reward_annealing_factor=1.0,
)
