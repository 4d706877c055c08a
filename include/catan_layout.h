/*
 * catan_layout.h — shared constants and plain-C data formats for the Catan
 * env-step hot path.  Included by the CUDA product (settlers_of_catan_rl_b200/csrc),
 * by the C-ABI header (catan_b200.h) and by the test-only CPU oracle (oracle/).
 *
 * Index conventions (all derived from the reference's enums, game/enums.py:4-50):
 *   player index  p = PlayerId - 1      : 0 White, 1 Blue, 2 Orange, 3 Red      (enums.py:8-12)
 *   resource index r = Resource - 1     : 0 Brick, 1 Wood, 2 Ore, 3 Sheep, 4 Wheat (enums.py:22-28)
 *       -> identical to the policy's resource-head order (env/wrapper.py:414-426)
 *   tile resource = Resource enum value : 0 Empty(desert) .. 5 Wheat; numerically equal to
 *       the Terrain enum (enums.py:14-28, game/components/tile.py:6)
 *   development card: 0 Knight, 1 VictoryPoint, 2 YearOfPlenty, 3 RoadBuilding, 4 Monopoly (enums.py:30-35)
 *   action type: 0..12 ActionTypes (enums.py:37-50)
 *   relative-seat label: 0 "next", 1 "next_next", 2 "next_next_next" (game/components/player.py:13-19)
 */
#ifndef CATAN_LAYOUT_H
#define CATAN_LAYOUT_H

#include <stdint.h>

#define CATAN_N_TILES    19
#define CATAN_N_CORNERS  54
#define CATAN_N_EDGES    72
#define CATAN_N_HARBOURS 9
#define CATAN_N_PLAYERS  4
#define CATAN_N_RES      5
#define CATAN_N_DEV      5
#define CATAN_DECK       25
#define CATAN_N_ACTION_TYPES 13

/* action types (enums.py:37-50) */
enum {
  CATAN_ACT_PLACE_SETTLEMENT = 0, CATAN_ACT_PLACE_ROAD = 1, CATAN_ACT_UPGRADE_CITY = 2,
  CATAN_ACT_BUY_DEV = 3, CATAN_ACT_PLAY_DEV = 4, CATAN_ACT_EXCHANGE = 5, CATAN_ACT_PROPOSE_TRADE = 6,
  CATAN_ACT_RESPOND = 7, CATAN_ACT_MOVE_ROBBER = 8, CATAN_ACT_ROLL_DICE = 9, CATAN_ACT_END_TURN = 10,
  CATAN_ACT_STEAL = 11, CATAN_ACT_DISCARD = 12
};
enum { CATAN_DEV_KNIGHT = 0, CATAN_DEV_VP = 1, CATAN_DEV_YOP = 2, CATAN_DEV_ROADBUILDING = 3, CATAN_DEV_MONOPOLY = 4 };
enum { CATAN_RES_BRICK = 0, CATAN_RES_WOOD = 1, CATAN_RES_ORE = 2, CATAN_RES_SHEEP = 3, CATAN_RES_WHEAT = 4 };

/* ---- composite action, int32[CATAN_ACTION_WORDS] per env (env/wrapper.py:114-166, RL/models/policy.py:192-199)
 * word 0 type | 1 corner | 2 edge (72 = dummy) | 3 tile | 4 dev card | 5 accept(0)/reject(1) |
 * 6 relative player 0..2 | 7..10 give list (0 = stop, 1..5 = resource index+1) | 11..14 receive list |
 * 15 resource A | 16 resource B | 17 discard resource | 18,19 unused */
#define CATAN_ACTION_WORDS 20
#define CATAN_A_TYPE 0
#define CATAN_A_CORNER 1
#define CATAN_A_EDGE 2
#define CATAN_A_TILE 3
#define CATAN_A_CARD 4
#define CATAN_A_ACCEPT 5
#define CATAN_A_PLAYER 6
#define CATAN_A_GIVE 7
#define CATAN_A_RECV 11
#define CATAN_A_RES_A 15
#define CATAN_A_RES_B 16
#define CATAN_A_DISCARD 17

/* ---- packed observation row, uint8[CATAN_OBS_STRIDE] per env (env/wrapper.py:52-83, :491-709)
 * Every feature is stored as the small non-negative integer the reference computes, with two
 * exceptions that the reference stores as ratios: longest-road length (ref: count/8.0) and army
 * size (ref: knights/4.0) are stored as the raw count; the decode multiplies by 1/8 and 1/4 (exact
 * in binary floating point). */
#define CATAN_OBS_PROPOSED_TRADE 0      /* 12  wrapper.py:61-69 */
#define CATAN_OBS_CURRENT_RES    12     /* 6   wrapper.py:70-71 */
#define CATAN_OBS_TILES          18     /* 19*60  wrapper.py:491-524 */
#define CATAN_OBS_TILE_DIM       60
#define CATAN_OBS_CUR_MAIN       1158   /* 152 wrapper.py:698-702 */
#define CATAN_OBS_CUR_MAIN_DIM   152
#define CATAN_OBS_OTHER_MAIN     1310   /* 3*159 wrapper.py:703-707 */
#define CATAN_OBS_OTHER_MAIN_DIM 159
#define CATAN_OBS_DEV_LISTS      1787   /* 5 lists * 25: cur played, cur hidden, next/nn/nnn played; value card+1, 0 pad */
#define CATAN_OBS_DEV_PAD        25
#define CATAN_OBS_META           1912   /* [0] acting PlayerId; [1..5] true lengths of the 5 lists; [6,7] zero */
#define CATAN_OBS_FEATURES       1787
#define CATAN_OBS_STRIDE         1920
/* offsets of the two ratio features inside a "main" block */
#define CATAN_OBS_CUR_LR_LEN     (50 + 50 + 1)   /* current: res 40 + vp 10 + prod 50 + [holder, len] */
#define CATAN_OBS_CUR_ARMY_LEN   (50 + 50 + 3)
#define CATAN_OBS_OTH_LR_LEN     (90 + 50 + 1)   /* other: min 40 + max 40 + vp 10 + prod 50 + [holder, len] */
#define CATAN_OBS_OTH_ARMY_LEN   (90 + 50 + 3)

/* ---- packed legal-action mask row, uint8[CATAN_MASK_STRIDE] per env (env/wrapper.py:168-185) */
#define CATAN_MASK_TYPE     0    /* 13 */
#define CATAN_MASK_CORNER   13   /* 3*54 */
#define CATAN_MASK_EDGE     175  /* 73 */
#define CATAN_MASK_TILE     248  /* 19 */
#define CATAN_MASK_DEV      267  /* 5 */
#define CATAN_MASK_ACCEPT   272  /* 2 */
#define CATAN_MASK_PLAYER   274  /* 3*3 */
#define CATAN_MASK_GIVE     283  /* 6 */
#define CATAN_MASK_RECV     289  /* 6 */
#define CATAN_MASK_RES_A    295  /* 4*5 */
#define CATAN_MASK_RES_B    315  /* 5 */
#define CATAN_MASK_DISCARD  320  /* 5 */
#define CATAN_MASK_ENTRIES  325
#define CATAN_MASK_STRIDE   336

/* ---- per-step info row, uint8[CATAN_INFO_STRIDE] per env */
#define CATAN_INFO_DONE        0   /* 1 if the applied action ended the game (wrapper.py:85-91) */
#define CATAN_INFO_WINNER      1   /* PlayerId of env.winner at the end of the step, 0 = none */
#define CATAN_INFO_FINAL_VP    2   /* [2..5] victory points per player index after the step (before any auto-reset) */
#define CATAN_INFO_ACTOR       6   /* PlayerId that takes the NEXT decision (game_manager.py:152-159) */
#define CATAN_INFO_ACTED       7   /* PlayerId that took THIS decision */
#define CATAN_INFO_ACT_TYPE    8   /* action type applied */
#define CATAN_INFO_ROLL        9   /* die_1 + die_2 if this step rolled, else 0 */
#define CATAN_INFO_ERR         10  /* non-zero: action rejected by validation this step (state unchanged) */
#define CATAN_INFO_RESET       11  /* 1 if the env was auto-reset inside this step */
#define CATAN_INFO_ACTOR_PRE   12  /* PlayerId that would act next in the state BEFORE any auto-reset (game_manager.py:99) */
#define CATAN_INFO_STRIDE      16

/* error codes stored in CATAN_INFO_ERR and OR-ed (1<<code) into the sticky err_flags */
enum {
  CATAN_ERR_NONE = 0, CATAN_ERR_BAD_TYPE = 1, CATAN_ERR_PHASE = 2, CATAN_ERR_CANNOT_AFFORD = 3,
  CATAN_ERR_BAD_LOCATION = 4, CATAN_ERR_BAD_CARD = 5, CATAN_ERR_BAD_RESOURCE = 6, CATAN_ERR_BAD_TARGET = 7,
  CATAN_ERR_BAD_HEAD_VALUE = 8,
  CATAN_ERR_NO_DEAL = 9       /* catan_randomise_uncertainty: the controlling player's beliefs admit no consistent deal */
};

/* ---- canonical unpacked game state (== Game.save_current_state, game/game.py:1013-1091, plus the
 * wrapper's curr_vps / winner, env/wrapper.py:711-716).  A flat int16 array: this is what
 * catan_export_state / catan_import_state move and what the parity tests compare field by field. */
typedef struct catan_state {
  int16_t tile_res[CATAN_N_TILES];        /* Resource enum 0..5 */
  int16_t tile_val[CATAN_N_TILES];        /* 2..12, desert 7 */
  int16_t robber_tile;
  int16_t corner_type[CATAN_N_CORNERS];   /* 0 none, 1 settlement, 2 city */
  int16_t corner_owner[CATAN_N_CORNERS];  /* 0 none, else PlayerId */
  int16_t edge_owner[CATAN_N_EDGES];      /* 0 none, else PlayerId */
  int16_t harbour_perm[CATAN_N_HARBOURS]; /* harbour id sitting at slot i (board.py:29-33,154) */
  int16_t player_order[4];                /* PlayerId at seat i */
  int16_t player_order_id;
  int16_t players_go;                     /* PlayerId */
  int16_t res[4][CATAN_N_RES];            /* [p][r] hands */
  int16_t vis[4][CATAN_N_RES];            /* visible_resources */
  int16_t est_min[4][3][CATAN_N_RES];     /* opponent_min_res[observer p][label][r] */
  int16_t est_max[4][3][CATAN_N_RES];
  int16_t vp[4];
  int16_t harbours[4];                    /* bit 0 any 3:1, bit (r+1) 2:1 for resource r */
  int16_t n_hidden[4];
  int16_t hidden[4][CATAN_DECK];          /* ordered, zero beyond n */
  int16_t n_played[4];
  int16_t played[4][CATAN_DECK];
  int16_t settlements_left[4];
  int16_t cities_left[4];
  int16_t init_settlements[4];
  int16_t init_roads[4];
  int16_t second_corner[4];               /* -1 = None */
  int16_t cur_longest_path[4];
  int16_t has_path_key[4];                /* key present in the defaultdict (wrapper.py:619) */
  int16_t cur_army[4];
  int16_t bank[CATAN_N_RES];
  int16_t deck_n;
  int16_t deck[CATAN_DECK];               /* pile, popped from index deck_n-1 (game.py:707) */
  int16_t lr_holder, lr_count, la_holder, la_count; /* holder PlayerId, 0 = None */
  int16_t initial_phase, dice_rolled, played_dev, must_use_dev, rb_active, rb_count;
  int16_t can_move_robber, just_moved_robber, must_respond, need_discard;
  int16_t n_discard;
  int16_t discard_queue[4];               /* PlayerIds, zero beyond n */
  int16_t trade_proposer, trade_target;   /* PlayerIds, 0 = no trade */
  int16_t n_give;
  int16_t give[4];                        /* resource index + 1 */
  int16_t n_recv;
  int16_t recv[4];
  int16_t die1, die2;                     /* 0 = None */
  int16_t trades_this_turn, actions_this_turn, turn;
  int16_t bought[CATAN_N_DEV];            /* development_cards_bought_this_turn as counts per card */
  int16_t curr_vps[4];                    /* wrapper.curr_vps */
  int16_t winner;                         /* wrapper.winner PlayerId, 0 = None */
  int16_t rng_ctr_lo, rng_ctr_hi;         /* game-stream Philox draw counter (not part of the reference state) */
} catan_state_t;
#define CATAN_STATE_WORDS ((int)(sizeof(catan_state_t) / sizeof(int16_t)))

/* environment options == EnvWrapper.__init__ kwargs (env/wrapper.py:12-28) */
typedef struct catan_config {
  int32_t max_actions_per_turn;          /* < 0 : None (np.inf) */
  int32_t max_proposed_trades_per_turn;  /* < 0 : None; default 4 */
  int32_t validate_actions;              /* default 1 */
  int32_t dense_reward;                  /* default 0 */
  int32_t auto_reset;                    /* 1: reset finished games inside step (vector API); 0: EnvWrapper.step semantics */
  float   win_reward;                    /* default 500 */
  float   reward_annealing_factor;       /* default 1.0 */
} catan_config_t;

/* ---- pinned RNG definition (replaces np.random / random; SURVEY 8c "shared-Philox mode")
 * Philox4x32-10, key = (seed_lo, seed_hi), counter = (draw >> 2, stream, env_lo, env_hi);
 * draw d of a stream uses output word (d & 3).  stream 0 = game (reset shuffles, dice, steal),
 * stream 1 = random-legal action sampler.
 *   bounded(n)    = (u32 * n) >> 32
 *   shuffle(a, n) = for i = n-1 .. 1 : j = bounded(i + 1); swap(a[i], a[j])
 *   die           = 1 + bounded(6)  (die_1 then die_2, game.py:139-140)
 *   steal         = list[bounded(len)] over the Brick,Wheat,Wood,Sheep,Ore multiset (game.py:638-643) */
#define CATAN_STREAM_GAME 0u
#define CATAN_STREAM_SAMPLER 1u

#endif /* CATAN_LAYOUT_H */
