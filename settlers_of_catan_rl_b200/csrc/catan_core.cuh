// catan_core.cuh — the Catan env-step engine: one warp advances one game.
//
// The same source compiles two ways:
//   * nvcc (product): every function is __device__, CATAN_LANES == 32, lanes of one warp cooperate
//     on one game whose packed record lives in shared memory;
//   * g++ with -DCATAN_HOST_EMU (tests only, tests/host_emu/): CATAN_LANES == 1, warp collectives
//     become identities.  This lets the CPU test-suite run exactly this logic against the oracle and
//     the golden fixtures before any GPU time is spent.  It is not a shipped fallback: the C-ABI
//     library is only ever built from the .cu files.
//
// Lane contracts used below:  [L0] = called by lane 0 only;  [W] = called by all lanes of the warp
// (contains warp syncs / collectives);  [P] = pure per-element predicate, any lane.
//
// Reference behaviour reproduced here (file:line are in /root/reference): game/game.py (Game),
// game/components/{board,corner,edge,player}.py, env/wrapper.py (EnvWrapper).  See SURVEY.md §8a.
#pragma once

#include <stdint.h>
#include <string.h>

#include "../../include/catan_layout.h"
#define CATAN_TOPO_MACROS_ONLY
#include "../../include/catan_topology.h"

#if defined(__CUDACC__) && !defined(CATAN_HOST_EMU)
#define CATAN_DEVICE 1
#define CATAN_FN __device__ __forceinline__
// phase-sized functions are emitted ONCE and called: the step kernel must stay small enough for the
// instruction caches (profiles/r1_notes.md: 60% of stalls were stall_no_inst with everything inlined)
#define CATAN_FN_NOINLINE __device__ __noinline__
#define CATAN_NO_UNROLL _Pragma("unroll 1")
#define CATAN_LANES 32
#else
#define CATAN_FN static inline
#define CATAN_FN_NOINLINE static
#define CATAN_NO_UNROLL
#define CATAN_LANES 1
#endif

namespace catanb {

enum { WHITE = 1, BLUE = 2, ORANGE = 3, RED = 4 };
enum { BRICK = 0, WOOD = 1, ORE = 2, SHEEP = 3, WHEAT = 4 };

// ------------------------------------------------------------------------------------------------
// constant board topology (generated from the reference's Board)
// ------------------------------------------------------------------------------------------------
struct Topo {
  int8_t tile_corners[19][6];
  int8_t edge_corners[72][2];
  int8_t corner_neigh[54][3];
  int8_t corner_neigh_edge[54][3];
  int8_t corner_tiles[54][3];
  int8_t corner_harbour_slot[54];
  int8_t tile_neigh[19][6];
  int8_t number_placement[19];
  int8_t default_number_order[18];
  int8_t terrain_to_place[19];
  int8_t harbour_res[9];
  int8_t deck_init[25];
  int8_t pad[6];
};
static_assert(sizeof(Topo) == 1008, "Topo must stay a multiple of 16 bytes");
#define CATAN_TOPO_INITIALIZER                                                                       \
  {CATAN_TILE_CORNERS_INIT, CATAN_EDGE_CORNERS_INIT, CATAN_CORNER_NEIGH_INIT, CATAN_CORNER_NEIGH_EDGE_INIT, \
   CATAN_CORNER_TILES_INIT, CATAN_CORNER_HARBOUR_SLOT_INIT, CATAN_TILE_NEIGH_INIT, CATAN_NUMBER_PLACEMENT_INIT, \
   CATAN_DEFAULT_NUMBER_ORDER_INIT, CATAN_TERRAIN_TO_PLACE_INIT, CATAN_HARBOUR_RES_INIT, CATAN_DECK_INIT_INIT, \
   {0, 0, 0, 0, 0, 0}}

// ------------------------------------------------------------------------------------------------
// packed per-game record: 832 bytes (6.5 x 128-B lines), contiguous per game so that one warp's
// load/store of its game is fully coalesced.  Field meaning == catan_state_t (catan_layout.h).
// ------------------------------------------------------------------------------------------------
struct alignas(16) GameRec {
  int16_t est_min[4][3][5];   // opponent_min_res[observer][label][r]        (player.py:38-43)
  int16_t est_max[4][3][5];
  int16_t vis[4][5];          // visible_resources (unbounded growth through trades -> 16 bit)
  uint32_t rng_ctr;           // game-stream Philox draw counter
  uint32_t decision_ctr;      // sampler-stream decision index
  uint32_t episode_steps;     // env steps since the last reset
  uint16_t actions_this_turn;
  uint16_t turn;
  uint8_t corner[54];         // (owner PlayerId << 2) | type (0 none, 1 settlement, 2 city)
  uint8_t edge[72];           // road owner PlayerId, 0 none
  uint8_t tile_res[19];
  uint8_t tile_val[19];
  uint8_t harbour_perm[9];
  uint8_t robber_tile;
  uint8_t res[4][5];
  int8_t vp[4];
  uint8_t harbours[4];
  uint8_t n_hidden[4];
  uint8_t n_played[4];
  uint8_t settlements_left[4];
  uint8_t cities_left[4];
  uint8_t init_settlements[4];
  uint8_t init_roads[4];
  int8_t second_corner[4];
  uint8_t cur_longest_path[4];
  uint8_t has_path_key[4];
  uint8_t cur_army[4];
  uint8_t hidden[4][25];
  uint8_t played[4][25];
  uint8_t bank[5];
  uint8_t deck_n;
  uint8_t deck[25];
  uint8_t player_order[4];
  uint8_t player_order_id, players_go;
  uint8_t lr_holder, lr_count, la_holder, la_count;
  uint8_t initial_phase, dice_rolled, played_dev, must_use_dev, rb_active, rb_count;
  uint8_t can_move_robber, just_moved_robber, must_respond, need_discard;
  uint8_t n_discard;
  uint8_t discard_queue[4];
  uint8_t trade_proposer, trade_target, n_give;
  uint8_t give[4];
  uint8_t n_recv;
  uint8_t recv[4];
  uint8_t die1, die2, trades_this_turn;
  uint8_t bought[5];
  int8_t curr_vps[4];
  uint8_t winner;
  uint8_t pad_[13];
};
static_assert(sizeof(GameRec) == 832, "GameRec layout changed: keep it a multiple of 16 bytes and update DESIGN.md");

// translated action (wrapper.py:114-166)
struct Act {
  int8_t type, corner, edge /* -1 = dummy */, tile, card, accept, target_pid, res_a, res_b, rate, discard;
  int8_t n_give, n_recv;
  int8_t give[4], recv[4];   // resource indices
};

struct EstReq {              // one update_player_resource_estimates call (game.py:921-971), deferred
  int8_t delta[5];
  uint8_t touched;           // bit r: resource r is a key of the `resources` dict
  uint8_t owner, thief;      // PlayerIds; thief 0 = None
  int16_t T_o, T_t;          // hand totals of owner / thief at call time
};

enum { EST_SPECIAL_NONE = 0, EST_SPECIAL_DICE = 1, EST_SPECIAL_MONOPOLY = 2 };

// per-warp scratch (shared memory on the device)
struct alignas(16) WarpScratch {
  int32_t action[CATAN_ACTION_WORDS];
  int32_t alloc[5][4];       // dice payout [r][player index]      (game.py:153-167)
  int16_t dice_T[4][5];      // clip bound of the (r, player) belief update of this roll
  int16_t mono_T[4];
  EstReq est[2];
  uint8_t n_est, est_special, granted, dice_roll;
  uint8_t mono_pid, mono_res;
  uint8_t mono_lost[4];
  uint8_t lr_pid;            // longest road must be re-evaluated for this PlayerId (0 = no)
  uint8_t err;
  uint8_t seat[5];           // seat of PlayerId (index 1..4)
  uint8_t acted_pid, act_type, roll_info, did_reset, done;
  Act act;                   // translated action (wrapper.py:114-166), filled by step_begin
  uint8_t lr_len, lr_shrunk;  // block-cooperative longest-road search: length of lr_pid, "holder's path shrank"
  uint8_t lr_other[5];       // lengths of the other players (index PlayerId), only when lr_shrunk
  uint8_t pad_[6];
};
static_assert(sizeof(WarpScratch) % 16 == 0, "WarpScratch must stay 16-byte granular");

struct Ctx {
  GameRec* g;
  const Topo* T;
  WarpScratch* ws;
  uint8_t* obs;              // staging row, CATAN_OBS_STRIDE bytes
  uint8_t* mask;             // staging row, CATAN_MASK_STRIDE bytes
  uint8_t* scratch;          // >= CATAN_LP_SCRATCH_BYTES, 16-byte aligned; may alias obs+mask (dead during the transition)
  const catan_config_t* cfg;
  uint64_t seed, env_id;
  int lane;
#ifdef CATAN_PROFILE_PHASES
  long long prof_t;
  unsigned long long* prof;  // [phase][4] = {sum cycles, max cycles, count, -}
#endif
};

// phase timers of the profiling build (profiles/phase_profile.py); compiled out of the product
enum { PH_LOAD = 0, PH_SCALAR, PH_DICE, PH_EST, PH_LROAD, PH_FINISH, PH_MASKS, PH_OBS, PH_SAMPLE, PH_STORE, PH_COUNT };
#ifdef CATAN_PROFILE_PHASES
__device__ __forceinline__ void prof_mark(Ctx& cx, int ph) {
  if (cx.lane == 0) {
    const long long t = clock64();
    const unsigned long long d = static_cast<unsigned long long>(t - cx.prof_t);
    atomicAdd(&cx.prof[ph * 4 + 0], d);
    atomicMax(&cx.prof[ph * 4 + 1], d);
    atomicAdd(&cx.prof[ph * 4 + 2], 1ull);
    cx.prof_t = clock64();
  }
}
#define CATAN_PROF(cx, ph) prof_mark(cx, ph)
#else
#define CATAN_PROF(cx, ph) ((void)0)
#endif

// ------------------------------------------------------------------------------------------------
// warp primitives
// ------------------------------------------------------------------------------------------------
#if CATAN_LANES == 32
CATAN_FN void wsync() { __syncwarp(); }
CATAN_FN bool wany(bool p) { return __any_sync(0xffffffffu, p) != 0; }
CATAN_FN int wmax(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { int t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
  return v;
}
CATAN_FN void sadd_i32(int32_t* p, int v) { atomicAdd(p, v); }
// byte add in shared memory: bytes never overflow here (small counters), so a word atomic is exact
CATAN_FN void sadd_u8(uint8_t* p, int v) {
  uintptr_t a = reinterpret_cast<uintptr_t>(p);
  atomicAdd(reinterpret_cast<unsigned int*>(a & ~uintptr_t(3)), static_cast<unsigned int>(v) << (8 * (a & 3)));
}
#else
CATAN_FN void wsync() {}
CATAN_FN bool wany(bool p) { return p; }
CATAN_FN int wmax(int v) { return v; }
CATAN_FN void sadd_i32(int32_t* p, int v) { *p += v; }
CATAN_FN void sadd_u8(uint8_t* p, int v) { *p = static_cast<uint8_t>(*p + v); }
#endif
#define CATAN_LANE_LOOP(i, n) for (int i = cx.lane; i < (n); i += CATAN_LANES)
#define CATAN_PER_LANE(n) (((n) + CATAN_LANES - 1) / CATAN_LANES)

// ------------------------------------------------------------------------------------------------
// pinned RNG (catan_layout.h): Philox4x32-10
// ------------------------------------------------------------------------------------------------
#if CATAN_LANES == 32
CATAN_FN uint32_t mulhi32(uint32_t a, uint32_t b) { return __umulhi(a, b); }
#else
CATAN_FN uint32_t mulhi32(uint32_t a, uint32_t b) { return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32); }
#endif

CATAN_FN_NOINLINE void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  CATAN_NO_UNROLL
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

CATAN_FN uint32_t rng_next(Ctx& cx) {   // [L0] next word of the game stream
  uint32_t d = cx.g->rng_ctr++;
  uint32_t w[4];
  philox4x32(d >> 2, CATAN_STREAM_GAME, static_cast<uint32_t>(cx.env_id), static_cast<uint32_t>(cx.env_id >> 32),
             static_cast<uint32_t>(cx.seed), static_cast<uint32_t>(cx.seed >> 32), w);
  return w[d & 3];
}
CATAN_FN int rng_bounded(Ctx& cx, int n) { return static_cast<int>(mulhi32(rng_next(cx), static_cast<uint32_t>(n))); }
CATAN_FN_NOINLINE void rng_shuffle(Ctx& cx, uint8_t* a, int n) {   // [L0]
  for (int i = n - 1; i >= 1; --i) {
    int j = rng_bounded(cx, i + 1);
    uint8_t t = a[i]; a[i] = a[j]; a[j] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
CATAN_FN int clipi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
CATAN_FN int hand_total(const GameRec& g, int pid) {
  const uint8_t* h = g.res[pid - 1];
  return h[0] + h[1] + h[2] + h[3] + h[4];
}
CATAN_FN void compute_seats(Ctx& cx) {   // [L0] seat of every PlayerId (player.py:13-19)
  for (int i = 0; i < 4; ++i) cx.ws->seat[cx.g->player_order[i]] = static_cast<uint8_t>(i);
}
// relative label of b seen from a: 0 next, 1 next_next, 2 next_next_next (player_lookup); -1 if a == b
CATAN_FN int label_of(const Ctx& cx, int a, int b) { return ((cx.ws->seat[b] - cx.ws->seat[a] + 4) & 3) - 1; }
CATAN_FN int pid_at_label(const Ctx& cx, int a, int label) { return cx.g->player_order[(cx.ws->seat[a] + 1 + label) & 3]; }
CATAN_FN int current_actor(const GameRec& g) {   // game_manager.py:152-159 / wrapper.py:53-58
  return g.need_discard ? g.discard_queue[0] : (g.must_respond ? g.trade_target : g.players_go);
}
CATAN_FN int best_exchange_rate(const GameRec& g, int pid, int r) {   // wrapper.py:428-438
  int h = g.harbours[pid - 1];
  return (h >> (r + 1)) & 1 ? 2 : ((h & 1) ? 3 : 4);
}
CATAN_FN int count_cards(const uint8_t* list, int n, int card) {
  int k = 0;
  CATAN_NO_UNROLL
  for (int i = 0; i < n; ++i) k += list[i] == card;
  return k;
}

// ------------------------------------------------------------------------------------------------
// placement predicates (corner.py:24-39, edge.py:23-42)  [P]
// ------------------------------------------------------------------------------------------------
CATAN_FN_NOINLINE bool can_place_settlement(const GameRec& g, const Topo& T, int c, int pid, bool initial) {
  if (g.corner[c]) return false;
  bool own_road = false;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int nb = T.corner_neigh[c][k];
    if (nb < 0) continue;
    if (g.corner[nb]) return false;
    own_road |= g.edge[T.corner_neigh_edge[c][k]] == pid;
  }
  return initial || own_road;
}

CATAN_FN_NOINLINE bool can_place_road(const GameRec& g, const Topo& T, int e, int pid, bool after_second, int second_corner) {
  if (g.edge[e]) return false;
  int c1 = T.edge_corners[e][0], c2 = T.edge_corners[e][1];
  if (after_second) return c1 == second_corner || c2 == second_corner;
  uint8_t b1 = g.corner[c1], b2 = g.corner[c2];
  if ((b1 && (b1 >> 2) == pid) || (b2 && (b2 >> 2) == pid)) return true;
  bool ok = false;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int e1 = T.corner_neigh_edge[c1][k], e2 = T.corner_neigh_edge[c2][k];
    ok |= (e1 >= 0 && g.edge[e1] == pid && !b1);
    ok |= (e2 >= 0 && g.edge[e2] == pid && !b2);
  }
  return ok;
}

// ------------------------------------------------------------------------------------------------
// reset: Board.reset (board.py:67-167) + Game.reset (game.py:39-136) + EnvWrapper.reset (wrapper.py:30-34)  [L0]
// ------------------------------------------------------------------------------------------------
CATAN_FN_NOINLINE bool number_order_ok(const Topo& T, const uint8_t* numbers, const uint8_t* terrain) {   // board.py:50-65
  uint8_t vals[19];
  int n = 0;
  for (int i = 0; i < 19; ++i) {
    int t = T.number_placement[i];
    vals[t] = terrain[t] == 0 ? 7 : numbers[n++];
  }
  for (int i = 0; i < 19; ++i) {
    if (vals[i] != 6 && vals[i] != 8) continue;
    for (int k = 0; k < 6; ++k) {
      int nb = T.tile_neigh[i][k];
      if (nb >= 0 && (vals[nb] == 6 || vals[nb] == 8)) return false;
    }
  }
  return true;
}

CATAN_FN_NOINLINE void reset_game(Ctx& cx) {   // [L0]
  GameRec& g = *cx.g;
  const Topo& T = *cx.T;
  const uint32_t rng = g.rng_ctr, dec = g.decision_ctr;
  memset(&g, 0, sizeof(GameRec));
  g.rng_ctr = rng;
  g.decision_ctr = dec;
  uint8_t numbers[18];
  for (int i = 0; i < 19; ++i) g.tile_res[i] = static_cast<uint8_t>(T.terrain_to_place[i]);
  rng_shuffle(cx, g.tile_res, 19);                                   // board.py:71-72
  for (int i = 0; i < 18; ++i) numbers[i] = static_cast<uint8_t>(T.default_number_order[i]);
  rng_shuffle(cx, numbers, 18);                                      // board.py:79
  while (!number_order_ok(T, numbers, g.tile_res)) rng_shuffle(cx, numbers, 18);   // board.py:80-81
  for (int i = 0; i < 9; ++i) g.harbour_perm[i] = static_cast<uint8_t>(i);
  rng_shuffle(cx, g.harbour_perm, 9);                                // board.py:83-84
  int n = 0;
  for (int i = 0; i < 19; ++i) {                                     // board.py:88-100
    int t = T.number_placement[i];
    if (g.tile_res[t] == 0) { g.tile_val[t] = 7; g.robber_tile = static_cast<uint8_t>(t); }
    else g.tile_val[t] = numbers[n++];
  }
  g.player_order[0] = WHITE; g.player_order[1] = BLUE; g.player_order[2] = ORANGE; g.player_order[3] = RED;
  rng_shuffle(cx, g.player_order, 4);                                // game.py:41-42
  g.players_go = g.player_order[0];
  for (int r = 0; r < 5; ++r) g.bank[r] = 19;                        // game.py:48-54
  for (int p = 0; p < 4; ++p) { g.settlements_left[p] = 5; g.cities_left[p] = 4; g.second_corner[p] = -1; }
  for (int i = 0; i < 25; ++i) g.deck[i] = static_cast<uint8_t>(T.deck_init[i]);
  rng_shuffle(cx, g.deck, 25);                                       // game.py:75-78
  g.deck_n = 25;
  g.initial_phase = 1;
  compute_seats(cx);
}

// ------------------------------------------------------------------------------------------------
// translate (wrapper.py:114-166, :414-486) and validate (game.py:264-525)  [L0]
// ------------------------------------------------------------------------------------------------
CATAN_FN_NOINLINE int translate_action(const Ctx& cx, const int32_t* a, Act& t) {
  const GameRec& g = *cx.g;
  memset(&t, 0, sizeof(Act));
  const int type = a[CATAN_A_TYPE], pg = g.players_go;
  t.type = static_cast<int8_t>(type);
  switch (type) {
    case CATAN_ACT_PLACE_SETTLEMENT:
    case CATAN_ACT_UPGRADE_CITY:
      if (a[CATAN_A_CORNER] < 0 || a[CATAN_A_CORNER] >= 54) return CATAN_ERR_BAD_HEAD_VALUE;
      t.corner = static_cast<int8_t>(a[CATAN_A_CORNER]);
      return 0;
    case CATAN_ACT_PLACE_ROAD:
      if (a[CATAN_A_EDGE] < 0 || a[CATAN_A_EDGE] > 72) return CATAN_ERR_BAD_HEAD_VALUE;
      t.edge = static_cast<int8_t>(a[CATAN_A_EDGE] == 72 ? -1 : a[CATAN_A_EDGE]);
      return 0;
    case CATAN_ACT_MOVE_ROBBER:
      if (a[CATAN_A_TILE] < 0 || a[CATAN_A_TILE] >= 19) return CATAN_ERR_BAD_HEAD_VALUE;
      t.tile = static_cast<int8_t>(a[CATAN_A_TILE]);
      return 0;
    case CATAN_ACT_STEAL:
    case CATAN_ACT_PROPOSE_TRADE: {
      if (a[CATAN_A_PLAYER] < 0 || a[CATAN_A_PLAYER] > 2) return CATAN_ERR_BAD_HEAD_VALUE;
      t.target_pid = static_cast<int8_t>(pid_at_label(cx, pg, a[CATAN_A_PLAYER]));
      if (type == CATAN_ACT_STEAL) return 0;
      for (int k = 0; k < 4; ++k) {                                  // wrapper.py:451-466: 0 ends the list
        int v = a[CATAN_A_GIVE + k];
        if (v == 0) break;
        if (v < 0 || v > 5) return CATAN_ERR_BAD_HEAD_VALUE;
        t.give[t.n_give++] = static_cast<int8_t>(v - 1);
      }
      for (int k = 0; k < 4; ++k) {
        int v = a[CATAN_A_RECV + k];
        if (v == 0) break;
        if (v < 0 || v > 5) return CATAN_ERR_BAD_HEAD_VALUE;
        t.recv[t.n_recv++] = static_cast<int8_t>(v - 1);
      }
      return 0;
    }
    case CATAN_ACT_PLAY_DEV: {
      int card = a[CATAN_A_CARD];
      if (card < 0 || card > 4) return CATAN_ERR_BAD_HEAD_VALUE;
      t.card = static_cast<int8_t>(card);
      if (card == CATAN_DEV_MONOPOLY || card == CATAN_DEV_YOP) {
        if (a[CATAN_A_RES_A] < 0 || a[CATAN_A_RES_A] > 4) return CATAN_ERR_BAD_HEAD_VALUE;
        t.res_a = static_cast<int8_t>(a[CATAN_A_RES_A]);
      }
      if (card == CATAN_DEV_YOP) {
        if (a[CATAN_A_RES_B] < 0 || a[CATAN_A_RES_B] > 4) return CATAN_ERR_BAD_HEAD_VALUE;
        t.res_b = static_cast<int8_t>(a[CATAN_A_RES_B]);
      }
      return 0;
    }
    case CATAN_ACT_EXCHANGE:
      if (a[CATAN_A_RES_A] < 0 || a[CATAN_A_RES_A] > 4 || a[CATAN_A_RES_B] < 0 || a[CATAN_A_RES_B] > 4)
        return CATAN_ERR_BAD_HEAD_VALUE;
      t.res_a = static_cast<int8_t>(a[CATAN_A_RES_A]);
      t.res_b = static_cast<int8_t>(a[CATAN_A_RES_B]);
      t.rate = static_cast<int8_t>(best_exchange_rate(g, pg, t.res_a));
      return 0;
    case CATAN_ACT_RESPOND:
      if (a[CATAN_A_ACCEPT] < 0 || a[CATAN_A_ACCEPT] > 1) return CATAN_ERR_BAD_HEAD_VALUE;
      t.accept = static_cast<int8_t>(a[CATAN_A_ACCEPT]);
      return 0;
    case CATAN_ACT_DISCARD:
      if (a[CATAN_A_DISCARD] < 0 || a[CATAN_A_DISCARD] > 4) return CATAN_ERR_BAD_HEAD_VALUE;
      t.discard = static_cast<int8_t>(a[CATAN_A_DISCARD]);
      return 0;
    case CATAN_ACT_BUY_DEV:
    case CATAN_ACT_ROLL_DICE:
    case CATAN_ACT_END_TURN:
      return 0;
    default:
      return CATAN_ERR_BAD_TYPE;
  }
}

CATAN_FN_NOINLINE int validate_action(const Ctx& cx, const Act& t) {
  const GameRec& g = *cx.g;
  const Topo& T = *cx.T;
  const int pid = g.players_go, p = pid - 1;
  const uint8_t* h = g.res[p];
  if (g.need_discard) {                                              // game.py:279-300
    if (t.type != CATAN_ACT_DISCARD) return CATAN_ERR_PHASE;
    const int d = g.discard_queue[0];
    if (hand_total(g, d) <= 7) return CATAN_ERR_PHASE;
    return g.res[d - 1][t.discard] > 0 ? 0 : CATAN_ERR_BAD_RESOURCE;
  }
  if (t.type == CATAN_ACT_DISCARD) return CATAN_ERR_PHASE;           // game.py:301-303
  // the common guard of most main-phase actions (must have rolled, nothing pending)
  const bool blocked_main = g.must_respond || g.initial_phase || !g.dice_rolled || g.must_use_dev || g.just_moved_robber;
  switch (t.type) {
    case CATAN_ACT_PLACE_SETTLEMENT:                                 // game.py:305-323
      if (g.must_respond || (!g.dice_rolled && !g.initial_phase) || g.must_use_dev || g.just_moved_robber) return CATAN_ERR_PHASE;
      if (g.initial_phase || (g.settlements_left[p] > 0 && h[WHEAT] && h[WOOD] && h[BRICK] && h[SHEEP])) {
        if (can_place_settlement(g, T, t.corner, pid, g.initial_phase)) {
          if (!g.initial_phase) return 0;
          return (g.init_settlements[p] == 0 || (g.init_settlements[p] == 1 && g.init_roads[p] == 1)) ? 0 : CATAN_ERR_BAD_LOCATION;
        }
      }
      return CATAN_ERR_CANNOT_AFFORD;
    case CATAN_ACT_PLACE_ROAD:                                       // game.py:324-357
      if (g.rb_active) {
        if (t.edge < 0) return 0;
        return can_place_road(g, T, t.edge, pid, false, 0) ? 0 : CATAN_ERR_BAD_LOCATION;
      }
      if (g.must_respond || (!g.dice_rolled && !g.initial_phase) || g.must_use_dev || g.just_moved_robber) return CATAN_ERR_PHASE;
      if (!(g.initial_phase || (h[WOOD] && h[BRICK]))) return CATAN_ERR_CANNOT_AFFORD;
      if (t.edge < 0) return CATAN_ERR_BAD_LOCATION;
      if (!can_place_road(g, T, t.edge, pid, false, 0)) return CATAN_ERR_BAD_LOCATION;
      if (!g.initial_phase) return 0;
      if (g.init_settlements[p] == 1 && g.init_roads[p] == 0) return 0;
      if (g.init_settlements[p] == 2 && g.init_roads[p] == 1)
        return can_place_road(g, T, t.edge, pid, true, g.second_corner[p]) ? 0 : CATAN_ERR_BAD_LOCATION;
      return CATAN_ERR_BAD_LOCATION;
    case CATAN_ACT_UPGRADE_CITY:                                     // game.py:358-376
      if (blocked_main) return CATAN_ERR_PHASE;
      if (g.cities_left[p] > 0 && h[WHEAT] > 1 && h[ORE] > 2) {
        const uint8_t b = g.corner[t.corner];
        if ((b & 3) != 1) return CATAN_ERR_BAD_LOCATION;
        if ((b >> 2) == pid) return 0;
      }
      return CATAN_ERR_CANNOT_AFFORD;
    case CATAN_ACT_BUY_DEV:                                          // game.py:377-393
      if (blocked_main) return CATAN_ERR_PHASE;
      if (h[WHEAT] && h[SHEEP] && h[ORE]) return g.deck_n > 0 ? 0 : CATAN_ERR_BAD_CARD;
      return CATAN_ERR_CANNOT_AFFORD;
    case CATAN_ACT_PLAY_DEV: {                                       // game.py:394-415
      if (g.must_respond || g.played_dev || g.initial_phase || g.just_moved_robber) return CATAN_ERR_PHASE;
      const int k = count_cards(g.hidden[p], g.n_hidden[p], t.card);
      return (k > 0 && k != g.bought[t.card]) ? 0 : CATAN_ERR_BAD_CARD;
    }
    case CATAN_ACT_EXCHANGE:                                         // game.py:416-443
      if (blocked_main) return CATAN_ERR_PHASE;
      if (h[t.res_a] < t.rate) return CATAN_ERR_CANNOT_AFFORD;
      return g.bank[t.res_b] > 0 ? 0 : CATAN_ERR_BAD_RESOURCE;
    case CATAN_ACT_PROPOSE_TRADE: {                                  // game.py:444-466
      if (blocked_main) return CATAN_ERR_PHASE;
      int cnt[5] = {0, 0, 0, 0, 0};
      for (int k = 0; k < t.n_give; ++k) cnt[t.give[k]]++;
      for (int r = 0; r < 5; ++r) if (h[r] < cnt[r]) return CATAN_ERR_CANNOT_AFFORD;
      return 0;
    }
    case CATAN_ACT_RESPOND: {                                        // game.py:467-482
      if (!g.must_respond) return CATAN_ERR_PHASE;
      if (t.accept == 1) return 0;                                   // head value 1 == "reject" (wrapper.py:157-160)
      int cnt[5] = {0, 0, 0, 0, 0};
      for (int k = 0; k < g.n_recv; ++k) cnt[g.recv[k] - 1]++;
      for (int r = 0; r < 5; ++r) if (g.res[g.trade_target - 1][r] < cnt[r]) return CATAN_ERR_CANNOT_AFFORD;
      return 0;
    }
    case CATAN_ACT_MOVE_ROBBER:                                      // game.py:483-490
      return (g.must_respond || g.must_use_dev || !g.can_move_robber) ? CATAN_ERR_PHASE : 0;
    case CATAN_ACT_ROLL_DICE:                                        // game.py:491-500
      return (g.must_respond || g.initial_phase || g.dice_rolled || g.just_moved_robber) ? CATAN_ERR_PHASE : 0;
    case CATAN_ACT_END_TURN:                                         // game.py:501-512
      return blocked_main ? CATAN_ERR_PHASE : 0;
    case CATAN_ACT_STEAL:                                            // game.py:513-525
      if (g.must_respond || !g.just_moved_robber) return CATAN_ERR_PHASE;
      for (int k = 0; k < 6; ++k) {
        const uint8_t b = g.corner[T.tile_corners[g.robber_tile][k]];
        if (b && (b >> 2) == t.target_pid) return 0;
      }
      return CATAN_ERR_BAD_TARGET;
  }
  return CATAN_ERR_BAD_TYPE;
}

// ------------------------------------------------------------------------------------------------
// apply_action, scalar part (game.py:527-815)  [L0].  Belief updates, the dice payout and the
// longest-road search are posted to the warp scratch and executed by all lanes afterwards.
// ------------------------------------------------------------------------------------------------
CATAN_FN void pay(GameRec& g, int p, int r, int n) {   // hand -n, visible floor 0, bank +n (game.py:197-208 etc.)
  g.res[p][r] = static_cast<uint8_t>(g.res[p][r] - n);
  int v = g.vis[p][r] - n;
  g.vis[p][r] = static_cast<int16_t>(v > 0 ? v : 0);
  g.bank[r] = static_cast<uint8_t>(g.bank[r] + n);
}

CATAN_FN EstReq& post_est(Ctx& cx, int owner, int thief) {
  EstReq& q = cx.ws->est[cx.ws->n_est++];
  memset(&q, 0, sizeof(EstReq));
  q.owner = static_cast<uint8_t>(owner);
  q.thief = static_cast<uint8_t>(thief);
  q.T_o = static_cast<int16_t>(hand_total(*cx.g, owner));
  q.T_t = static_cast<int16_t>(thief ? hand_total(*cx.g, thief) : 0);
  return q;
}
CATAN_FN void est_set(EstReq& q, int r, int d) { q.delta[r] = static_cast<int8_t>(d); q.touched |= static_cast<uint8_t>(1u << r); }

CATAN_FN void advance_seat(GameRec& g, bool left) {   // game.py:253-262
  g.player_order_id = static_cast<uint8_t>((g.player_order_id + (left ? 3 : 1)) & 3);
  g.players_go = g.player_order[g.player_order_id];
}

CATAN_FN_NOINLINE void update_largest_army(GameRec& g) {   // game.py:817-841  [L0]
  const int order[4] = {BLUE, WHITE, RED, ORANGE};
  int max_count = 0, cp = 0;
  for (int i = 0; i < 4; ++i) {
    const int q = order[i];
    const int k = count_cards(g.played[q - 1], g.n_played[q - 1], CATAN_DEV_KNIGHT);
    g.cur_army[q - 1] = static_cast<uint8_t>(k);
    if (k >= 3 && k > max_count) { max_count = k; cp = q; }
  }
  if (!cp) return;
  if (!g.la_holder) { g.la_holder = static_cast<uint8_t>(cp); g.la_count = static_cast<uint8_t>(max_count); g.vp[cp - 1] += 2; }
  else if (g.la_holder == cp) g.la_count = static_cast<uint8_t>(max_count);
  else if (max_count > g.la_count) {
    g.vp[g.la_holder - 1] -= 2;
    g.la_holder = static_cast<uint8_t>(cp); g.la_count = static_cast<uint8_t>(max_count);
    g.vp[cp - 1] += 2;
  }
}

CATAN_FN_NOINLINE void apply_scalar(Ctx& cx) {   // [L0]
  GameRec& g = *cx.g;
  const Topo& T = *cx.T;
  WarpScratch& ws = *cx.ws;
  const Act& t = ws.act;
  const int pid = g.players_go, p = pid - 1;
  switch (t.type) {
    case CATAN_ACT_PLACE_SETTLEMENT: {                               // game.py:530-555, :195-212
      const int c = t.corner;
      if (!g.initial_phase) { pay(g, p, WHEAT, 1); pay(g, p, SHEEP, 1); pay(g, p, WOOD, 1); pay(g, p, BRICK, 1); }
      g.corner[c] = static_cast<uint8_t>((pid << 2) | 1);
      const int slot = T.corner_harbour_slot[c];                     // board.py:182-183
      if (slot >= 0) {
        const int hres = T.harbour_res[g.harbour_perm[slot]];
        g.harbours[p] |= static_cast<uint8_t>(1u << hres);           // bit 0 = generic, bit Resource = 2:1
      }
      g.settlements_left[p] -= 1;
      g.vp[p] += 1;
      if (g.initial_phase) {
        g.init_settlements[p] += 1;
        if (g.init_settlements[p] == 2) {
          int gain[5] = {0, 0, 0, 0, 0};
          for (int k = 0; k < 3; ++k) {
            const int tl = T.corner_tiles[c][k];
            if (tl < 0 || g.tile_res[tl] == 0) continue;
            const int r = g.tile_res[tl] - 1;
            g.res[p][r] += 1; g.vis[p][r] += 1; g.bank[r] -= 1; gain[r] += 1;
          }
          EstReq& q = post_est(cx, pid, 0);
          for (int r = 0; r < 5; ++r) if (gain[r]) est_set(q, r, gain[r]);
          g.second_corner[p] = static_cast<int8_t>(c);
        }
      } else {
        EstReq& q = post_est(cx, pid, 0);
        est_set(q, BRICK, -1); est_set(q, WOOD, -1); est_set(q, WHEAT, -1); est_set(q, SHEEP, -1);
        if (g.lr_holder) ws.lr_pid = g.lr_holder;                    // game.py:552-553
      }
      break;
    }
    case CATAN_ACT_PLACE_ROAD: {                                     // game.py:556-597, :222-232
      bool final_init = false;
      if (t.edge >= 0) {
        if (!g.initial_phase && !g.rb_active) { pay(g, p, WOOD, 1); pay(g, p, BRICK, 1); }
        g.edge[t.edge] = static_cast<uint8_t>(pid);
        if (g.initial_phase) {
          g.init_roads[p] += 1;
          int first = 0, second = 0;
          for (int q = 0; q < 4; ++q) { first += g.init_settlements[q] >= 1; second += g.init_settlements[q] == 2; }
          if (first < 4) advance_seat(g, false);
          else if (second == 0) { /* last seat places twice in a row */ }
          else if (second < 4) advance_seat(g, true);
          else { g.initial_phase = 0; final_init = true; }
        }
      }
      ws.lr_pid = static_cast<uint8_t>(pid);                         // game.py:585 (also for the dummy edge)
      if (g.rb_active) {
        g.rb_count += 1;
        if (g.rb_count >= 2) { g.rb_active = 0; g.rb_count = 0; g.must_use_dev = 0; }
      } else if (!g.initial_phase && !final_init) {
        EstReq& q = post_est(cx, pid, 0);
        est_set(q, BRICK, -1); est_set(q, WOOD, -1);
      }
      break;
    }
    case CATAN_ACT_UPGRADE_CITY: {                                   // game.py:598-604, :240-251
      pay(g, p, WHEAT, 2); pay(g, p, ORE, 3);
      g.corner[t.corner] = static_cast<uint8_t>((pid << 2) | 2);
      g.vp[p] += 1; g.cities_left[p] -= 1; g.settlements_left[p] += 1;
      EstReq& q = post_est(cx, pid, 0);
      est_set(q, ORE, -3); est_set(q, WHEAT, -2);
      break;
    }
    case CATAN_ACT_ROLL_DICE: {                                      // game.py:605-611, :138-150
      g.die1 = static_cast<uint8_t>(1 + rng_bounded(cx, 6));
      g.die2 = static_cast<uint8_t>(1 + rng_bounded(cx, 6));
      const int roll = g.die1 + g.die2;
      ws.roll_info = static_cast<uint8_t>(roll);
      g.dice_rolled = 1;
      if (roll == 7) {
        g.can_move_robber = 1;
        for (int i = 0; i < 4; ++i) {
          const int q = g.player_order[i];
          if (hand_total(g, q) > 7) { g.need_discard = 1; g.discard_queue[g.n_discard++] = static_cast<uint8_t>(q); }
        }
      } else {
        ws.dice_roll = static_cast<uint8_t>(roll);                   // payout + beliefs: dice_payout() [W]
      }
      break;
    }
    case CATAN_ACT_END_TURN:                                         // game.py:612-622
      g.can_move_robber = 0; g.dice_rolled = 0; g.played_dev = 0;
      advance_seat(g, false);
      g.turn += 1;
      for (int c = 0; c < 5; ++c) g.bought[c] = 0;
      g.trades_this_turn = 0; g.actions_this_turn = 0;
      break;
    case CATAN_ACT_MOVE_ROBBER: {                                    // game.py:623-634
      g.robber_tile = static_cast<uint8_t>(t.tile);
      g.can_move_robber = 0;
      for (int k = 0; k < 6; ++k) {
        const uint8_t b = g.corner[T.tile_corners[t.tile][k]];
        if (b && (b >> 2) != pid) g.just_moved_robber = 1;
      }
      break;
    }
    case CATAN_ACT_STEAL: {                                          // game.py:635-652
      const int v = t.target_pid;
      const int n = hand_total(g, v);
      if (n > 0) {
        const int order[5] = {BRICK, WHEAT, WOOD, SHEEP, ORE};       // game.py:638
        int idx = rng_bounded(cx, n), r = BRICK;
        for (int i = 0; i < 5; ++i) {
          const int cnt = g.res[v - 1][order[i]];
          if (idx < cnt) { r = order[i]; break; }
          idx -= cnt;
        }
        g.res[p][r] += 1; g.res[v - 1][r] -= 1;
        for (int q = 0; q < 5; ++q) if (g.vis[v - 1][q] > 0) g.vis[v - 1][q] -= 1;
        EstReq& rq = post_est(cx, v, pid);
        est_set(rq, r, -1);
      }
      g.just_moved_robber = 0;
      break;
    }
    case CATAN_ACT_PLAY_DEV: {                                       // game.py:653-693
      const int n = g.n_hidden[p];
      int at = 0;
      while (at < n && g.hidden[p][at] != t.card) ++at;
      if (at < n) {                                                  // (always true for a validated action)
        for (int i = at; i + 1 < n; ++i) g.hidden[p][i] = g.hidden[p][i + 1];
        g.hidden[p][n - 1] = 0;
        g.n_hidden[p] -= 1;
      }
      if (g.n_played[p] < 25) g.played[p][g.n_played[p]++] = static_cast<uint8_t>(t.card);
      g.played_dev = 1;
      if (t.card == CATAN_DEV_VP) g.vp[p] += 1;
      else if (t.card == CATAN_DEV_KNIGHT) { g.can_move_robber = 1; update_largest_army(g); }
      else if (t.card == CATAN_DEV_ROADBUILDING) { g.rb_active = 1; g.rb_count = 0; g.must_use_dev = 1; }
      else if (t.card == CATAN_DEV_MONOPOLY) {
        const int r = t.res_a;
        ws.est_special = EST_SPECIAL_MONOPOLY;
        ws.mono_pid = static_cast<uint8_t>(pid); ws.mono_res = static_cast<uint8_t>(r);
        for (int o = 0; o < 4; ++o) {
          ws.mono_lost[o] = 0;
          if (o == p) continue;
          const int cnt = g.res[o][r];
          g.res[o][r] = 0; g.vis[o][r] = 0;
          g.res[p][r] = static_cast<uint8_t>(g.res[p][r] + cnt); g.vis[p][r] = static_cast<int16_t>(g.vis[p][r] + cnt);
          ws.mono_lost[o] = static_cast<uint8_t>(cnt);
        }
        for (int o = 0; o < 4; ++o) ws.mono_T[o] = static_cast<int16_t>(hand_total(g, o + 1));
      } else {                                                       // Year of Plenty
        const int rr[2] = {t.res_a, t.res_b};
        for (int i = 0; i < 2; ++i) {
          const int r = rr[i];
          if (g.bank[r] > 0) {
            g.bank[r] -= 1; g.res[p][r] += 1; g.vis[p][r] += 1;
            EstReq& q = post_est(cx, pid, 0);
            est_set(q, r, 1);
          }
        }
      }
      break;
    }
    case CATAN_ACT_BUY_DEV: {                                        // game.py:694-710
      pay(g, p, SHEEP, 1); pay(g, p, ORE, 1); pay(g, p, WHEAT, 1);
      EstReq& q = post_est(cx, pid, 0);
      est_set(q, SHEEP, -1); est_set(q, ORE, -1); est_set(q, WHEAT, -1);
      if (g.deck_n > 0 && g.n_hidden[p] < 25) {                      // (always true for a validated action)
        const int card = g.deck[g.deck_n - 1];                       // deque.pop(): right end
        g.deck[g.deck_n - 1] = 0; g.deck_n -= 1;
        g.hidden[p][g.n_hidden[p]++] = static_cast<uint8_t>(card);
        g.bought[card] += 1;
      }
      break;
    }
    case CATAN_ACT_EXCHANGE: {                                       // game.py:711-734
      const int d = t.res_b, tr = t.res_a, rate = t.rate;
      g.res[p][d] += 1; g.vis[p][d] += 1;
      g.res[p][tr] = static_cast<uint8_t>(g.res[p][tr] - rate);
      const int v = g.vis[p][tr] - rate;
      g.vis[p][tr] = static_cast<int16_t>(v > 0 ? v : 0);
      g.bank[tr] = static_cast<uint8_t>(g.bank[tr] + rate); g.bank[d] -= 1;
      EstReq& q = post_est(cx, pid, 0);
      if (d == tr) est_set(q, d, 1 - rate);
      else { est_set(q, d, 1); est_set(q, tr, -rate); }
      break;
    }
    case CATAN_ACT_PROPOSE_TRADE:                                    // game.py:735-750
      g.must_respond = 1;
      g.trade_proposer = static_cast<uint8_t>(pid); g.trade_target = static_cast<uint8_t>(t.target_pid);
      g.n_give = static_cast<uint8_t>(t.n_give); g.n_recv = static_cast<uint8_t>(t.n_recv);
      for (int k = 0; k < 4; ++k) {
        g.give[k] = static_cast<uint8_t>(k < t.n_give ? t.give[k] + 1 : 0);
        g.recv[k] = static_cast<uint8_t>(k < t.n_recv ? t.recv[k] + 1 : 0);
      }
      g.trades_this_turn += 1;
      break;
    case CATAN_ACT_RESPOND: {                                        // game.py:751-784
      if (t.accept == 0) {
        const int p1 = g.trade_proposer - 1, p2 = g.trade_target - 1;
        int d1[5] = {0, 0, 0, 0, 0};
        uint8_t touched = 0;
        for (int k = 0; k < g.n_give; ++k) {
          const int r = g.give[k] - 1;
          g.res[p1][r] -= 1; if (g.vis[p1][r] > 0) g.vis[p1][r] -= 1;
          g.res[p2][r] += 1; g.vis[p2][r] += 1;
          d1[r] -= 1; touched |= static_cast<uint8_t>(1u << r);
        }
        for (int k = 0; k < g.n_recv; ++k) {
          const int r = g.recv[k] - 1;
          g.res[p1][r] += 1; g.vis[p1][r] += 1;
          g.res[p2][r] -= 1; if (g.vis[p2][r] > 0) g.vis[p2][r] -= 1;
          d1[r] += 1; touched |= static_cast<uint8_t>(1u << r);
        }
        EstReq& q1 = post_est(cx, p1 + 1, 0);
        EstReq& q2 = post_est(cx, p2 + 1, 0);
        for (int r = 0; r < 5; ++r) { q1.delta[r] = static_cast<int8_t>(d1[r]); q2.delta[r] = static_cast<int8_t>(-d1[r]); }
        q1.touched = touched; q2.touched = touched;
      }
      g.must_respond = 0;
      g.trade_proposer = 0; g.trade_target = 0; g.n_give = 0; g.n_recv = 0;
      for (int k = 0; k < 4; ++k) { g.give[k] = 0; g.recv[k] = 0; }
      break;
    }
    case CATAN_ACT_DISCARD: {                                        // game.py:785-807
      const int d = g.discard_queue[0], r = t.discard;
      g.res[d - 1][r] -= 1; g.bank[r] += 1;
      EstReq& q = post_est(cx, d, 0);
      est_set(q, r, -1);
      if (hand_total(g, d) <= 7) {
        for (int i = 0; i < 3; ++i) g.discard_queue[i] = g.discard_queue[i + 1];
        g.discard_queue[3] = 0;
        g.n_discard -= 1;
        if (g.n_discard == 0) g.need_discard = 0;
      }
      break;
    }
  }
  if (t.type != CATAN_ACT_RESPOND && t.type != CATAN_ACT_END_TURN && t.type != CATAN_ACT_DISCARD)
    g.actions_this_turn += 1;                                        // game.py:809-810
}

// ------------------------------------------------------------------------------------------------
// dice payout (game.py:151-175)  [W]
// ------------------------------------------------------------------------------------------------
CATAN_FN_NOINLINE void dice_payout(Ctx& cx) {
  GameRec& g = *cx.g;
  const Topo& T = *cx.T;
  WarpScratch& ws = *cx.ws;
  const int roll = ws.dice_roll;
  CATAN_LANE_LOOP(i, 20) (&ws.alloc[0][0])[i] = 0;
  wsync();
  CATAN_LANE_LOOP(i, 114) {                                          // (tile, corner) pairs
    const int t = i / 6, k = i - 6 * t;
    if (g.tile_val[t] != roll || t == g.robber_tile) continue;
    const uint8_t b = g.corner[T.tile_corners[t][k]];
    if (b) sadd_i32(&ws.alloc[g.tile_res[t] - 1][(b >> 2) - 1], b & 3);   // settlement +1, city +2
  }
  wsync();
  if (cx.lane == 0) {
    const int res_order[5] = {WOOD, ORE, BRICK, WHEAT, SHEEP};       // game.py:153-155
    int tot[4];
    for (int p = 0; p < 4; ++p) tot[p] = hand_total(g, p + 1);
    uint8_t granted = 0;
    for (int ri = 0; ri < 5; ++ri) {
      const int r = res_order[ri];
      const int total = ws.alloc[r][0] + ws.alloc[r][1] + ws.alloc[r][2] + ws.alloc[r][3];
      if (total > g.bank[r]) continue;                               // all-or-nothing per resource (game.py:171)
      granted |= static_cast<uint8_t>(1u << r);
      for (int p = 0; p < 4; ++p) {
        g.res[p][r] = static_cast<uint8_t>(g.res[p][r] + ws.alloc[r][p]);
        g.bank[r] = static_cast<uint8_t>(g.bank[r] - ws.alloc[r][p]);
        tot[p] += ws.alloc[r][p];
        ws.dice_T[p][r] = static_cast<int16_t>(tot[p]);              // owner's running total when (r, p) is re-clipped (Q4)
      }
    }
    ws.granted = granted;
    ws.est_special = EST_SPECIAL_DICE;
  }
  wsync();
}

// ------------------------------------------------------------------------------------------------
// belief updates, one lane per (observer, label, resource) entry  [W]
//   generic  : update_player_resource_estimates (game.py:921-971)
//   dice     : the 20 calls of one roll folded into one pass (game.py:170-175; Q4)
//   monopoly : update_resource_estimates_monopoly (game.py:973-1010)
// ------------------------------------------------------------------------------------------------
CATAN_FN_NOINLINE void est_apply(Ctx& cx) {
  GameRec& g = *cx.g;
  WarpScratch& ws = *cx.ws;
  int16_t* emin = &g.est_min[0][0][0];
  int16_t* emax = &g.est_max[0][0][0];
  for (int qi = 0; qi < ws.n_est; ++qi) {
    const EstReq rq = ws.est[qi];
    int16_t nmin[CATAN_PER_LANE(60)], nmax[CATAN_PER_LANE(60)];
    bool wr[CATAN_PER_LANE(60)];
#pragma unroll
    for (int it = 0; it < CATAN_PER_LANE(60); ++it) {
      const int i = cx.lane + it * CATAN_LANES;
      wr[it] = false;
      if (i >= 60) continue;
      const int o = i / 15, l = (i / 5) % 3, r = i % 5;
      const int observer = o + 1, target = pid_at_label(cx, observer, l);
      const bool touched = (rq.touched >> r) & 1;
      int mn = emin[i], mx = emax[i];
      if (!rq.thief || observer == rq.thief) {                       // game.py:936-954
        if (target == rq.owner && observer != rq.owner && touched) {
          mx = clipi(mx + rq.delta[r], 0, rq.T_o); mn = clipi(mn + rq.delta[r], 0, rq.T_o); wr[it] = true;
        }
      } else if (observer == rq.owner) {                             // victim knows what was taken (game.py:929-933)
        if (target == rq.thief && touched) { mx -= rq.delta[r]; mn -= rq.delta[r]; wr[it] = true; }
      } else {                                                       // third party (game.py:955-971)
        if (target == rq.owner) {
          mx = clipi(mx, 0, rq.T_o); mn = clipi(mn - 1, 0, rq.T_o); wr[it] = true;
        } else if (target == rq.thief) {
          const int m0 = emax[(o * 3 + label_of(cx, observer, rq.owner)) * 5 + r];   // the victim entry BEFORE its clip
          if (m0 > 0) { mx = clipi(mx + 1, 0, rq.T_t); mn = clipi(mn, 0, rq.T_t); wr[it] = true; }
        }
      }
      nmin[it] = static_cast<int16_t>(mn); nmax[it] = static_cast<int16_t>(mx);
    }
    wsync();
#pragma unroll
    for (int it = 0; it < CATAN_PER_LANE(60); ++it) {
      const int i = cx.lane + it * CATAN_LANES;
      if (i < 60 && wr[it]) { emin[i] = nmin[it]; emax[i] = nmax[it]; }
    }
    wsync();
  }
  if (ws.est_special == EST_SPECIAL_DICE) {
    CATAN_LANE_LOOP(i, 60) {
      const int o = i / 15, l = (i / 5) % 3, r = i % 5;
      if (!((ws.granted >> r) & 1)) continue;
      const int tp = pid_at_label(cx, o + 1, l) - 1;
      const int gain = ws.alloc[r][tp], T = ws.dice_T[tp][r];
      emax[i] = static_cast<int16_t>(clipi(emax[i] + gain, 0, T));
      emin[i] = static_cast<int16_t>(clipi(emin[i] + gain, 0, T));
    }
    wsync();
  } else if (ws.est_special == EST_SPECIAL_MONOPOLY) {
    int tot = 0;
    for (int q = 0; q < 4; ++q) tot += ws.mono_lost[q];
    CATAN_LANE_LOOP(i, 60) {
      const int o = i / 15, l = (i / 5) % 3, r = i % 5;
      const int target = pid_at_label(cx, o + 1, l);
      if (target == ws.mono_pid) {                                   // game.py:984-991, unclipped
        if (r == ws.mono_res) { emin[i] = static_cast<int16_t>(emin[i] + tot); emax[i] = static_cast<int16_t>(emax[i] + tot); }
      } else {                                                       // game.py:993-1010
        const int lost = r == ws.mono_res ? ws.mono_lost[target - 1] : 0, T = ws.mono_T[target - 1];
        emax[i] = static_cast<int16_t>(clipi(emax[i] - lost, 0, T));
        emin[i] = static_cast<int16_t>(clipi(emin[i] - lost, 0, T));
      }
    }
    wsync();
  }
}

// ------------------------------------------------------------------------------------------------
// longest road: node-simple longest path (game.py:843-862, utils.py:3-15; Q7)  [W]
//
// The reference enumerates every simple path of the player's road graph from every start corner.  Here
// the graph is first reduced to one 64-bit adjacency mask per corner (a corner holding an opponent's
// building keeps its incoming arcs but gets no outgoing ones, game.py:851-858).  The enumeration is
// cut into 324 independent work items (start corner, 1st branch, 2nd branch) that the lanes claim from a
// shared counter, and every lane runs the same branch-free push/pop state machine over its own path
// stack, so the warp stays converged while the lanes sit at different depths of different subtrees.
// Scratch (cx.scratch): adj[54] u64 | counter | path[54][LANES] bytes.
// ------------------------------------------------------------------------------------------------
#define CATAN_LP_ADJ_BYTES 432
#define CATAN_LP_PATH_OFF 464
#define CATAN_LP_SCRATCH_BYTES (CATAN_LP_PATH_OFF + 54 * CATAN_LANES)
#define CATAN_LP_ITEMS 324
#if CATAN_LANES == 32
CATAN_FN int ctz64(uint64_t x) { return __ffsll(static_cast<long long>(x)) - 1; }
CATAN_FN int fetch_add_i32(int32_t* p) { return atomicAdd(p, 1); }
CATAN_FN void smax_i32(int32_t* p, int v) { atomicMax(p, v); }
#else
CATAN_FN int ctz64(uint64_t x) { return __builtin_ctzll(x); }
CATAN_FN int fetch_add_i32(int32_t* p) { return (*p)++; }
CATAN_FN void smax_i32(int32_t* p, int v) { if (v > *p) *p = v; }
#endif
CATAN_FN int kth_bit(uint64_t m, int k) {   // index of the k-th (0-based) set bit, -1 if there is none
  for (int j = 0; j < k; ++j) m &= m - 1;
  return m ? ctz64(m) : -1;
}

// adjacency masks of PlayerId pid's road graph: adj[v] = corners reachable from v over one own road;
// a corner holding an opponent's building has no outgoing arcs (game.py:851-858).  [W]; caller syncs.
CATAN_FN void lp_build_adj(const GameRec& g, const Topo& T, int pid, uint64_t* adj, int lane) {
  for (int v = lane; v < 54; v += CATAN_LANES) {
    uint64_t a = 0;
    const uint8_t b = g.corner[v];
    if (!(b && (b >> 2) != pid)) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int e = T.corner_neigh_edge[v][k];
        if (e >= 0 && g.edge[e] == pid) a |= 1ull << T.corner_neigh[v][k];
      }
    }
    adj[v] = a;
  }
}

// Cooperative search over n_jobs graphs (adj[job][54]), in ROUNDS of at most `budget` loop iterations.
// A unit of work is a DFS state: in round 0 the 324 static prefixes (start corner, 1st branch, 2nd branch) per job
// claimed from *counter, in later rounds the tasks queued in `ring`.  Every lane walks its unit with the push/pop
// state machine; the deepest level seen per job is max-ed into best[job].  When the budget of a round is used up,
// every lane that is still inside a subtree (a) donates the untried siblings of the shallowest open level of its
// path as fresh tasks (the biggest pieces of what is left) and (b) parks the rest of its state as a RESUME task;
// the next round hands all of these to whatever lanes are free.  Parallelism therefore grows geometrically round
// by round and a dense road network (10^5 path visits) is spread over the whole block instead of pinning a lane.
// path: byte stacks, element (depth, lane) at path[depth * path_stride + path_lane]; bit 7 of a stack byte marks a
// level whose remaining siblings were donated.  ctl[0] = claim counter, ctl[1] = ring reservation cursor (absolute,
// monotonic), ctl[2] = first absolute slot that is NOT valid this round, ctl[3] = absolute slot of this round's first task.
// Any number of warps may call this concurrently on the same arguments; the caller synchronises between rounds.
struct alignas(8) LpTask {
  uint8_t job, depth, base, pad_[5];
  uint64_t above;            // untried-candidate filter at the deepest level (resume tasks); ~0 for fresh tasks
  uint8_t path[56];          // path[0..depth], flag bits included
};
static_assert(sizeof(LpTask) == 72, "LpTask layout");

CATAN_FN_NOINLINE void lp_round(const uint64_t* adj_all, int n_jobs, bool first_round, int32_t* ctl, int32_t* best,
                                uint8_t* path, int path_stride, int path_lane, LpTask* ring, int ring_cap, int n_in, int budget) {
  const int total = first_round ? n_jobs * CATAN_LP_ITEMS : n_in;
  const int round_base = ctl[3];
  const uint64_t* adj = adj_all;
  int job = 0, lbest = 0, node = 0, depth = 0, base = 0, iters = 0;
  uint64_t visited = 0, above = ~0ull;
  bool active = false, exhausted = false;
  for (;;) {
    if (!active && !exhausted) {                                     // claim the next non-empty unit
      for (;;) {
        const int i = fetch_add_i32(&ctl[0]);
        if (i >= total) { exhausted = true; break; }
        int nj;
        if (first_round) {
          nj = i / CATAN_LP_ITEMS;
          const int it = i - nj * CATAN_LP_ITEMS;
          if (nj != job) { if (lbest) smax_i32(&best[job], lbest); job = nj; lbest = 0; }
          adj = adj_all + nj * 54;
          const int v = it / 6, k1 = (it % 6) >> 1, k2 = it & 1;
          const int t1 = kth_bit(adj[v], k1);
          if (t1 < 0) continue;
          const uint64_t c2 = adj[t1] & ~(1ull << v);
          if (!c2) { if (k2 == 0 && lbest < 1) lbest = 1; continue; }
          const int t2 = kth_bit(c2, k2);
          if (t2 < 0) continue;
          path[path_lane] = static_cast<uint8_t>(v);
          path[path_stride + path_lane] = static_cast<uint8_t>(t1);
          path[2 * path_stride + path_lane] = static_cast<uint8_t>(t2);
          visited = (1ull << v) | (1ull << t1) | (1ull << t2);
          depth = 2; base = 2; node = t2; above = ~0ull;
        } else {
          const LpTask& tk = ring[(round_base + i) % ring_cap];
          nj = tk.job;
          if (nj != job) { if (lbest) smax_i32(&best[job], lbest); job = nj; lbest = 0; }
          adj = adj_all + nj * 54;
          depth = tk.depth; base = tk.base; above = tk.above;
          visited = 0;
          for (int q = 0; q <= depth; ++q) {
            const int raw = tk.path[q];
            path[q * path_stride + path_lane] = static_cast<uint8_t>(raw);
            visited |= 1ull << (raw & 63);
          }
          node = tk.path[depth] & 63;
        }
        active = true;
        if (lbest < depth) lbest = depth;
        break;
      }
    }
    if (!wany(active)) break;
    if (++iters >= budget) {                                         // (warp-uniform) budget used up: queue what is left
      iters = 0;
      if (active) {
        // shallowest open level q in [base, depth] that still has untried siblings -> they become fresh tasks
        int dq = -1;
        uint64_t dc = 0, vis = 0;
        for (int q = 0; q <= depth; ++q) {
          const int raw = path[q * path_stride + path_lane], u = raw & 63;
          vis |= 1ull << u;
          if (q < base || (raw & 128)) continue;
          uint64_t c = adj[u] & ~vis;
          c &= q == depth ? above : ~((2ull << (path[(q + 1) * path_stride + path_lane] & 63)) - 1ull);
          if (c) { dq = q; dc = c; break; }
        }
        int need = 1;                                                // the resume task
        for (uint64_t t = dc; t; t &= t - 1) ++need;
        // Reserve `need` consecutive ring slots.  A reservation that does not fit leaves a hole at the END of the
        // round's list (every later reservation fails too); ctl[2] remembers where the valid part stops.
#if CATAN_LANES == 32
        const int slot = atomicAdd(&ctl[1], need);
#else
        const int slot = ctl[1]; ctl[1] += need;
#endif
        if (slot + need - round_base <= ring_cap) {
          int w = slot;
          for (uint64_t t = dc; t; t &= t - 1) {                     // fresh tasks: prefix path[0..dq] + one untried sibling
            LpTask& tk = ring[w++ % ring_cap];
            tk.job = static_cast<uint8_t>(job);
            tk.depth = static_cast<uint8_t>(dq + 1);
            tk.base = static_cast<uint8_t>(dq + 1);
            tk.above = ~0ull;
            for (int z = 0; z <= dq; ++z) tk.path[z] = path[z * path_stride + path_lane] & 63;
            tk.path[dq + 1] = static_cast<uint8_t>(ctz64(t));
          }
          if (dq >= 0) {
            path[dq * path_stride + path_lane] |= 128;              // those siblings are no longer this unit's business
            if (dq == depth) above = 0;
          }
          LpTask& rk = ring[w % ring_cap];                           // resume task: the state machine's registers + stack
          rk.job = static_cast<uint8_t>(job);
          rk.depth = static_cast<uint8_t>(depth);
          rk.base = static_cast<uint8_t>(base);
          rk.above = above;
          for (int z = 0; z <= depth; ++z) rk.path[z] = path[z * path_stride + path_lane];
          active = false;
        } else {
#if CATAN_LANES == 32
          atomicMin(&ctl[2], slot);                                  // ring full: keep the unit and carry on in this round
#else
          if (slot < ctl[2]) ctl[2] = slot;
#endif
        }
      }
      continue;
    }
    if (active) {
      const uint64_t cand = adj[node] & ~visited & above;
      if (cand) {                                                    // push the lowest untried neighbour
        const int t = ctz64(cand);
        ++depth;
        path[depth * path_stride + path_lane] = static_cast<uint8_t>(t);
        visited |= 1ull << t;
        node = t; above = ~0ull;
        if (depth > lbest) lbest = depth;
      } else if (depth == base) {
        active = false;                                              // unit exhausted
      } else {                                                       // pop; resume the parent above the popped child
        visited &= ~(1ull << node);
        --depth;
        const int raw = path[depth * path_stride + path_lane];
        above = (raw & 128) ? 0ull : ~((2ull << node) - 1ull);       // siblings of a donated level belong to other lanes
        node = raw & 63;
      }
    }
  }
  if (lbest) smax_i32(&best[job], lbest);
}

// Round driver shared by the warp-local and the block-cooperative callers: SYNC_ is the barrier of the participating
// threads, leader_ is true for exactly one of them.  ctl_ holds TWO sets of four control words used by alternate
// rounds, so that the leader can prepare the next round's set while the others still read this round's: two barriers
// per round.  The caller must have made the adjacency tables and best_[] visible (one barrier) before entering.
#define CATAN_LP_RUN(adj_, n_jobs_, ctl_, best_, path_, stride_, plane_, ring_, cap_, budget_, leader_, SYNC_, ROUNDS_)   \
  do {                                                                                                                     \
    int n_in_ = 0, set_ = 0;                                                                                               \
    bool first_ = true;                                                                                                    \
    if (leader_) { (ctl_)[0] = 0; (ctl_)[1] = 0; (ctl_)[2] = 0x7fffffff; (ctl_)[3] = 0; }                                  \
    do {                                                                                                                   \
      SYNC_;                                                                                                               \
      int32_t* c_ = (ctl_) + 4 * set_;                                                                                     \
      lp_round(adj_, n_jobs_, first_, c_, best_, path_, stride_, plane_, ring_, cap_, n_in_, budget_);                     \
      SYNC_;                                                                                                               \
      {                                                                                                                    \
        const int base_ = c_[3] + n_in_;                              /* first task queued during this round */           \
        const int end_ = c_[1] < c_[2] ? c_[1] : c_[2];                                                                    \
        n_in_ = end_ - base_;                                                                                              \
        set_ ^= 1;                                                                                                         \
        if (leader_) { int32_t* d_ = (ctl_) + 4 * set_; d_[0] = 0; d_[1] = end_; d_[2] = 0x7fffffff; d_[3] = base_; }      \
      }                                                                                                                    \
      first_ = false;                                                                                                      \
      ROUNDS_;                                                                                                             \
    } while (n_in_ > 0);                                                                                                   \
  } while (0)

// warp-local longest path of one player (used by the host emulation and by callers without a block)  [W]
// scratch: adj 432 | ctl[8] | path stacks | best | task ring of CATAN_LP_WARP_TASKS
#define CATAN_LP_WARP_TASKS 96
#undef CATAN_LP_SCRATCH_BYTES
#define CATAN_LP_TASK_OFF ((CATAN_LP_PATH_OFF + 54 * CATAN_LANES + 8 + 7) & ~7)
#define CATAN_LP_SCRATCH_BYTES (CATAN_LP_TASK_OFF + CATAN_LP_WARP_TASKS * 72)
#ifndef CATAN_LP_BUDGET
#define CATAN_LP_BUDGET 160
#endif
CATAN_FN int longest_path(Ctx& cx, int pid) {
  uint64_t* adj = reinterpret_cast<uint64_t*>(cx.scratch);
  int32_t* ctl = reinterpret_cast<int32_t*>(cx.scratch + CATAN_LP_ADJ_BYTES);
  int32_t* best = reinterpret_cast<int32_t*>(cx.scratch + CATAN_LP_TASK_OFF - 8);
  LpTask* ring = reinterpret_cast<LpTask*>(cx.scratch + CATAN_LP_TASK_OFF);
  lp_build_adj(*cx.g, *cx.T, pid, adj, cx.lane);
  if (cx.lane == 0) *best = 0;
  wsync();
  CATAN_LP_RUN(adj, 1, ctl, best, cx.scratch + CATAN_LP_PATH_OFF, CATAN_LANES, cx.lane, ring, CATAN_LP_WARP_TASKS, CATAN_LP_BUDGET,
               cx.lane == 0, wsync(), (void)0);
  wsync();
  const int r = *best;
  wsync();
  return r;
}

// game.py:880-881: the holder's own path got shorter -> the other three players must be re-measured
CATAN_FN bool lr_is_shrunk(const GameRec& g, int pid, int len) { return g.lr_holder == pid && g.lr_count > len; }

// game.py:864-919 given the measured lengths: len of `pid`, and (only when shrunk) other_len[PlayerId] of the rest  [L0]
CATAN_FN_NOINLINE void lr_apply(GameRec& g, int pid, int len, bool shrunk, const uint8_t* other_len) {
  const int holder = g.lr_holder, count = g.lr_count;
  g.cur_longest_path[pid - 1] = static_cast<uint8_t>(len);
  g.has_path_key[pid - 1] = 1;
  if (!holder) {
    if (len >= 5) { g.lr_holder = static_cast<uint8_t>(pid); g.lr_count = static_cast<uint8_t>(len); g.vp[pid - 1] += 2; }
  } else if (holder == pid) {
    if (shrunk) {
      int max_len = len, player = pid;
      bool tied = false;
      for (int o = WHITE; o <= RED; ++o) {                           // game.py:886 order White,Blue,Orange,Red
        if (o == pid) continue;
        const int pl = other_len[o];
        if (pl == max_len) tied = true;
        else if (pl > max_len) { max_len = pl; tied = false; player = o; }
      }
      if (max_len >= 5) {
        if (tied) {
          if (player == pid) g.lr_count = static_cast<uint8_t>(len);
          else { g.lr_holder = 0; g.lr_count = 0; g.vp[pid - 1] -= 2; }
        } else {
          g.lr_holder = static_cast<uint8_t>(player); g.lr_count = static_cast<uint8_t>(max_len);
          g.vp[player - 1] += 2; g.vp[pid - 1] -= 2;
        }
      } else { g.lr_holder = 0; g.lr_count = 0; g.vp[pid - 1] -= 2; }
    } else {
      g.lr_count = static_cast<uint8_t>(len);
    }
  } else if (len > count) {
    g.vp[holder - 1] -= 2; g.vp[pid - 1] += 2;
    g.lr_holder = static_cast<uint8_t>(pid); g.lr_count = static_cast<uint8_t>(len);
  }
}

CATAN_FN_NOINLINE void update_longest_road(Ctx& cx, int pid) {   // game.py:864-919  [W]
  GameRec& g = *cx.g;
  const int len = longest_path(cx, pid);
  const bool shrunk = lr_is_shrunk(g, pid, len);
  uint8_t other_len[5] = {0, 0, 0, 0, 0};
  if (shrunk) {
    for (int o = WHITE; o <= RED; ++o) if (o != pid) other_len[o] = static_cast<uint8_t>(longest_path(cx, o));
  }
  if (cx.lane == 0) lr_apply(g, pid, len, shrunk, other_len);
  wsync();
}

// ------------------------------------------------------------------------------------------------
// one env step (wrapper.py:36-50 without the observation)  [W]
//   ws.action must hold the composite action.  reward_out: float[4]; info_out: uint8[CATAN_INFO_STRIDE]
//   (both written by lane 0; may point to global memory).  Returns the error code (uniform).
// ------------------------------------------------------------------------------------------------
// phase 1 [L0]: clear the scratch, translate (wrapper.py:114-166) and validate (game.py:264-525) -> ws.err, ws.act
CATAN_FN_NOINLINE void step_begin(Ctx& cx) {
  GameRec& g = *cx.g;
  WarpScratch& ws = *cx.ws;
  ws.n_est = 0; ws.est_special = EST_SPECIAL_NONE; ws.dice_roll = 0; ws.lr_pid = 0; ws.roll_info = 0;
  ws.did_reset = 0; ws.done = 0;
  compute_seats(cx);
  ws.acted_pid = static_cast<uint8_t>(current_actor(g));
  ws.act_type = static_cast<uint8_t>(ws.action[CATAN_A_TYPE]);
  int err = translate_action(cx, ws.action, ws.act);
  if (!err && cx.cfg->validate_actions) err = validate_action(cx, ws.act);
  ws.err = static_cast<uint8_t>(err);
}

CATAN_FN_NOINLINE void step_finish(Ctx& cx, float* reward_out, uint8_t* info_out);

CATAN_FN int step_game(Ctx& cx, float* reward_out, uint8_t* info_out) {
  WarpScratch& ws = *cx.ws;
  if (cx.lane == 0) {
    step_begin(cx);
    if (!ws.err) apply_scalar(cx);
  }
  wsync();
  CATAN_PROF(cx, PH_SCALAR);
  const int err = ws.err;
  if (!err) {
    if (ws.dice_roll) { dice_payout(cx); CATAN_PROF(cx, PH_DICE); }
    if (ws.n_est || ws.est_special) { est_apply(cx); CATAN_PROF(cx, PH_EST); }
    if (ws.lr_pid) { update_longest_road(cx, ws.lr_pid); CATAN_PROF(cx, PH_LROAD); }
  }
  if (cx.lane == 0) step_finish(cx, reward_out, info_out);
  wsync();
  CATAN_PROF(cx, PH_FINISH);
  return err;
}

// last phase [L0]: done / reward / info (wrapper.py:85-112) and the optional auto-reset
CATAN_FN_NOINLINE void step_finish(Ctx& cx, float* reward_out, uint8_t* info_out) {
  GameRec& g = *cx.g;
  WarpScratch& ws = *cx.ws;
  const int err = ws.err;
  {
    struct alignas(16) V16 { uint32_t w[4]; };
    struct alignas(16) F4 { float v[4]; };
    F4 rew = {{0.f, 0.f, 0.f, 0.f}};
    int done = 0;
    if (!err) {
      g.episode_steps += 1;
      // game.py:18-23: dict order Blue, Red, Orange, White; the LAST player with >= 10 VP becomes env.winner
      if (g.vp[BLUE - 1] >= 10) { done = 1; g.winner = BLUE; }
      if (g.vp[RED - 1] >= 10) { done = 1; g.winner = RED; }
      if (g.vp[ORANGE - 1] >= 10) { done = 1; g.winner = ORANGE; }
      if (g.vp[WHITE - 1] >= 10) { done = 1; g.winner = WHITE; }
      if (cx.cfg->dense_reward) {                                    // wrapper.py:95-106
        const int ty = ws.act_type;
        const double bonus = (ty == CATAN_ACT_PLAY_DEV ? 5.0 : 0.0) + (ty == CATAN_ACT_MOVE_ROBBER ? 1.0 : 0.0) -
                             (ty == CATAN_ACT_DISCARD ? 0.3 : 0.0) + (ty == CATAN_ACT_UPGRADE_CITY ? 2.5 : 0.0);
        CATAN_NO_UNROLL
        for (int p = 0; p < 4; ++p) {
          // python: ((5*dvp + 5) + 1 - 0.3 + 2.5) * factor with at most one bonus non-zero => same double value
          const double r = (5.0 * static_cast<double>(g.vp[p] - g.curr_vps[p]) + bonus) * static_cast<double>(cx.cfg->reward_annealing_factor);
          rew.v[p] = static_cast<float>(r);
        }
      }
      CATAN_NO_UNROLL
      for (int p = 0; p < 4; ++p) g.curr_vps[p] = g.vp[p];
      if (done) {
        const int wp = g.winner - 1;
        rew.v[wp] = static_cast<float>(static_cast<double>(rew.v[wp]) + static_cast<double>(cx.cfg->win_reward));
      }
    }
    V16 info;
    info.w[0] = static_cast<uint32_t>(done) | (static_cast<uint32_t>(g.winner) << 8) |
                (static_cast<uint32_t>(static_cast<uint8_t>(g.vp[0])) << 16) | (static_cast<uint32_t>(static_cast<uint8_t>(g.vp[1])) << 24);
    uint32_t reset_flag = 0;
    const uint32_t actor_before_reset = static_cast<uint32_t>(current_actor(g));   // game_manager.py:99 reads it before env.reset()
    const uint32_t vp23 = static_cast<uint32_t>(static_cast<uint8_t>(g.vp[2])) | (static_cast<uint32_t>(static_cast<uint8_t>(g.vp[3])) << 8);
    if (done && cx.cfg->auto_reset) { reset_game(cx); reset_flag = 1; }
    info.w[1] = vp23 | (static_cast<uint32_t>(current_actor(g)) << 16) | (static_cast<uint32_t>(ws.acted_pid) << 24);
    info.w[2] = static_cast<uint32_t>(ws.act_type) | (static_cast<uint32_t>(ws.roll_info) << 8) |
                (static_cast<uint32_t>(err) << 16) | (reset_flag << 24);
    info.w[3] = actor_before_reset;
    ws.done = static_cast<uint8_t>(done);
    *reinterpret_cast<F4*>(reward_out) = rew;
    *reinterpret_cast<V16*>(info_out) = info;
  }
}

// ------------------------------------------------------------------------------------------------
// legal-action masks (wrapper.py:168-412)  [W]
// ------------------------------------------------------------------------------------------------
CATAN_FN_NOINLINE bool mask_play_dev(Ctx& cx, int pid) {   // wrapper.py:221-228 / :262-269 / :368-388; lane-0 writes
  const GameRec& g = *cx.g;
  uint8_t* m = cx.mask;
  const int p = pid - 1;
  if (g.n_hidden[p] == 0 || g.played_dev) return false;
  if (cx.lane == 0) {
    const int bank_total = g.bank[0] + g.bank[1] + g.bank[2] + g.bank[3] + g.bank[4];
    uint8_t valid[5];
    bool any = false;
    for (int c = 0; c < 5; ++c) {
      const int k = count_cards(g.hidden[p], g.n_hidden[p], c);
      valid[c] = (k > 0 && g.bought[c] < k && (c != CATAN_DEV_YOP || bank_total > 0)) ? 1 : 0;
      any |= valid[c] != 0;
    }
    if (any) {
      m[CATAN_MASK_TYPE + CATAN_ACT_PLAY_DEV] = 1;
      for (int c = 0; c < 5; ++c) m[CATAN_MASK_DEV + c] = valid[c];
      if (valid[CATAN_DEV_YOP]) {                                    // Q11: bank mask lands on row 2 of head 9 and on head 10
        for (int r = 0; r < 5; ++r) {
          const uint8_t b = g.bank[r] > 0;
          m[CATAN_MASK_RES_A + 10 + r] = b;
          m[CATAN_MASK_RES_B + r] = b;
        }
      }
    }
  }
  return true;
}

// road head (wrapper.py:322-339).  mode 0: main phase (write only when something is placeable, returns that);
// mode 1: initial phase (always write, dummy 0); mode 2: road building (always write, dummy iff nothing placeable)
CATAN_FN_NOINLINE bool mask_roads(Ctx& cx, int pid, int mode) {
  const GameRec& g = *cx.g;
  const Topo& T = *cx.T;
  uint8_t* m = cx.mask + CATAN_MASK_EDGE;
  bool after_second = false;
  int second = -1;
  if (g.initial_phase && g.init_settlements[g.players_go - 1] == 2) { after_second = true; second = g.second_corner[g.players_go - 1]; }
  uint8_t v[CATAN_PER_LANE(72)];
  bool mine = false;
#pragma unroll
  for (int it = 0; it < CATAN_PER_LANE(72); ++it) {
    const int e = cx.lane + it * CATAN_LANES;
    v[it] = (e < 72 && can_place_road(g, T, e, pid, after_second, second)) ? 1 : 0;
    mine |= v[it] != 0;
  }
  const bool placed = wany(mine);
  if (mode != 0 || placed) {
#pragma unroll
    for (int it = 0; it < CATAN_PER_LANE(72); ++it) {
      const int e = cx.lane + it * CATAN_LANES;
      if (e < 72) m[e] = v[it];
    }
    if (cx.lane == 0) m[72] = (mode == 2 && !placed) ? 1 : 0;
  }
  return placed;
}

CATAN_FN_NOINLINE void encode_masks(Ctx& cx) {
  const GameRec& g = *cx.g;
  const Topo& T = *cx.T;
  uint8_t* m = cx.mask;
  CATAN_LANE_LOOP(w, CATAN_MASK_STRIDE / 4) {                         // zeros for the type head and the pad, ones elsewhere
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int idx = 4 * w + b;
      if (idx >= CATAN_MASK_CORNER && idx < CATAN_MASK_ENTRIES) v |= 1u << (8 * b);
    }
    reinterpret_cast<uint32_t*>(m)[w] = v;
  }
  wsync();
  const int pid = g.players_go, p = pid - 1;
  const uint8_t* h = g.res[p];
  if (g.need_discard) {                                              // wrapper.py:186-192
    const int d = g.discard_queue[0];
    if (cx.lane == 0) m[CATAN_MASK_TYPE + CATAN_ACT_DISCARD] = 1;
    CATAN_LANE_LOOP(r, 5) if (g.res[d - 1][r] == 0) m[CATAN_MASK_DISCARD + r] = 0;
  } else if (g.initial_phase) {                                      // wrapper.py:195-204
    if (g.init_settlements[p] == 0 || (g.init_settlements[p] == 1 && g.init_roads[p] == 1)) {
      if (cx.lane == 0) m[CATAN_MASK_TYPE + CATAN_ACT_PLACE_SETTLEMENT] = 1;
      CATAN_LANE_LOOP(c, 54) m[CATAN_MASK_CORNER + c] = can_place_settlement(g, T, c, pid, true);
    } else {
      if (cx.lane == 0) m[CATAN_MASK_TYPE + CATAN_ACT_PLACE_ROAD] = 1;
      mask_roads(cx, pid, 1);
    }
  } else if (g.rb_active) {                                          // wrapper.py:206-209
    if (cx.lane == 0) m[CATAN_MASK_TYPE + CATAN_ACT_PLACE_ROAD] = 1;
    mask_roads(cx, pid, 2);
  } else if (g.just_moved_robber) {                                  // wrapper.py:210-213, :341-351
    if (cx.lane == 0) {
      m[CATAN_MASK_TYPE + CATAN_ACT_STEAL] = 1;
      uint8_t row[3] = {0, 0, 0};
      for (int k = 0; k < 6; ++k) {
        const uint8_t b = g.corner[T.tile_corners[g.robber_tile][k]];
        if (b && (b >> 2) != pid) row[label_of(cx, pid, b >> 2)] = 1;
      }
      for (int l = 0; l < 3; ++l) m[CATAN_MASK_PLAYER + 3 + l] = row[l];
    }
  } else if (g.must_respond) {                                       // wrapper.py:214-218, :353-365
    if (cx.lane == 0) {
      m[CATAN_MASK_TYPE + CATAN_ACT_RESPOND] = 1;
      int cnt[5] = {0, 0, 0, 0, 0};
      bool ok = true;
      for (int k = 0; k < g.n_recv; ++k) cnt[g.recv[k] - 1]++;
      for (int r = 0; r < 5; ++r) ok &= g.res[g.trade_target - 1][r] >= cnt[r];
      m[CATAN_MASK_ACCEPT] = ok;
    }
  } else if (!g.dice_rolled) {                                       // wrapper.py:219-229
    if (cx.lane == 0) m[CATAN_MASK_TYPE + CATAN_ACT_ROLL_DICE] = 1;
    mask_play_dev(cx, pid);
  } else {
    if (cx.lane == 0) m[CATAN_MASK_TYPE + CATAN_ACT_END_TURN] = 1;   // wrapper.py:232
    if (!(cx.cfg->max_actions_per_turn >= 0 && g.actions_this_turn > cx.cfg->max_actions_per_turn)) {
      if (h[WHEAT] && h[SHEEP] && h[WOOD] && h[BRICK]) {             // wrapper.py:238-243
        uint8_t v[CATAN_PER_LANE(54)];
        bool mine = false;
#pragma unroll
        for (int it = 0; it < CATAN_PER_LANE(54); ++it) {
          const int c = cx.lane + it * CATAN_LANES;
          v[it] = (c < 54 && can_place_settlement(g, T, c, pid, false)) ? 1 : 0;
          mine |= v[it] != 0;
        }
        if (wany(mine) && g.settlements_left[p] > 0) {
          if (cx.lane == 0) m[CATAN_MASK_TYPE + CATAN_ACT_PLACE_SETTLEMENT] = 1;
#pragma unroll
          for (int it = 0; it < CATAN_PER_LANE(54); ++it) {
            const int c = cx.lane + it * CATAN_LANES;
            if (c < 54) m[CATAN_MASK_CORNER + c] = v[it];
          }
        }
      }
      if (h[WHEAT] >= 2 && h[ORE] >= 3 && g.cities_left[p] > 0) {    // wrapper.py:245-250
        const uint8_t want = static_cast<uint8_t>((pid << 2) | 1);
        bool mine = false;
        CATAN_LANE_LOOP(c, 54) mine |= g.corner[c] == want;
        if (wany(mine)) {
          if (cx.lane == 0) m[CATAN_MASK_TYPE + CATAN_ACT_UPGRADE_CITY] = 1;
          CATAN_LANE_LOOP(c, 54) m[CATAN_MASK_CORNER + 54 + c] = g.corner[c] == want;
        }
      }
      if (h[WOOD] && h[BRICK]) {                                     // wrapper.py:252-256
        if (mask_roads(cx, pid, 0) && cx.lane == 0) m[CATAN_MASK_TYPE + CATAN_ACT_PLACE_ROAD] = 1;
      }
      if (cx.lane == 0 && h[WHEAT] && h[SHEEP] && h[ORE] && g.deck_n > 0) m[CATAN_MASK_TYPE + CATAN_ACT_BUY_DEV] = 1;
      mask_play_dev(cx, pid);                                        // wrapper.py:262-269
      wsync();                                                       // the exchange block may overwrite head 10
      if (cx.lane == 0) {
        uint8_t give[5], get[5];                                     // wrapper.py:271-276, :390-412 (Q13)
        bool ag = false, ar = false;
        for (int r = 0; r < 5; ++r) {
          give[r] = h[r] >= best_exchange_rate(g, pid, r);
          get[r] = g.bank[r] > 0;
          ag |= give[r] != 0; ar |= get[r] != 0;
        }
        if (ag && ar) {
          m[CATAN_MASK_TYPE + CATAN_ACT_EXCHANGE] = 1;
          for (int r = 0; r < 5; ++r) { m[CATAN_MASK_RES_A + r] = give[r]; m[CATAN_MASK_RES_B + r] = get[r]; }
        }
        if (hand_total(g, pid) > 0 &&                                // wrapper.py:283-289
            (cx.cfg->max_proposed_trades_per_turn < 0 || g.trades_this_turn < cx.cfg->max_proposed_trades_per_turn))
          m[CATAN_MASK_TYPE + CATAN_ACT_PROPOSE_TRADE] = 1;
        if (g.can_move_robber) m[CATAN_MASK_TYPE + CATAN_ACT_MOVE_ROBBER] = 1;
      }
      if (g.can_move_robber) {                                       // wrapper.py:278-281, :308-320 (Q1: any building)
        CATAN_LANE_LOOP(t, 19) {
          bool any = false;
#pragma unroll
          for (int k = 0; k < 6; ++k) any |= g.corner[T.tile_corners[t][k]] != 0;
          m[CATAN_MASK_TILE + t] = any;
        }
      }
    }
  }
  wsync();
}

// ------------------------------------------------------------------------------------------------
// observation (wrapper.py:52-83, :491-524, :526-709)  [W]
// ------------------------------------------------------------------------------------------------
CATAN_FN int bucket8(int n) { return n < 5 ? n : (n < 8 ? 5 : (n < 11 ? 6 : 7)); }                        // wrapper.py:554-561
CATAN_FN int bucket7(int n) { return n <= 2 ? n : (n <= 5 ? 3 : (n <= 7 ? 4 : (n <= 10 ? 5 : 6))); }       // wrapper.py:662-671
// slot of resource index r in the obs order Wood,Brick,Wheat,Ore,Sheep (wrapper.py:550)
CATAN_FN int obs_res_slot(int r) { return (0x24301 >> (4 * r)) & 7; }   // BRICK->1 WOOD->0 ORE->3 SHEEP->4 WHEAT->2

CATAN_FN_NOINLINE void encode_obs(Ctx& cx) {
  const GameRec& g = *cx.g;
  const Topo& T = *cx.T;
  uint8_t* o = cx.obs;
  {
    struct alignas(16) V16 { uint32_t a, b, c, d; };
    V16* ov = reinterpret_cast<V16*>(o);
    const V16 z = {0u, 0u, 0u, 0u};
    CATAN_LANE_LOOP(w, CATAN_OBS_STRIDE / 16) ov[w] = z;
  }
  const int actor = current_actor(g), ap = actor - 1;
  // relative seating in registers: REL(pid) = block of PlayerId pid seen from the actor (0 self, 1 next, ...),
  // PID_AT(rel) = PlayerId sitting rel seats after the actor (player.py:13-19)
  uint32_t relpack = 0, pidrel = 0;
  {
    const uint8_t* seat = cx.ws->seat;
    const int aseat = seat[actor];
#pragma unroll
    for (int p = 1; p <= 4; ++p) relpack |= static_cast<uint32_t>((seat[p] - aseat + 4) & 3) << (2 * p);
#pragma unroll
    for (int r = 0; r < 4; ++r) pidrel |= static_cast<uint32_t>(g.player_order[(aseat + r) & 3]) << (4 * r);
  }
#define CATAN_REL(pid_) ((relpack >> (2 * (pid_))) & 3u)
#define CATAN_PID_AT(rel_) ((pidrel >> (4 * (rel_))) & 15u)
#define CATAN_BLOCK(rel_) ((rel_) == 0 ? CATAN_OBS_CUR_MAIN : CATAN_OBS_OTHER_MAIN + ((rel_) - 1) * CATAN_OBS_OTHER_MAIN_DIM)
  wsync();
  CATAN_LANE_LOOP(i, 114) {                                          // (tile, corner) pairs: wrapper.py:499-521 and :595-610
    const int t = i / 6, k = i - 6 * t;
    const uint8_t b = g.corner[T.tile_corners[t][k]];
    uint8_t* cf = o + CATAN_OBS_TILES + t * CATAN_OBS_TILE_DIM + 18 + k * 7;
    cf[b & 3] = 1;                                                   // none / settlement / city
    if (b) {
      const int rel = CATAN_REL(b >> 2);
      cf[3 + rel] = 1;                                               // owner relative to the actor
      const int v = g.tile_val[t];
      if (v != 7)                                                    // production table of the owner's block (+1 / +2)
        sadd_u8(o + CATAN_BLOCK(rel) + (rel == 0 ? 50 : 90) + obs_res_slot(g.tile_res[t] - 1) * 10 + (v <= 6 ? v - 2 : v - 3), b & 3);
    }
  }
  CATAN_LANE_LOOP(t, 19) {                                           // wrapper.py:494-498
    uint8_t* f = o + CATAN_OBS_TILES + t * CATAN_OBS_TILE_DIM;
    f[0] = g.robber_tile == t;
    f[1 + g.tile_val[t] - 2] = 1;
    f[12 + g.tile_res[t]] = 1;
  }
#pragma unroll
  for (int li = 0; li < 5; ++li) {                                   // development-card lists, wrapper.py:642-655
    const int tp = (li < 2 ? actor : static_cast<int>(CATAN_PID_AT(li - 1))) - 1;
    const uint8_t* list = li == 1 ? g.hidden[tp] : g.played[tp];
    const int n = li == 1 ? g.n_hidden[tp] : g.n_played[tp];
    CATAN_LANE_LOOP(j, n) o[CATAN_OBS_DEV_LISTS + li * CATAN_OBS_DEV_PAD + j] = static_cast<uint8_t>(list[j] + 1);
  }
  wsync();   // word atomics of the production tables above share 32-bit words with the byte stores below
  CATAN_LANE_LOOP(vl, 32) {                                          // player blocks: lane = (block rel, feature group sub)
    const int rel = vl >> 3, sub = vl & 7;
    const int target = CATAN_PID_AT(rel), tp = target - 1;
    uint8_t* m = o + CATAN_BLOCK(rel);
    uint8_t* c = m + (rel == 0 ? 40 : 80);                           // vp 10 | production 50 | road 2 | army 2 | harbours 6
    if (sub < 5) {
      const int r = sub, slot = obs_res_slot(r);
      if (rel == 0) {
        m[slot * 8 + bucket8(g.res[ap][r])] = 1;                     // wrapper.py:550-562
        m[110 + slot * 7 + bucket7(g.bank[r])] = 1;                  // wrapper.py:657-672
        o[CATAN_OBS_CURRENT_RES + 1 + r] = g.res[ap][r];             // wrapper.py:70-71
      } else {
        m[slot * 8 + bucket8(g.est_min[ap][rel - 1][r])] = 1;        // wrapper.py:563-585
        m[40 + slot * 8 + bucket8(g.est_max[ap][rel - 1][r])] = 1;
      }
    } else if (sub == 5) {
      const int vps = g.vp[tp];
      c[vps < 10 ? vps : 9] = 1;                                     // wrapper.py:587-593
      if (g.lr_holder) {                                             // wrapper.py:613-620 (Q9)
        if (g.lr_holder == target) { c[60] = 1; c[61] = g.lr_count; }
        else if (g.has_path_key[tp]) c[61] = g.cur_longest_path[tp];
      }
    } else if (sub == 6) {
      if (g.la_holder == target) c[62] = 1;                          // wrapper.py:623-627 (Q10)
      c[63] = g.cur_army[tp];
      const int hb = g.harbours[tp];
#pragma unroll
      for (int b = 0; b < 6; ++b) c[64 + b] = (hb >> b) & 1;         // wrapper.py:632-637
    } else if (rel == 0) {
      m[145 + bucket7(g.deck_n)] = 1;                                // wrapper.py:674-686
      o[CATAN_OBS_META] = static_cast<uint8_t>(actor);
      o[CATAN_OBS_META + 1] = g.n_played[ap];
      o[CATAN_OBS_META + 2] = g.n_hidden[ap];
      if (g.trade_proposer) {                                        // wrapper.py:65-69 (Q15)
        for (int k = 0; k < g.n_give; ++k) o[CATAN_OBS_PROPOSED_TRADE + g.give[k]] = 1;
        for (int k = 0; k < g.n_recv; ++k) o[CATAN_OBS_PROPOSED_TRADE + g.recv[k] + 5] = 1;
      }
    } else {
      m[150 + rel - 1] = 1;                                          // wrapper.py:532-541
      const int nh = g.n_hidden[tp];
      m[153 + (nh <= 4 ? nh : 5)] = 1;                               // wrapper.py:690-695
      o[CATAN_OBS_META + 2 + rel] = g.n_played[tp];
    }
  }
#undef CATAN_REL
#undef CATAN_PID_AT
#undef CATAN_BLOCK
  wsync();
}

// ------------------------------------------------------------------------------------------------
// pinned random-legal sampler (twin of oracle/ref_harness.py:sample_action)  [W]
// m / hand may point to shared or global memory.  Result in a[CATAN_ACTION_WORDS] of every lane's
// registers is NOT materialised; lane 0 writes the words to `out`.
// ------------------------------------------------------------------------------------------------
#if CATAN_LANES == 32
CATAN_FN_NOINLINE int pick(const uint8_t* bits, int n, uint32_t w, int lane) {   // n <= 96; index of the floor(w*k/2^32)-th set entry
  const bool s0 = lane < n && bits[lane] != 0, s1 = lane + 32 < n && bits[lane + 32] != 0, s2 = lane + 64 < n && bits[lane + 64] != 0;
  const unsigned b0 = __ballot_sync(0xffffffffu, s0), b1 = __ballot_sync(0xffffffffu, s1), b2 = __ballot_sync(0xffffffffu, s2);
  const int k0 = __popc(b0), k1 = __popc(b1), k = k0 + k1 + __popc(b2);
  if (!k) return 0;
  const int j = static_cast<int>(__umulhi(w, static_cast<uint32_t>(k)));
  const unsigned below = (1u << lane) - 1u;
  // the lane whose set entry has rank j announces itself
  const bool h0 = s0 && __popc(b0 & below) == j;
  const bool h1 = s1 && k0 + __popc(b1 & below) == j;
  const bool h2 = s2 && k0 + k1 + __popc(b2 & below) == j;
  const unsigned m0 = __ballot_sync(0xffffffffu, h0), m1 = __ballot_sync(0xffffffffu, h1), m2 = __ballot_sync(0xffffffffu, h2);
  return m0 ? __ffs(m0) - 1 : (m1 ? 31 + __ffs(m1) : 63 + __ffs(m2));
}
#else
CATAN_FN_NOINLINE int pick(const uint8_t* bits, int n, uint32_t w, int) {
  int k = 0;
  for (int i = 0; i < n; ++i) k += bits[i] != 0;
  if (!k) return 0;
  int j = static_cast<int>(mulhi32(w, static_cast<uint32_t>(k)));
  for (int i = 0; i < n; ++i) if (bits[i]) { if (j == 0) return i; --j; }
  return 0;
}
#endif

// `hand`: the acting player's 5 resource counts (== obs current_resources[1..5], wrapper.py:70-71)
CATAN_FN_NOINLINE void sample_action(const uint8_t* m, const uint8_t* hand, uint64_t seed, uint64_t env_id, uint32_t decision,
                            int lane, int32_t* out) {
  uint32_t w[4];
  philox4x32(decision, CATAN_STREAM_SAMPLER, static_cast<uint32_t>(env_id), static_cast<uint32_t>(env_id >> 32),
             static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), w);
  const int t = pick(m + CATAN_MASK_TYPE, 13, w[0], lane);
  int corner = 0, edge = 0, tile = 0, card = 0, accept = 0, player = 0, give = 0, recv = 0, res_a = 0, res_b = 0, discard = 0;
  switch (t) {                                                       // t is warp-uniform
    case CATAN_ACT_PLACE_SETTLEMENT: corner = pick(m + CATAN_MASK_CORNER, 54, w[1], lane); break;
    case CATAN_ACT_UPGRADE_CITY: corner = pick(m + CATAN_MASK_CORNER + 54, 54, w[1], lane); break;
    case CATAN_ACT_PLACE_ROAD: edge = pick(m + CATAN_MASK_EDGE, 73, w[1], lane); break;
    case CATAN_ACT_MOVE_ROBBER: tile = pick(m + CATAN_MASK_TILE, 19, w[1], lane); break;
    case CATAN_ACT_PLAY_DEV:
      card = pick(m + CATAN_MASK_DEV, 5, w[1], lane);
      if (card == CATAN_DEV_MONOPOLY) res_a = pick(m + CATAN_MASK_RES_A + 10, 5, w[2], lane);
      else if (card == CATAN_DEV_YOP) {
        res_a = pick(m + CATAN_MASK_RES_A + 15, 5, w[2], lane);
        res_b = pick(m + CATAN_MASK_RES_B, 5, w[3], lane);
      }
      break;
    case CATAN_ACT_EXCHANGE:
      res_a = pick(m + CATAN_MASK_RES_A, 5, w[1], lane);
      res_b = pick(m + CATAN_MASK_RES_B, 5, w[2], lane);
      break;
    case CATAN_ACT_PROPOSE_TRADE:
      player = pick(m + CATAN_MASK_PLAYER, 3, w[1], lane);
      give = 1 + pick(hand, 5, w[2], lane);                            // a resource the proposer holds
      recv = 1 + static_cast<int>(mulhi32(w[3], 5u));
      break;
    case CATAN_ACT_RESPOND: accept = pick(m + CATAN_MASK_ACCEPT, 2, w[1], lane); break;
    case CATAN_ACT_STEAL: player = pick(m + CATAN_MASK_PLAYER + 3, 3, w[1], lane); break;
    case CATAN_ACT_DISCARD: discard = pick(m + CATAN_MASK_DISCARD, 5, w[1], lane); break;
    default: break;
  }
  if (lane < CATAN_ACTION_WORDS * (CATAN_LANES == 32 ? 1 : CATAN_ACTION_WORDS)) {
    CATAN_NO_UNROLL
    for (int i = lane; i < CATAN_ACTION_WORDS; i += CATAN_LANES) {       // one coalesced 80-byte row
      int v = 0;
      switch (i) {
        case CATAN_A_TYPE: v = t; break;
        case CATAN_A_CORNER: v = corner; break;
        case CATAN_A_EDGE: v = edge; break;
        case CATAN_A_TILE: v = tile; break;
        case CATAN_A_CARD: v = card; break;
        case CATAN_A_ACCEPT: v = accept; break;
        case CATAN_A_PLAYER: v = player; break;
        case CATAN_A_GIVE: v = give; break;
        case CATAN_A_RECV: v = recv; break;
        case CATAN_A_RES_A: v = res_a; break;
        case CATAN_A_RES_B: v = res_b; break;
        case CATAN_A_DISCARD: v = discard; break;
        default: break;
      }
      out[i] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// packed record <-> canonical state (host side of catan_export_state / catan_import_state)
// ------------------------------------------------------------------------------------------------
static inline void rec_to_state(const GameRec& g, catan_state_t& s) {
  memset(&s, 0, sizeof(s));
  for (int i = 0; i < 19; ++i) { s.tile_res[i] = g.tile_res[i]; s.tile_val[i] = g.tile_val[i]; }
  s.robber_tile = g.robber_tile;
  for (int i = 0; i < 54; ++i) { s.corner_type[i] = g.corner[i] & 3; s.corner_owner[i] = g.corner[i] >> 2; }
  for (int i = 0; i < 72; ++i) s.edge_owner[i] = g.edge[i];
  for (int i = 0; i < 9; ++i) s.harbour_perm[i] = g.harbour_perm[i];
  for (int i = 0; i < 4; ++i) s.player_order[i] = g.player_order[i];
  s.player_order_id = g.player_order_id; s.players_go = g.players_go;
  for (int p = 0; p < 4; ++p) {
    for (int r = 0; r < 5; ++r) {
      s.res[p][r] = g.res[p][r]; s.vis[p][r] = g.vis[p][r];
      for (int l = 0; l < 3; ++l) { s.est_min[p][l][r] = g.est_min[p][l][r]; s.est_max[p][l][r] = g.est_max[p][l][r]; }
    }
    s.vp[p] = g.vp[p]; s.harbours[p] = g.harbours[p];
    s.n_hidden[p] = g.n_hidden[p]; s.n_played[p] = g.n_played[p];
    for (int i = 0; i < 25; ++i) { s.hidden[p][i] = g.hidden[p][i]; s.played[p][i] = g.played[p][i]; }
    s.settlements_left[p] = g.settlements_left[p]; s.cities_left[p] = g.cities_left[p];
    s.init_settlements[p] = g.init_settlements[p]; s.init_roads[p] = g.init_roads[p];
    s.second_corner[p] = g.second_corner[p];
    s.cur_longest_path[p] = g.cur_longest_path[p]; s.has_path_key[p] = g.has_path_key[p]; s.cur_army[p] = g.cur_army[p];
    s.curr_vps[p] = g.curr_vps[p];
  }
  for (int r = 0; r < 5; ++r) s.bank[r] = g.bank[r];
  s.deck_n = g.deck_n;
  for (int i = 0; i < 25; ++i) s.deck[i] = g.deck[i];
  s.lr_holder = g.lr_holder; s.lr_count = g.lr_count; s.la_holder = g.la_holder; s.la_count = g.la_count;
  s.initial_phase = g.initial_phase; s.dice_rolled = g.dice_rolled; s.played_dev = g.played_dev;
  s.must_use_dev = g.must_use_dev; s.rb_active = g.rb_active; s.rb_count = g.rb_count;
  s.can_move_robber = g.can_move_robber; s.just_moved_robber = g.just_moved_robber;
  s.must_respond = g.must_respond; s.need_discard = g.need_discard;
  s.n_discard = g.n_discard;
  for (int i = 0; i < 4; ++i) { s.discard_queue[i] = g.discard_queue[i]; s.give[i] = g.give[i]; s.recv[i] = g.recv[i]; }
  s.trade_proposer = g.trade_proposer; s.trade_target = g.trade_target; s.n_give = g.n_give; s.n_recv = g.n_recv;
  s.die1 = g.die1; s.die2 = g.die2;
  s.trades_this_turn = g.trades_this_turn; s.actions_this_turn = static_cast<int16_t>(g.actions_this_turn);
  s.turn = static_cast<int16_t>(g.turn);
  for (int c = 0; c < 5; ++c) s.bought[c] = g.bought[c];
  s.winner = g.winner;
  s.rng_ctr_lo = static_cast<int16_t>(g.rng_ctr & 0xFFFF); s.rng_ctr_hi = static_cast<int16_t>(g.rng_ctr >> 16);
}

static inline void state_to_rec(const catan_state_t& s, GameRec& g) {
  const uint32_t dec = g.decision_ctr, steps = g.episode_steps;
  memset(&g, 0, sizeof(g));
  g.decision_ctr = dec; g.episode_steps = steps;
  for (int i = 0; i < 19; ++i) { g.tile_res[i] = static_cast<uint8_t>(s.tile_res[i]); g.tile_val[i] = static_cast<uint8_t>(s.tile_val[i]); }
  g.robber_tile = static_cast<uint8_t>(s.robber_tile);
  for (int i = 0; i < 54; ++i) g.corner[i] = static_cast<uint8_t>(s.corner_type[i] ? ((s.corner_owner[i] << 2) | s.corner_type[i]) : 0);
  for (int i = 0; i < 72; ++i) g.edge[i] = static_cast<uint8_t>(s.edge_owner[i]);
  for (int i = 0; i < 9; ++i) g.harbour_perm[i] = static_cast<uint8_t>(s.harbour_perm[i]);
  for (int i = 0; i < 4; ++i) g.player_order[i] = static_cast<uint8_t>(s.player_order[i]);
  g.player_order_id = static_cast<uint8_t>(s.player_order_id); g.players_go = static_cast<uint8_t>(s.players_go);
  for (int p = 0; p < 4; ++p) {
    for (int r = 0; r < 5; ++r) {
      g.res[p][r] = static_cast<uint8_t>(s.res[p][r]); g.vis[p][r] = s.vis[p][r];
      for (int l = 0; l < 3; ++l) { g.est_min[p][l][r] = s.est_min[p][l][r]; g.est_max[p][l][r] = s.est_max[p][l][r]; }
    }
    g.vp[p] = static_cast<int8_t>(s.vp[p]); g.harbours[p] = static_cast<uint8_t>(s.harbours[p]);
    g.n_hidden[p] = static_cast<uint8_t>(s.n_hidden[p]); g.n_played[p] = static_cast<uint8_t>(s.n_played[p]);
    for (int i = 0; i < 25; ++i) { g.hidden[p][i] = static_cast<uint8_t>(s.hidden[p][i]); g.played[p][i] = static_cast<uint8_t>(s.played[p][i]); }
    g.settlements_left[p] = static_cast<uint8_t>(s.settlements_left[p]); g.cities_left[p] = static_cast<uint8_t>(s.cities_left[p]);
    g.init_settlements[p] = static_cast<uint8_t>(s.init_settlements[p]); g.init_roads[p] = static_cast<uint8_t>(s.init_roads[p]);
    g.second_corner[p] = static_cast<int8_t>(s.second_corner[p]);
    g.cur_longest_path[p] = static_cast<uint8_t>(s.cur_longest_path[p]); g.has_path_key[p] = static_cast<uint8_t>(s.has_path_key[p]);
    g.cur_army[p] = static_cast<uint8_t>(s.cur_army[p]);
    g.curr_vps[p] = static_cast<int8_t>(s.curr_vps[p]);
  }
  for (int r = 0; r < 5; ++r) g.bank[r] = static_cast<uint8_t>(s.bank[r]);
  g.deck_n = static_cast<uint8_t>(s.deck_n);
  for (int i = 0; i < 25; ++i) g.deck[i] = static_cast<uint8_t>(s.deck[i]);
  g.lr_holder = static_cast<uint8_t>(s.lr_holder); g.lr_count = static_cast<uint8_t>(s.lr_count);
  g.la_holder = static_cast<uint8_t>(s.la_holder); g.la_count = static_cast<uint8_t>(s.la_count);
  g.initial_phase = static_cast<uint8_t>(s.initial_phase); g.dice_rolled = static_cast<uint8_t>(s.dice_rolled);
  g.played_dev = static_cast<uint8_t>(s.played_dev); g.must_use_dev = static_cast<uint8_t>(s.must_use_dev);
  g.rb_active = static_cast<uint8_t>(s.rb_active); g.rb_count = static_cast<uint8_t>(s.rb_count);
  g.can_move_robber = static_cast<uint8_t>(s.can_move_robber); g.just_moved_robber = static_cast<uint8_t>(s.just_moved_robber);
  g.must_respond = static_cast<uint8_t>(s.must_respond); g.need_discard = static_cast<uint8_t>(s.need_discard);
  g.n_discard = static_cast<uint8_t>(s.n_discard);
  for (int i = 0; i < 4; ++i) {
    g.discard_queue[i] = static_cast<uint8_t>(s.discard_queue[i]);
    g.give[i] = static_cast<uint8_t>(s.give[i]); g.recv[i] = static_cast<uint8_t>(s.recv[i]);
  }
  g.trade_proposer = static_cast<uint8_t>(s.trade_proposer); g.trade_target = static_cast<uint8_t>(s.trade_target);
  g.n_give = static_cast<uint8_t>(s.n_give); g.n_recv = static_cast<uint8_t>(s.n_recv);
  g.die1 = static_cast<uint8_t>(s.die1); g.die2 = static_cast<uint8_t>(s.die2);
  g.trades_this_turn = static_cast<uint8_t>(s.trades_this_turn);
  g.actions_this_turn = static_cast<uint16_t>(s.actions_this_turn); g.turn = static_cast<uint16_t>(s.turn);
  for (int c = 0; c < 5; ++c) g.bought[c] = static_cast<uint8_t>(s.bought[c]);
  g.winner = static_cast<uint8_t>(s.winner);
  g.rng_ctr = static_cast<uint32_t>(static_cast<uint16_t>(s.rng_ctr_lo)) | (static_cast<uint32_t>(static_cast<uint16_t>(s.rng_ctr_hi)) << 16);
}

}  // namespace catanb
