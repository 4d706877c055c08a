// catan_core.cuh — what the engine (catan_game.cuh) is built on: the board topology, the packed game record and its
// conversion to the canonical state, the pinned Philox RNG, and the cooperative longest-road search.
//
// The same source compiles two ways:
//   * nvcc (product): every function is __device__, CATAN_LANES == 32;
//   * g++ with -DCATAN_HOST_EMU (tests only, tests/host_emu/): CATAN_LANES == 1, warp collectives become identities.
//     This lets the CPU test-suite run exactly this logic against the oracle and the golden fixtures before any GPU
//     time is spent.  It is not a shipped fallback: the C-ABI library is only ever built from the .cu files.
//
// Reference behaviour reproduced here (file:line are in /root/reference): game/game.py (Game),
// game/components/{board,corner,edge,player}.py, env/wrapper.py (EnvWrapper).  See SURVEY.md §8a.
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../../include/catan_layout.h"
#define CATAN_TOPO_MACROS_ONLY
#include "../../include/catan_topology.h"

#if defined(__CUDACC__) && !defined(CATAN_HOST_EMU)
#define CATAN_DEVICE 1
#define CATAN_FN __device__ __forceinline__
// phase-sized functions are emitted ONCE and called: the kernels must stay small enough for the instruction caches
// (profiles/r1_notes.md: 60% of stalls were stall_no_inst with everything inlined)
#define CATAN_FN_NOINLINE __device__ __noinline__
#define CATAN_NO_UNROLL _Pragma("unroll 1")
#define CATAN_LANES 32
// The phase functions take their game through a pointer the compiler cannot trace back to the kernel's shared memory: without a
// hint every access to the staged chunk and topology is a generic LD / ST (L1TEX address check, "long scoreboard" stalls) instead
// of LDS / STS.  A function that only ever runs on staged data says so.
#define CATAN_IN_SMEM(p_) __builtin_assume(__isShared(p_))
#define CATAN_IN_LOCAL(p_) __builtin_assume(__isLocal(p_))     // (a caller's stack object passed by reference)
#else
#define CATAN_IN_SMEM(p_) ((void)0)
#define CATAN_IN_LOCAL(p_) ((void)0)
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
// packed per-game record: 832 bytes (the card lists are 4-bit codes: 208 -> 104 bytes, which pays for the production cache).  In HBM 32 records are interleaved into one chunk (catan_game.cuh: GameView).
// Field meaning == catan_state_t (catan_layout.h).
// ------------------------------------------------------------------------------------------------
struct alignas(16) GameRec {
  // ---- cold, 16 bit: the belief tables and the production cache (follow-up passes, observation)
  int16_t est_min[4][3][5];   // opponent_min_res[observer][label][r]        (player.py:38-43)
  int16_t est_max[4][3][5];
  uint16_t prod[4][13];       // engine-private cache of the observation's production table (wrapper.py:595-610): 4-bit counts,
                              // entry j = slot * 10 + number slot in the obs order at bits 4 * (j & 3) of prod[p][j >> 2]; kept
                              // current by the transition (a settlement adds 1, a city one more per adjacent tile)
  // ---- HOT [CATAN_HOT_BEGIN, CATAN_HOT_END): what the rules of almost every action read and write.  The transition kernel stages
  // only this part of a chunk in shared memory (6 KB instead of 26 KB) and touches the cold parts in place.
  int16_t vis[4][5];          // visible_resources (unbounded growth through trades -> 16 bit)
  uint32_t rng_ctr;           // game-stream Philox draw counter
  uint32_t decision_ctr;      // sampler-stream decision index
  uint32_t episode_steps;     // env steps since the last reset
  uint16_t actions_this_turn;
  uint16_t turn;
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
  uint8_t bank[5];
  uint8_t deck_n;
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
  uint8_t lr_dirty[4];        // engine-private (not part of the canonical state): an opponent built next to this player's
                              // roads since cur_longest_path was measured -> the incremental update is not allowed
  // ---- cold, bytes: the board and the card lists (placements, dice payout, development cards, observation)
  uint8_t corner[54];         // (owner PlayerId << 2) | type (0 none, 1 settlement, 2 city)
  uint8_t edge[72];           // road owner PlayerId, 0 none
  uint8_t tile_res[19];
  uint8_t tile_val[19];
  uint8_t harbour_perm[9];
  uint8_t hidden[4][13];      // ordered card lists, two cards per byte (card i at bits 4 * (i & 1) of byte i >> 1; unused = 0)
  uint8_t played[4][13];
  uint8_t deck[25];
  uint8_t pad_[1];
};
#define CATAN_HOT_BEGIN 344   /* offsetof(GameRec, vis) */
#define CATAN_HOT_END 529     /* offsetof(GameRec, corner) */
static_assert(sizeof(GameRec) == 832, "GameRec layout changed: keep it a multiple of 16 bytes and update DESIGN.md");
static_assert(offsetof(GameRec, vis) == CATAN_HOT_BEGIN && offsetof(GameRec, corner) == CATAN_HOT_END && CATAN_HOT_BEGIN % 8 == 0, "the hot range of a record");

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

// ------------------------------------------------------------------------------------------------
// warp primitives
// ------------------------------------------------------------------------------------------------
#if CATAN_LANES == 32
CATAN_FN void wsync() { __syncwarp(); }
CATAN_FN bool wany(bool p) { return __any_sync(0xffffffffu, p) != 0; }
// byte add in shared memory: bytes never overflow here (small counters), so a word atomic is exact
CATAN_FN void sadd_u8(uint8_t* p, int v) {
  uintptr_t a = reinterpret_cast<uintptr_t>(p);
  atomicAdd(reinterpret_cast<unsigned int*>(a & ~uintptr_t(3)), static_cast<unsigned int>(v) << (8 * (a & 3)));
}
CATAN_FN uint32_t mulhi32(uint32_t a, uint32_t b) { return __umulhi(a, b); }
CATAN_FN int ctz64(uint64_t x) { return __ffsll(static_cast<long long>(x)) - 1; }
CATAN_FN int fetch_add_i32(int32_t* p, int v) { return atomicAdd(p, v); }
CATAN_FN void smax_i32(int32_t* p, int v) { atomicMax(p, v); }
CATAN_FN void smin_i32(int32_t* p, int v) { atomicMin(p, v); }
#else
CATAN_FN void wsync() {}
CATAN_FN bool wany(bool p) { return p; }
CATAN_FN void sadd_u8(uint8_t* p, int v) { *p = static_cast<uint8_t>(*p + v); }
CATAN_FN uint32_t mulhi32(uint32_t a, uint32_t b) { return static_cast<uint32_t>((static_cast<uint64_t>(a) * b) >> 32); }
CATAN_FN int ctz64(uint64_t x) { return __builtin_ctzll(x); }
CATAN_FN int fetch_add_i32(int32_t* p, int v) { const int o = *p; *p += v; return o; }
CATAN_FN void smax_i32(int32_t* p, int v) { if (v > *p) *p = v; }
CATAN_FN void smin_i32(int32_t* p, int v) { if (v < *p) *p = v; }
#endif

// ------------------------------------------------------------------------------------------------
// pinned RNG (catan_layout.h): Philox4x32-10
// ------------------------------------------------------------------------------------------------
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
// word `idx` of the same block, returned in a register: a caller that wants ONE word has no output array on its stack (the
// address of such an array, handed to the called function, defeats the compiler's stack-slot lifetime analysis -- on the device
// Game.randomise_uncertainty's hand totals were overwritten by the four words)
CATAN_FN_NOINLINE uint32_t philox4x32_word(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, int idx) {
  CATAN_NO_UNROLL
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return idx == 0 ? c0 : idx == 1 ? c1 : idx == 2 ? c2 : c3;
}


// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
CATAN_FN int clipi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
CATAN_FN void est_set(EstReq& q, int r, int d) { q.delta[r] = static_cast<int8_t>(d); q.touched |= static_cast<uint8_t>(1u << r); }

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


// ------------------------------------------------------------------------------------------------
// Longest road: cooperative enumeration of node-simple paths (game.py:843-862, utils.py:3-15; Q7).
//
// The reference enumerates every simple path of the player's road graph from every start corner.  Here the graph is
// one 64-bit adjacency mask per corner (a corner holding an opponent's building keeps its incoming arcs but gets no
// outgoing ones, game.py:851-858) and the enumeration is a POOL of tasks that any number of lanes drain together.
// A task is a subtree of the search: (set of visited corners, current corner, depth) -- nothing else is needed to walk
// down from there.  A lane walks its task with a branch-free push/pop state machine over its own small stack.  Whenever
// the pool runs low it cuts the untried siblings of the SHALLOWEST open level of its stack (the biggest pieces of what
// it has left) off as new tasks; idle lanes pick them up at once.  There are no rounds and no barriers inside the
// search: a dense network (10^4..10^5 path visits) spreads over the whole group within a few iterations, a chain-like
// one is simply walked by one lane.  The deepest level seen is max-ed into *best.
//
// "Through" mode (sw >= 0): the walk first follows adjb (arcs INTO the current corner, plus bit sw at every corner =
// "stop walking backwards and jump to corner sw"), and follows adj once sw has been visited.  With the tables of
// t_lp_build_adj2 (catan_game.cuh) for a new road a -> sw and the single seed (a) this enumerates exactly the simple
// paths that contain that arc, the depth being the path length.
//
// Pool protocol (ctl[0] = claim cursor C, ctl[1] = write cursor R, ctl[2] = unfinished tasks W; ring slot i & mask holds
// task i once its seq == i + 1): a producer adds to W, reserves slots by advancing R, writes the tasks and publishes
// each with its seq; an idle lane advances C to get a slot number -- possibly one that is not written yet, which it
// then simply waits for -- and the pool is dead when W == 0.  Producers only refill while R - C is small, so the ring
// (a power of two, several times the number of lanes) is never lapped.
// ------------------------------------------------------------------------------------------------
#define CATAN_LP_ADJ_BYTES 432
#define CATAN_LP_CTL_WORDS 4
#define CATAN_LP_BURST 4
struct alignas(16) LpTask {
  uint64_t visited;
  uint8_t node, depth, pad_[2];
  uint32_t seq;
};
static_assert(sizeof(LpTask) == 16, "LpTask layout");
#if CATAN_LANES == 32
CATAN_FN int32_t vload_i32(const int32_t* p) { return *reinterpret_cast<const volatile int32_t*>(p); }
CATAN_FN uint32_t vload_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
CATAN_FN void vstore_u32(uint32_t* p, uint32_t v) { *reinterpret_cast<volatile uint32_t*>(p) = v; }
CATAN_FN void group_fence() { __threadfence_block(); }
#ifndef CATAN_LP_IDLE_NS
#define CATAN_LP_IDLE_NS 200
#endif
CATAN_FN void group_idle() { __nanosleep(CATAN_LP_IDLE_NS); }
#else
CATAN_FN int32_t vload_i32(const int32_t* p) { return *p; }
CATAN_FN uint32_t vload_u32(const uint32_t* p) { return *p; }
CATAN_FN void vstore_u32(uint32_t* p, uint32_t v) { *p = v; }
CATAN_FN void group_fence() {}
CATAN_FN void group_idle() {}
#endif

CATAN_FN_NOINLINE void lp_pool(const uint64_t* adj, const uint64_t* adjb, int sw, int32_t* ctl, int32_t* best,
                               uint8_t* path, int path_stride, int path_lane, LpTask* ring, int ring_mask, int low_water) {
#define CATAN_LP_CAND(u_, vis_) ((sw >= 0 && !(((vis_) >> sw) & 1ull)) ? adjb[(u_)] : adj[(u_)])
  int lbest = 0, node = 0, depth = 0, base = 0, my = -1, steps = 0;
  uint64_t visited = 0, vbase = 0, above = ~0ull;
  bool active = false;
  for (;;) {
    if (!active) {
      if (my < 0 && vload_i32(&ctl[0]) < vload_i32(&ctl[1])) my = fetch_add_i32(&ctl[0], 1);   // a slot number, maybe of a future task
      if (my >= 0) {
        LpTask& tk = ring[my & ring_mask];
        if (vload_u32(&tk.seq) == static_cast<uint32_t>(my) + 1u) {
          group_fence();
          visited = vbase = tk.visited; node = tk.node; depth = base = tk.depth; above = ~0ull;
          path[base * path_stride + path_lane] = static_cast<uint8_t>(node);
          if (lbest < depth) lbest = depth;
          active = true; my = -1;
        }
      }
    }
    if (!wany(active || vload_i32(&ctl[2]) > 0)) break;             // nobody in this warp works and the pool is dead
    if (!wany(active)) group_idle();                                 // a warp without work must not take issue slots from the others
    if (!active) continue;
    if (vload_i32(&ctl[1]) - vload_i32(&ctl[0]) < low_water) {
      // the pool runs low: give the untried siblings of the shallowest open level away
      uint64_t vis = vbase;
      for (int q = base; q <= depth; ++q) {
        const int raw = path[q * path_stride + path_lane], u = raw & 63;
        vis |= 1ull << u;
        if (raw & 128) continue;                                     // already given away
        uint64_t c = CATAN_LP_CAND(u, vis) & ~vis;
        c &= q == depth ? above : ~((2ull << (path[(q + 1) * path_stride + path_lane] & 63)) - 1ull);
        if (!c) continue;
        int need = 0;
        for (uint64_t t = c; t; t &= t - 1) ++need;
        fetch_add_i32(&ctl[2], need);
        int w = fetch_add_i32(&ctl[1], need);
        for (; c; c &= c - 1, ++w) {
          const int t = ctz64(c);
          LpTask& tk = ring[w & ring_mask];
          tk.visited = vis | (1ull << t); tk.node = static_cast<uint8_t>(t); tk.depth = static_cast<uint8_t>(q + 1);
          group_fence();
          vstore_u32(&tk.seq, static_cast<uint32_t>(w) + 1u);
        }
        path[q * path_stride + path_lane] = static_cast<uint8_t>(raw | 128);
        if (q == depth) above = 0ull;
        break;
      }
    }
    CATAN_NO_UNROLL
    for (int k = 0; k < CATAN_LP_BURST && active; ++k) {             // a burst of walk steps between two looks at the pool
      ++steps;
      const uint64_t cand = CATAN_LP_CAND(node, visited) & ~visited & above;
      if (cand) {                                                    // push the lowest untried neighbour
        const int t = ctz64(cand);
        ++depth;
        path[depth * path_stride + path_lane] = static_cast<uint8_t>(t);
        visited |= 1ull << t;
        node = t; above = ~0ull;
        if (depth > lbest) lbest = depth;
      } else if (depth == base) {
        active = false;                                              // task exhausted
        fetch_add_i32(&ctl[2], -1);
      } else {                                                       // pop; resume the parent above the popped child
        visited &= ~(1ull << node);
        --depth;
        const int raw = path[depth * path_stride + path_lane];
        above = (raw & 128) ? 0ull : ~((2ull << node) - 1ull);       // the siblings of a level that was given away are not ours
        node = raw & 63;
      }
    }
  }
  if (lbest) smax_i32(best, lbest);
  if (steps) fetch_add_i32(&ctl[3], steps);                          // diagnostics: walk steps of this search
#undef CATAN_LP_CAND
}

// One round of the level-synchronous opening (CATAN_LP_RUN): the lane consumes task `idx` of the ring and appends one task per
// untried neighbour (the cursors: ctl[0] first live task, ctl[1] write cursor; barriers between rounds are the caller's).
CATAN_FN void lp_open_level(const uint64_t* adj, const uint64_t* adjb, int sw, int32_t* ctl, LpTask* ring, int ring_mask, int idx,
                            int& lbest, int& steps) {
  const LpTask tk = ring[idx & ring_mask];
  const uint64_t vis = tk.visited;
  uint64_t c = ((sw >= 0 && !((vis >> (sw < 0 ? 0 : sw)) & 1ull)) ? adjb[tk.node] : adj[tk.node]) & ~vis;
  ++steps;
  if (lbest < tk.depth) lbest = tk.depth;
  if (!c) return;
  int need = 0;
  for (uint64_t t = c; t; t &= t - 1) ++need;
  int w = fetch_add_i32(&ctl[1], need);
  for (; c; c &= c - 1, ++w) {
    const int t = ctz64(c);
    LpTask& nt = ring[w & ring_mask];
    nt.visited = vis | (1ull << t); nt.node = static_cast<uint8_t>(t); nt.depth = static_cast<uint8_t>(tk.depth + 1);
    nt.seq = static_cast<uint32_t>(w) + 1u;
  }
}

// Driver shared by the group-local and the block-cooperative callers: SYNC_ is the barrier of the participating threads,
// leader_ is true for exactly one of them.  The leader seeds the pool: one task per corner with an outgoing arc
// (start_ < 0), or the single corner start_.  lanes_ = number of participating threads (>= 1; the ring holds at least
// 4 * lanes_ + 64 tasks).  The caller must have made the adjacency tables and *best_ visible (one barrier) before.
#define CATAN_LP_RUN(adj_, adjb_, sw_, start_, ctl_, best_, path_, lanes_, plane_, ring_, cap_, leader_, SYNC_)            \
  do {                                                                                                                     \
    for (int v_ = (plane_); v_ < (cap_); v_ += (lanes_)) (ring_)[v_].seq = 0u;   /* no task of an earlier search is valid */ \
    if (leader_) { (ctl_)[0] = 0; (ctl_)[1] = 0; (ctl_)[2] = 0; (ctl_)[3] = 0; }                                           \
    SYNC_;                                                                                                                 \
    /* the seeds: the single corner start_, or (every lane its share of the corners) each corner with an outgoing arc */   \
    for (int v_ = (start_) >= 0 ? ((leader_) ? (start_) : 54) : (plane_); v_ < 54; v_ += (start_) >= 0 ? 54 : (lanes_)) {  \
      if ((start_) < 0 && (adj_)[v_] == 0ull) continue;                                                                    \
      const int n_ = fetch_add_i32(&(ctl_)[1], 1);                                                                         \
      LpTask& tk_ = (ring_)[n_];                                                                                           \
      tk_.visited = 1ull << v_; tk_.node = static_cast<uint8_t>(v_); tk_.depth = 0; tk_.seq = static_cast<uint32_t>(n_) + 1u; \
    }                                                                                                                      \
    SYNC_;                                                                                                                 \
    if ((lanes_) < 32) { if (leader_) (ctl_)[2] = (ctl_)[1]; SYNC_; }                                                      \
    /* Level-synchronous opening (>= 32 lanes): as long as the frontier of the search tree fits the lanes, every lane unfolds ONE */ \
    /* open subtree by one level per round (lp_open_level) -- a deep, narrow tree (a chain of roads: 64-256 walk steps at ~200   */ \
    /* cycles each for one walking lane) is finished in `depth` rounds, and a bushy one reaches the pool with a task for every   */ \
    /* lane instead of growing there one hand-over at a time (profiles/r2_notes.md).  While the frontier fits ONE warp the       */ \
    /* rounds are that warp's alone (a warp barrier costs a fifth of a block barrier); then the whole group joins.               */ \
    if ((lanes_) >= 32) {                                                                                                  \
      int lbest_ = 0, steps_ = 0;                                                                                          \
      if ((plane_) < 32) {                                                                                                 \
        for (int round_ = 0; round_ < 64; ++round_) {                                                                      \
          const int lo_ = vload_i32(&(ctl_)[0]), hi_ = vload_i32(&(ctl_)[1]), nf_ = hi_ - lo_;                             \
          wsync();                                                                                                         \
          if (nf_ == 0 || nf_ > 32 || 4 * nf_ > (cap_) - 64) break;                                                        \
          if ((plane_) < nf_) lp_open_level(adj_, adjb_, sw_, ctl_, ring_, (cap_) - 1, lo_ + (plane_), lbest_, steps_);    \
          if (leader_) (ctl_)[0] = hi_;                                                                                    \
          wsync();                                                                                                         \
        }                                                                                                                  \
      }                                                                                                                    \
      SYNC_;                                                                                                               \
      for (int round_ = 0; round_ < 64; ++round_) {                                                                        \
        const int lo_ = vload_i32(&(ctl_)[0]), hi_ = vload_i32(&(ctl_)[1]), nf_ = hi_ - lo_;                               \
        SYNC_;                                               /* everybody has read the cursors of this round */            \
        if (nf_ == 0 || nf_ > (lanes_) || 4 * nf_ > (cap_) - 64) break;                                                    \
        if ((plane_) < nf_) lp_open_level(adj_, adjb_, sw_, ctl_, ring_, (cap_) - 1, lo_ + (plane_), lbest_, steps_);      \
        if (leader_) (ctl_)[0] = hi_;                        /* this level is consumed */                                  \
        SYNC_;                                               /* the next level and the cursors are visible */              \
      }                                                                                                                    \
      if (lbest_) smax_i32(best_, lbest_);                                                                                 \
      if (steps_) fetch_add_i32(&(ctl_)[3], steps_);                                                                       \
      if (leader_) (ctl_)[2] = vload_i32(&(ctl_)[1]) - vload_i32(&(ctl_)[0]);   /* what is left for the pool */             \
      SYNC_;                                                                                                               \
    }                                                                                                                      \
    lp_pool(adj_, adjb_, sw_, ctl_, best_, path_, lanes_, plane_, ring_, (cap_) - 1, (lanes_) < 4 ? 2 : (lanes_) / 2);   \
    SYNC_;                                                                                                                 \
  } while (0)

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
    for (int i = 0; i < 25; ++i) {
      s.hidden[p][i] = i < g.n_hidden[p] ? (g.hidden[p][i >> 1] >> (4 * (i & 1))) & 15 : 0;
      s.played[p][i] = i < g.n_played[p] ? (g.played[p][i >> 1] >> (4 * (i & 1))) & 15 : 0;
    }
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

// obs slot of resource index r (BRICK..WHEAT) in the order Wood, Brick, Wheat, Ore, Sheep (wrapper.py:550), and the number slot of
// a token (2-6 -> 0-4, 8-12 -> 5-9; wrapper.py:600-604)
#define CATAN_OBS_RES_SLOT(r_) ((0x24301 >> (4 * (r_))) & 7)
#define CATAN_OBS_NUM_SLOT(v_) ((v_) <= 6 ? (v_) - 2 : (v_) - 3)
// the production cache from scratch (host side: catan_import_state)
static inline void rebuild_prod(GameRec& g) {
  static const Topo T = CATAN_TOPO_INITIALIZER;
  memset(g.prod, 0, sizeof(g.prod));
  for (int t = 0; t < 19; ++t) {
    if (g.tile_res[t] == 0 || g.tile_val[t] == 7) continue;
    const int j = CATAN_OBS_RES_SLOT(g.tile_res[t] - 1) * 10 + CATAN_OBS_NUM_SLOT(g.tile_val[t]);
    for (int k = 0; k < 6; ++k) {
      const int b = g.corner[T.tile_corners[t][k]];
      if (b) g.prod[(b >> 2) - 1][j >> 2] = static_cast<uint16_t>(g.prod[(b >> 2) - 1][j >> 2] + ((b & 3) << (4 * (j & 3))));
    }
  }
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
    for (int i = 0; i < 25; ++i) {                              // (entries beyond the list length stay zero: the encoder relies on it)
      if (i < s.n_hidden[p]) g.hidden[p][i >> 1] |= static_cast<uint8_t>((s.hidden[p][i] & 15) << (4 * (i & 1)));
      if (i < s.n_played[p]) g.played[p][i >> 1] |= static_cast<uint8_t>((s.played[p][i] & 15) << (4 * (i & 1)));
    }
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
  for (int p = 0; p < 4; ++p) g.lr_dirty[p] = 1;   // nothing is known about how cur_longest_path relates to the imported board
  rebuild_prod(g);
  g.rng_ctr = static_cast<uint32_t>(static_cast<uint16_t>(s.rng_ctr_lo)) | (static_cast<uint32_t>(static_cast<uint16_t>(s.rng_ctr_hi)) << 16);
}

}  // namespace catanb
