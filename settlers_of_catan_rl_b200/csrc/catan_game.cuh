// catan_game.cuh — the Catan env-step engine, ONE THREAD PER GAME.
//
// Why thread-per-game (profiles/r2_notes.md): the first engine gave every game a warp.  Its lane-parallel loops are
// 19/54/72/114 wide and its rule code is scalar, so a warp instruction did useful work for ~16 lanes of ONE game and the
// kernel needed 2.9 k warp instructions per env step (issue-bound at 6 % of the HBM roofline).  Here a warp instruction
// serves 32 games.  That only pays if the 32 lanes touch memory together, so the packed game records are stored
// LANE-INTERLEAVED: 32 consecutive games form a chunk in which byte k of game l lives at chunk + 32*k + l (16- and 32-bit
// fields interleave in units of their own size).  Every `g.field(i)` below is then one fully used 32-byte sector per
// warp, and a per-game pointer chase costs no more than a coalesced load.
//
// The same source compiles two ways, like catan_core.cuh: nvcc (product, CATAN_W == 32) and g++ -DCATAN_HOST_EMU
// (tests/host_emu only, CATAN_W == 1: a "chunk" is one plain GameRec), so the CPU test-suite runs exactly this logic
// against the oracle and the golden fixtures.
//
// Reference behaviour reproduced here (file:line are in /root/reference): game/game.py (Game),
// game/components/{board,corner,edge,player}.py, env/wrapper.py (EnvWrapper).  See SURVEY.md §8a.
#pragma once

#include <stddef.h>
#include <type_traits>

#include "catan_core.cuh"

namespace catanb {

#ifdef CATAN_DEVICE
#define CATAN_W 32
#define CATAN_MFN __device__ __forceinline__
#else
#define CATAN_W 1
#define CATAN_MFN inline
#endif
#define CATAN_CHUNK_BYTES (sizeof(GameRec) * CATAN_W)

// ------------------------------------------------------------------------------------------------
// view of one game inside a lane-interleaved chunk.  Field names / meaning == GameRec (catan_core.cuh).
// ------------------------------------------------------------------------------------------------
struct GameView {
  uint8_t* base;   // where the HOT fields start (record offset CATAN_HOT_BEGIN of the chunk; the range ends at CATAN_HOT_END) ...
  int lane;        // game index inside the chunk
  uint8_t* cold;   // ... and the chunk base, for all the other fields.  One chunk (base == cold + CATAN_HOT_BEGIN * CATAN_W) except in
                   // the transition kernel, which stages only the hot part of a chunk in shared memory and leaves the rest in place.
  GameView() = default;
  CATAN_MFN GameView(uint8_t* chunk, int l) : base(chunk + static_cast<size_t>(CATAN_HOT_BEGIN) * CATAN_W), lane(l), cold(chunk) {}
  CATAN_MFN GameView(uint8_t* hot, int l, uint8_t* chunk) : base(hot), lane(l), cold(chunk) {}
  // field at record offset `off` (a compile-time constant at every call site: the choice of the pointer folds away), element k
  template <class T>
  CATAN_MFN T& at(int off, int k) const {
    const bool h = off >= CATAN_HOT_BEGIN && off < CATAN_HOT_END;
    return *reinterpret_cast<T*>((h ? base : cold) + (static_cast<size_t>(h ? off - CATAN_HOT_BEGIN : off) + static_cast<size_t>(k) * sizeof(T)) * CATAN_W +
                                 static_cast<size_t>(lane) * sizeof(T));
  }
  // raw access by the record offset of the ELEMENT (copies, clears): the pointer is chosen at run time
  template <class T>
  CATAN_MFN T& raw(int byte_off) const {
    const bool h = byte_off >= CATAN_HOT_BEGIN && byte_off < CATAN_HOT_END;
    return *reinterpret_cast<T*>((h ? base : cold) + static_cast<size_t>(h ? byte_off - CATAN_HOT_BEGIN : byte_off) * CATAN_W + static_cast<size_t>(lane) * sizeof(T));
  }
#define CATAN_F0(T, name) CATAN_MFN T& name() const { return at<T>(offsetof(GameRec, name), 0); }
#define CATAN_F1(T, name) CATAN_MFN T& name(int i) const { return at<T>(offsetof(GameRec, name), i); }
#define CATAN_F2(T, name, B) CATAN_MFN T& name(int i, int j) const { return at<T>(offsetof(GameRec, name), i * (B) + j); }
#define CATAN_F3(T, name, B, C_) CATAN_MFN T& name(int i, int j, int k) const { return at<T>(offsetof(GameRec, name), (i * (B) + j) * (C_) + k); }
  CATAN_F3(int16_t, est_min, 3, 5) CATAN_F3(int16_t, est_max, 3, 5) CATAN_F2(int16_t, vis, 5) CATAN_F2(uint16_t, prod, 13)
  CATAN_F0(uint32_t, rng_ctr) CATAN_F0(uint32_t, decision_ctr) CATAN_F0(uint32_t, episode_steps)
  CATAN_F0(uint16_t, actions_this_turn) CATAN_F0(uint16_t, turn)
  CATAN_F1(uint8_t, corner) CATAN_F1(uint8_t, edge) CATAN_F1(uint8_t, tile_res) CATAN_F1(uint8_t, tile_val)
  CATAN_F1(uint8_t, harbour_perm) CATAN_F0(uint8_t, robber_tile) CATAN_F2(uint8_t, res, 5) CATAN_F1(int8_t, vp)
  CATAN_F1(uint8_t, harbours) CATAN_F1(uint8_t, n_hidden) CATAN_F1(uint8_t, n_played) CATAN_F1(uint8_t, settlements_left)
  CATAN_F1(uint8_t, cities_left) CATAN_F1(uint8_t, init_settlements) CATAN_F1(uint8_t, init_roads) CATAN_F1(int8_t, second_corner)
  CATAN_F1(uint8_t, cur_longest_path) CATAN_F1(uint8_t, has_path_key) CATAN_F1(uint8_t, cur_army)
  CATAN_F2(uint8_t, hidden, 13) CATAN_F2(uint8_t, played, 13) CATAN_F1(uint8_t, bank) CATAN_F0(uint8_t, deck_n) CATAN_F1(uint8_t, deck)
  CATAN_F1(uint8_t, player_order) CATAN_F0(uint8_t, player_order_id) CATAN_F0(uint8_t, players_go)
  CATAN_F0(uint8_t, lr_holder) CATAN_F0(uint8_t, lr_count) CATAN_F0(uint8_t, la_holder) CATAN_F0(uint8_t, la_count)
  CATAN_F0(uint8_t, initial_phase) CATAN_F0(uint8_t, dice_rolled) CATAN_F0(uint8_t, played_dev) CATAN_F0(uint8_t, must_use_dev)
  CATAN_F0(uint8_t, rb_active) CATAN_F0(uint8_t, rb_count) CATAN_F0(uint8_t, can_move_robber) CATAN_F0(uint8_t, just_moved_robber)
  CATAN_F0(uint8_t, must_respond) CATAN_F0(uint8_t, need_discard) CATAN_F0(uint8_t, n_discard) CATAN_F1(uint8_t, discard_queue)
  CATAN_F0(uint8_t, trade_proposer) CATAN_F0(uint8_t, trade_target) CATAN_F0(uint8_t, n_give) CATAN_F1(uint8_t, give)
  CATAN_F0(uint8_t, n_recv) CATAN_F1(uint8_t, recv) CATAN_F0(uint8_t, die1) CATAN_F0(uint8_t, die2) CATAN_F0(uint8_t, trades_this_turn)
  CATAN_F1(uint8_t, bought) CATAN_F1(int8_t, curr_vps) CATAN_F0(uint8_t, winner) CATAN_F1(uint8_t, lr_dirty)
#undef CATAN_F0
#undef CATAN_F1
#undef CATAN_F2
#undef CATAN_F3
};

// view of game `i` of a record array stored as lane-interleaved chunks
CATAN_FN GameView game_view(uint8_t* recs, size_t i) {
  return GameView(recs + (i / CATAN_W) * CATAN_CHUNK_BYTES, static_cast<int>(i % CATAN_W));
}

// host side: one game between a chunked buffer and a plain GameRec (catan_export_state / catan_import_state)
static inline void chunk_get(const uint8_t* chunk, int lane, int W, GameRec& out) {
  // 1-, 2- and 4-byte fields interleave in units of their own size; GameRec's layout: 16-bit [0,384), uint32 [384,396),
  // uint16 [396,400), bytes from 400 on
  uint8_t* o = reinterpret_cast<uint8_t*>(&out);
  for (size_t off = 0; off < sizeof(GameRec);) {
    const size_t sz = off < offsetof(GameRec, rng_ctr) ? 2 : (off < offsetof(GameRec, actions_this_turn) ? 4 : (off < offsetof(GameRec, robber_tile) ? 2 : 1));
    memcpy(o + off, chunk + off * W + static_cast<size_t>(lane) * sz, sz);
    off += sz;
  }
}
static inline void chunk_put(uint8_t* chunk, int lane, int W, const GameRec& in) {
  const uint8_t* o = reinterpret_cast<const uint8_t*>(&in);
  for (size_t off = 0; off < sizeof(GameRec);) {
    const size_t sz = off < offsetof(GameRec, rng_ctr) ? 2 : (off < offsetof(GameRec, actions_this_turn) ? 4 : (off < offsetof(GameRec, robber_tile) ? 2 : 1));
    memcpy(chunk + off * W + static_cast<size_t>(lane) * sz, o + off, sz);
    off += sz;
  }
}
static_assert(offsetof(GameRec, est_min) == 0 && offsetof(GameRec, rng_ctr) == 384 && offsetof(GameRec, actions_this_turn) == 396 &&
              offsetof(GameRec, robber_tile) == 400, "chunk_get/chunk_put assume this field order");

// OR over the lanes of a group (host build: one lane)
#ifdef CATAN_DEVICE
CATAN_FN uint32_t group_or32(uint32_t v) { return __reduce_or_sync(0xffffffffu, v); }
#else
CATAN_FN uint32_t group_or32(uint32_t v) { return v; }
#endif
CATAN_FN uint64_t group_or64(uint64_t v) {
  return static_cast<uint64_t>(group_or32(static_cast<uint32_t>(v))) | (static_cast<uint64_t>(group_or32(static_cast<uint32_t>(v >> 32))) << 32);
}

// ------------------------------------------------------------------------------------------------
// bit tables derived from the topology (built once per block into shared memory)
// ------------------------------------------------------------------------------------------------
struct alignas(16) TopoX {
  uint64_t corner_nb[54];        // neighbouring corners of corner c
  uint64_t corner_edges_lo[54];  // edges 0..63 incident to corner c
  uint64_t tile_cmask[19];       // the six corners of tile t
  uint32_t corner_edges_hi[54];  // edges 64..71 incident to corner c (bit e - 64)
  uint32_t pad_[2];
};
CATAN_FN void build_topox(const Topo& T, TopoX& X, int tid, int nthreads) {
  for (int c = tid; c < 54; c += nthreads) {
    uint64_t nb = 0, lo = 0;
    uint32_t hi = 0;
    for (int k = 0; k < 3; ++k) {
      if (T.corner_neigh[c][k] >= 0) nb |= 1ull << T.corner_neigh[c][k];
      const int e = T.corner_neigh_edge[c][k];
      if (e >= 64) hi |= 1u << (e - 64);
      else if (e >= 0) lo |= 1ull << e;
    }
    X.corner_nb[c] = nb; X.corner_edges_lo[c] = lo; X.corner_edges_hi[c] = hi;
  }
  for (int t = tid; t < 19; t += nthreads) {
    uint64_t m = 0;
    for (int k = 0; k < 6; ++k) m |= 1ull << T.tile_corners[t][k];
    X.tile_cmask[t] = m;
  }
}

// ------------------------------------------------------------------------------------------------
// relative seating in registers (player.py:13-19): sp = seat of PlayerId p at bits [2p, 2p+2), op = PlayerId at seat s at
// bits [4s, 4s+4)
// ------------------------------------------------------------------------------------------------
struct Seats { uint32_t sp, op; };
CATAN_FN Seats load_seats(const GameView& g) {
  Seats s = {0u, 0u};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t p = g.player_order(i);
    s.op |= p << (4 * i);
    s.sp |= static_cast<uint32_t>(i) << (2 * p);
  }
  return s;
}
CATAN_FN int seat_of(Seats s, int pid) { return (s.sp >> (2 * pid)) & 3; }
CATAN_FN int pid_at_seat(Seats s, int seat) { return (s.op >> (4 * (seat & 3))) & 15; }
// relative label of b seen from a: 0 next, 1 next_next, 2 next_next_next (player_lookup); -1 if a == b
CATAN_FN int label_of(Seats s, int a, int b) { return ((seat_of(s, b) - seat_of(s, a) + 4) & 3) - 1; }
CATAN_FN int pid_at_label(Seats s, int a, int label) { return pid_at_seat(s, seat_of(s, a) + 1 + label); }

struct TCx {
  GameView g;
  const Topo* T;
  const TopoX* X;
  const catan_config_t* cfg;
  uint64_t seed, env_id;
  Seats s;
};
// address-space hints (CATAN_IN_SMEM).  ENCODE: functions that only the encode kernels call -- the whole chunk and both topology
// tables are staged.  RULES: functions that only the transition kernel calls -- the hot range of the chunk and the topology are
// staged, the cold fields may be at home in global memory (transition_kernel<DIRECT>).
#define CATAN_STAGED_ENCODE_G(g_) do { CATAN_IN_SMEM((g_).base); CATAN_IN_SMEM((g_).cold); } while (0)
#define CATAN_STAGED_ENCODE(cx_) do { CATAN_STAGED_ENCODE_G((cx_).g); CATAN_IN_SMEM((cx_).T); CATAN_IN_SMEM((cx_).X); } while (0)
#define CATAN_STAGED_RULES_G(g_) CATAN_IN_SMEM((g_).base)
#define CATAN_STAGED_RULES(cx_) do { CATAN_STAGED_RULES_G((cx_).g); CATAN_IN_SMEM((cx_).T); } while (0)

enum { CATAN_LR_ROAD = 0, CATAN_LR_SETTLE = 1 };
// what apply_action leaves for the follow-up passes of the same step
struct StepTmp {             // (kept small: 32 of them sit next to the staged chunk in the transition kernel's shared memory)
  EstReq est[2];
  uint8_t dice_T[4][5];      // clip bound of the (r, player) belief update of this roll (a hand total)
  uint8_t alloc[5][4];       // dice payout [r][player index]      (game.py:153-167)
  uint8_t mono_T[4];
  uint8_t mono_lost[4];
  uint8_t n_est, est_special, granted, dice_roll;
  uint8_t mono_pid, mono_res, lr_pid, err;
  uint8_t acted_pid, act_type, roll_info, follow;   // follow: dice payout / belief updates are pending (t_followups_group)
  uint8_t lr_kind, lr_loc;   // what triggered the longest-road update of lr_pid: CATAN_LR_ROAD + edge (0xff: none), CATAN_LR_SETTLE + corner
  Seats s;                   // seating of the game (for the follow-up pass)
};

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
CATAN_FN int t_hand_total(const GameView& g, int pid) {
  const int p = pid - 1;
  return g.res(p, 0) + g.res(p, 1) + g.res(p, 2) + g.res(p, 3) + g.res(p, 4);
}
CATAN_FN int t_current_actor(const GameView& g) {   // game_manager.py:152-159 / wrapper.py:53-58
  return g.need_discard() ? g.discard_queue(0) : (g.must_respond() ? g.trade_target() : g.players_go());
}
CATAN_FN int t_best_exchange_rate(const GameView& g, int pid, int r) {   // wrapper.py:428-438
  const int h = g.harbours(pid - 1);
  return (h >> (r + 1)) & 1 ? 2 : ((h & 1) ? 3 : 4);
}
// the ordered card lists hold two cards per byte
CATAN_FN int t_hidden_at(const GameView& g, int p, int i) { return (g.hidden(p, i >> 1) >> (4 * (i & 1))) & 15; }
CATAN_FN int t_played_at(const GameView& g, int p, int i) { return (g.played(p, i >> 1) >> (4 * (i & 1))) & 15; }
CATAN_FN void t_hidden_set(const GameView& g, int p, int i, int card) {
  uint8_t& b = g.hidden(p, i >> 1);
  b = static_cast<uint8_t>((b & ~(15u << (4 * (i & 1)))) | (static_cast<uint32_t>(card) << (4 * (i & 1))));
}
CATAN_FN void t_played_set(const GameView& g, int p, int i, int card) {
  uint8_t& b = g.played(p, i >> 1);
  b = static_cast<uint8_t>((b & ~(15u << (4 * (i & 1)))) | (static_cast<uint32_t>(card) << (4 * (i & 1))));
}
// number of cards of every kind in a player's hidden list, 6 bits per kind
CATAN_FN uint32_t t_hidden_counts(const GameView& g, int p) {
  uint32_t k = 0;
  const int n = g.n_hidden(p);
  CATAN_NO_UNROLL
  for (int i = 0; i < n; ++i) k += 1u << (6 * t_hidden_at(g, p, i));
  return k;
}
CATAN_FN int t_count_played(const GameView& g, int p, int card) {
  int k = 0;
  const int n = g.n_played(p);
  CATAN_NO_UNROLL
  for (int i = 0; i < n; ++i) k += t_played_at(g, p, i) == card;
  return k;
}
// the production cache (GameRec::prod): `w` more production for player index p from every tile around corner c
CATAN_FN void t_prod_add(const GameView& g, const Topo& T, int p, int c, int w) {
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int tl = T.corner_tiles[c][k];
    if (tl < 0) continue;
    const int tr = g.tile_res(tl), val = g.tile_val(tl);
    if (tr == 0 || val == 7) continue;
    const int j = CATAN_OBS_RES_SLOT(tr - 1) * 10 + CATAN_OBS_NUM_SLOT(val);
    g.prod(p, j >> 2) = static_cast<uint16_t>(g.prod(p, j >> 2) + (w << (4 * (j & 3))));
  }
}

CATAN_FN uint32_t t_rng_next(TCx& cx) {   // next word of the game stream
  const uint32_t d = cx.g.rng_ctr();
  cx.g.rng_ctr() = d + 1;
  return philox4x32_word(d >> 2, CATAN_STREAM_GAME, static_cast<uint32_t>(cx.env_id), static_cast<uint32_t>(cx.env_id >> 32),
                         static_cast<uint32_t>(cx.seed), static_cast<uint32_t>(cx.seed >> 32), static_cast<int>(d & 3));
}
CATAN_FN int t_rng_bounded(TCx& cx, int n) { return static_cast<int>(mulhi32(t_rng_next(cx), static_cast<uint32_t>(n))); }

// ------------------------------------------------------------------------------------------------
// placement predicates for ONE location (corner.py:24-39, edge.py:23-42); the mask encoder uses bit boards instead
// ------------------------------------------------------------------------------------------------
CATAN_FN_NOINLINE bool t_can_place_settlement(const GameView& g_, const Topo& T, int c, int pid, bool initial) {
  const GameView g = g_;
  CATAN_STAGED_RULES_G(g); CATAN_IN_SMEM(&T);
  if (g.corner(c)) return false;
  bool own_road = false;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int nb = T.corner_neigh[c][k];
    if (nb < 0) continue;
    if (g.corner(nb)) return false;
    own_road |= g.edge(T.corner_neigh_edge[c][k]) == pid;
  }
  return initial || own_road;
}

CATAN_FN_NOINLINE bool t_can_place_road(const GameView& g_, const Topo& T, int e, int pid, bool after_second, int second_corner) {
  const GameView g = g_;
  CATAN_STAGED_RULES_G(g); CATAN_IN_SMEM(&T);
  if (g.edge(e)) return false;
  const int c1 = T.edge_corners[e][0], c2 = T.edge_corners[e][1];
  if (after_second) return c1 == second_corner || c2 == second_corner;
  const uint8_t b1 = g.corner(c1), b2 = g.corner(c2);
  if ((b1 && (b1 >> 2) == pid) || (b2 && (b2 >> 2) == pid)) return true;
  bool ok = false;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int e1 = T.corner_neigh_edge[c1][k], e2 = T.corner_neigh_edge[c2][k];
    ok |= (e1 >= 0 && !b1 && g.edge(e1) == pid);
    ok |= (e2 >= 0 && !b2 && g.edge(e2) == pid);
  }
  return ok;
}

// ------------------------------------------------------------------------------------------------
// translate (wrapper.py:114-166, :414-486) and validate (game.py:264-525)
// ------------------------------------------------------------------------------------------------
CATAN_FN_NOINLINE int t_translate_action(const TCx& cx_, const int32_t* a, Act& t) {
  TCx cx = cx_;   // (a private copy: the address-space hints below then hold for every access, whatever the stores in between)
  CATAN_IN_LOCAL(&t);
  CATAN_STAGED_RULES(cx);
  const GameView& g = cx.g;
  memset(&t, 0, sizeof(Act));
  const int type = a[CATAN_A_TYPE], pg = g.players_go();
  t.type = static_cast<int8_t>(type);
  switch (type) {
    case CATAN_ACT_PLACE_SETTLEMENT:
    case CATAN_ACT_UPGRADE_CITY: {
      const int v = a[CATAN_A_CORNER];
      if (v < 0 || v >= 54) return CATAN_ERR_BAD_HEAD_VALUE;
      t.corner = static_cast<int8_t>(v);
      return 0;
    }
    case CATAN_ACT_PLACE_ROAD: {
      const int v = a[CATAN_A_EDGE];
      if (v < 0 || v > 72) return CATAN_ERR_BAD_HEAD_VALUE;
      t.edge = static_cast<int8_t>(v == 72 ? -1 : v);
      return 0;
    }
    case CATAN_ACT_MOVE_ROBBER: {
      const int v = a[CATAN_A_TILE];
      if (v < 0 || v >= 19) return CATAN_ERR_BAD_HEAD_VALUE;
      t.tile = static_cast<int8_t>(v);
      return 0;
    }
    case CATAN_ACT_STEAL:
    case CATAN_ACT_PROPOSE_TRADE: {
      const int pl = a[CATAN_A_PLAYER];
      if (pl < 0 || pl > 2) return CATAN_ERR_BAD_HEAD_VALUE;
      t.target_pid = static_cast<int8_t>(pid_at_label(cx.s, pg, pl));
      if (type == CATAN_ACT_STEAL) return 0;
      for (int k = 0; k < 4; ++k) {                                  // wrapper.py:451-466: 0 ends the list
        const int v = a[CATAN_A_GIVE + k];
        if (v == 0) break;
        if (v < 0 || v > 5) return CATAN_ERR_BAD_HEAD_VALUE;
        t.give[t.n_give++] = static_cast<int8_t>(v - 1);
      }
      for (int k = 0; k < 4; ++k) {
        const int v = a[CATAN_A_RECV + k];
        if (v == 0) break;
        if (v < 0 || v > 5) return CATAN_ERR_BAD_HEAD_VALUE;
        t.recv[t.n_recv++] = static_cast<int8_t>(v - 1);
      }
      return 0;
    }
    case CATAN_ACT_PLAY_DEV: {
      const int card = a[CATAN_A_CARD];
      if (card < 0 || card > 4) return CATAN_ERR_BAD_HEAD_VALUE;
      t.card = static_cast<int8_t>(card);
      if (card == CATAN_DEV_MONOPOLY || card == CATAN_DEV_YOP) {
        const int v = a[CATAN_A_RES_A];
        if (v < 0 || v > 4) return CATAN_ERR_BAD_HEAD_VALUE;
        t.res_a = static_cast<int8_t>(v);
      }
      if (card == CATAN_DEV_YOP) {
        const int v = a[CATAN_A_RES_B];
        if (v < 0 || v > 4) return CATAN_ERR_BAD_HEAD_VALUE;
        t.res_b = static_cast<int8_t>(v);
      }
      return 0;
    }
    case CATAN_ACT_EXCHANGE: {
      const int va = a[CATAN_A_RES_A], vb = a[CATAN_A_RES_B];
      if (va < 0 || va > 4 || vb < 0 || vb > 4) return CATAN_ERR_BAD_HEAD_VALUE;
      t.res_a = static_cast<int8_t>(va);
      t.res_b = static_cast<int8_t>(vb);
      t.rate = static_cast<int8_t>(t_best_exchange_rate(g, pg, va));
      return 0;
    }
    case CATAN_ACT_RESPOND: {
      const int v = a[CATAN_A_ACCEPT];
      if (v < 0 || v > 1) return CATAN_ERR_BAD_HEAD_VALUE;
      t.accept = static_cast<int8_t>(v);
      return 0;
    }
    case CATAN_ACT_DISCARD: {
      const int v = a[CATAN_A_DISCARD];
      if (v < 0 || v > 4) return CATAN_ERR_BAD_HEAD_VALUE;
      t.discard = static_cast<int8_t>(v);
      return 0;
    }
    case CATAN_ACT_BUY_DEV:
    case CATAN_ACT_ROLL_DICE:
    case CATAN_ACT_END_TURN:
      return 0;
    default:
      return CATAN_ERR_BAD_TYPE;
  }
}

CATAN_FN_NOINLINE int t_validate_action(const TCx& cx_, const Act& t) {
  TCx cx = cx_;   // (a private copy: the address-space hints below then hold for every access, whatever the stores in between)
  CATAN_IN_LOCAL(&t);
  CATAN_STAGED_RULES(cx);
  const GameView& g = cx.g;
  const Topo& T = *cx.T;
  const int pid = g.players_go(), p = pid - 1;
  if (g.need_discard()) {                                            // game.py:279-300
    if (t.type != CATAN_ACT_DISCARD) return CATAN_ERR_PHASE;
    const int d = g.discard_queue(0);
    if (t_hand_total(g, d) <= 7) return CATAN_ERR_PHASE;
    return g.res(d - 1, t.discard) > 0 ? 0 : CATAN_ERR_BAD_RESOURCE;
  }
  if (t.type == CATAN_ACT_DISCARD) return CATAN_ERR_PHASE;           // game.py:301-303
  const bool must_respond = g.must_respond(), initial = g.initial_phase(), rolled = g.dice_rolled(), must_use = g.must_use_dev(),
             just_moved = g.just_moved_robber();
  // the common guard of most main-phase actions (must have rolled, nothing pending)
  const bool blocked_main = must_respond || initial || !rolled || must_use || just_moved;
  switch (t.type) {
    case CATAN_ACT_PLACE_SETTLEMENT:                                 // game.py:305-323
      if (must_respond || (!rolled && !initial) || must_use || just_moved) return CATAN_ERR_PHASE;
      if (initial || (g.settlements_left(p) > 0 && g.res(p, WHEAT) && g.res(p, WOOD) && g.res(p, BRICK) && g.res(p, SHEEP))) {
        if (t_can_place_settlement(g, T, t.corner, pid, initial)) {
          if (!initial) return 0;
          return (g.init_settlements(p) == 0 || (g.init_settlements(p) == 1 && g.init_roads(p) == 1)) ? 0 : CATAN_ERR_BAD_LOCATION;
        }
      }
      return CATAN_ERR_CANNOT_AFFORD;
    case CATAN_ACT_PLACE_ROAD:                                       // game.py:324-357
      if (g.rb_active()) {
        if (t.edge < 0) return 0;
        return t_can_place_road(g, T, t.edge, pid, false, 0) ? 0 : CATAN_ERR_BAD_LOCATION;
      }
      if (must_respond || (!rolled && !initial) || must_use || just_moved) return CATAN_ERR_PHASE;
      if (!(initial || (g.res(p, WOOD) && g.res(p, BRICK)))) return CATAN_ERR_CANNOT_AFFORD;
      if (t.edge < 0) return CATAN_ERR_BAD_LOCATION;
      if (!t_can_place_road(g, T, t.edge, pid, false, 0)) return CATAN_ERR_BAD_LOCATION;
      if (!initial) return 0;
      if (g.init_settlements(p) == 1 && g.init_roads(p) == 0) return 0;
      if (g.init_settlements(p) == 2 && g.init_roads(p) == 1)
        return t_can_place_road(g, T, t.edge, pid, true, g.second_corner(p)) ? 0 : CATAN_ERR_BAD_LOCATION;
      return CATAN_ERR_BAD_LOCATION;
    case CATAN_ACT_UPGRADE_CITY:                                     // game.py:358-376
      if (blocked_main) return CATAN_ERR_PHASE;
      if (g.cities_left(p) > 0 && g.res(p, WHEAT) > 1 && g.res(p, ORE) > 2) {
        const uint8_t b = g.corner(t.corner);
        if ((b & 3) != 1) return CATAN_ERR_BAD_LOCATION;
        if ((b >> 2) == pid) return 0;
      }
      return CATAN_ERR_CANNOT_AFFORD;
    case CATAN_ACT_BUY_DEV:                                          // game.py:377-393
      if (blocked_main) return CATAN_ERR_PHASE;
      if (g.res(p, WHEAT) && g.res(p, SHEEP) && g.res(p, ORE)) return g.deck_n() > 0 ? 0 : CATAN_ERR_BAD_CARD;
      return CATAN_ERR_CANNOT_AFFORD;
    case CATAN_ACT_PLAY_DEV: {                                       // game.py:394-415
      if (must_respond || g.played_dev() || initial || just_moved) return CATAN_ERR_PHASE;
      const int k = (t_hidden_counts(g, p) >> (6 * t.card)) & 63;
      return (k > 0 && k != g.bought(t.card)) ? 0 : CATAN_ERR_BAD_CARD;
    }
    case CATAN_ACT_EXCHANGE:                                         // game.py:416-443
      if (blocked_main) return CATAN_ERR_PHASE;
      if (g.res(p, t.res_a) < t.rate) return CATAN_ERR_CANNOT_AFFORD;
      return g.bank(t.res_b) > 0 ? 0 : CATAN_ERR_BAD_RESOURCE;
    case CATAN_ACT_PROPOSE_TRADE: {                                  // game.py:444-466
      if (blocked_main) return CATAN_ERR_PHASE;
      uint32_t cnt = 0;                                              // 4 bits per resource
      for (int k = 0; k < t.n_give; ++k) cnt += 1u << (4 * t.give[k]);
      for (int r = 0; r < 5; ++r) if (g.res(p, r) < static_cast<int>((cnt >> (4 * r)) & 15)) return CATAN_ERR_CANNOT_AFFORD;
      return 0;
    }
    case CATAN_ACT_RESPOND: {                                        // game.py:467-482
      if (!must_respond) return CATAN_ERR_PHASE;
      if (t.accept == 1) return 0;                                   // head value 1 == "reject" (wrapper.py:157-160)
      uint32_t cnt = 0;
      const int nr = g.n_recv(), tt = g.trade_target() - 1;
      for (int k = 0; k < nr; ++k) cnt += 1u << (4 * (g.recv(k) - 1));
      for (int r = 0; r < 5; ++r) if (g.res(tt, r) < static_cast<int>((cnt >> (4 * r)) & 15)) return CATAN_ERR_CANNOT_AFFORD;
      return 0;
    }
    case CATAN_ACT_MOVE_ROBBER:                                      // game.py:483-490
      return (must_respond || must_use || !g.can_move_robber()) ? CATAN_ERR_PHASE : 0;
    case CATAN_ACT_ROLL_DICE:                                        // game.py:491-500
      return (must_respond || initial || rolled || just_moved) ? CATAN_ERR_PHASE : 0;
    case CATAN_ACT_END_TURN:                                         // game.py:501-512
      return blocked_main ? CATAN_ERR_PHASE : 0;
    case CATAN_ACT_STEAL: {                                          // game.py:513-525
      if (must_respond || !just_moved) return CATAN_ERR_PHASE;
      const int rt = g.robber_tile();
      for (int k = 0; k < 6; ++k) {
        const uint8_t b = g.corner(T.tile_corners[rt][k]);
        if (b && (b >> 2) == t.target_pid) return 0;
      }
      return CATAN_ERR_BAD_TARGET;
    }
  }
  return CATAN_ERR_BAD_TYPE;
}

// ------------------------------------------------------------------------------------------------
// apply_action, scalar part (game.py:527-815).  Belief updates, the dice payout and the longest-road search are posted
// to `tmp` and executed afterwards.
// ------------------------------------------------------------------------------------------------
CATAN_FN void t_pay(const GameView& g, int p, int r, int n) {   // hand -n, visible floor 0, bank +n (game.py:197-208 etc.)
  g.res(p, r) = static_cast<uint8_t>(g.res(p, r) - n);
  const int v = g.vis(p, r) - n;
  g.vis(p, r) = static_cast<int16_t>(v > 0 ? v : 0);
  g.bank(r) = static_cast<uint8_t>(g.bank(r) + n);
}

CATAN_FN EstReq& t_post_est(TCx& cx, StepTmp& tmp, int owner, int thief) {
  EstReq& q = tmp.est[tmp.n_est++];
  memset(&q, 0, sizeof(EstReq));
  q.owner = static_cast<uint8_t>(owner);
  q.thief = static_cast<uint8_t>(thief);
  q.T_o = static_cast<int16_t>(t_hand_total(cx.g, owner));
  q.T_t = static_cast<int16_t>(thief ? t_hand_total(cx.g, thief) : 0);
  return q;
}

CATAN_FN void t_advance_seat(const GameView& g, bool left) {   // game.py:253-262
  const int id = (g.player_order_id() + (left ? 3 : 1)) & 3;
  g.player_order_id() = static_cast<uint8_t>(id);
  g.players_go() = g.player_order(id);
}

CATAN_FN_NOINLINE void t_update_largest_army(const GameView& g_) {   // game.py:817-841
  const GameView g = g_;
  CATAN_STAGED_RULES_G(g);
  int max_count = 0, cp = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = (0x3412 >> (4 * i)) & 15;                    // order Blue, White, Red, Orange (game.py:820)
    const int k = t_count_played(g, q - 1, CATAN_DEV_KNIGHT);
    g.cur_army(q - 1) = static_cast<uint8_t>(k);
    if (k >= 3 && k > max_count) { max_count = k; cp = q; }
  }
  if (!cp) return;
  const int holder = g.la_holder();
  if (!holder) { g.la_holder() = static_cast<uint8_t>(cp); g.la_count() = static_cast<uint8_t>(max_count); g.vp(cp - 1) += 2; }
  else if (holder == cp) g.la_count() = static_cast<uint8_t>(max_count);
  else if (max_count > g.la_count()) {
    g.vp(holder - 1) -= 2;
    g.la_holder() = static_cast<uint8_t>(cp); g.la_count() = static_cast<uint8_t>(max_count);
    g.vp(cp - 1) += 2;
  }
}

CATAN_FN_NOINLINE void t_apply_scalar(TCx& cx_, StepTmp& tmp, const Act& t) {
  TCx cx = cx_;   // (a private copy: the address-space hints below then hold for every access, whatever the stores in between)
  CATAN_IN_LOCAL(&t); CATAN_IN_SMEM(&tmp);
  CATAN_STAGED_RULES(cx);
  const GameView& g = cx.g;
  const Topo& T = *cx.T;
  const int pid = g.players_go(), p = pid - 1;
  switch (t.type) {
    case CATAN_ACT_PLACE_SETTLEMENT: {                               // game.py:530-555, :195-212
      const int c = t.corner;
      const bool initial = g.initial_phase();
      if (!initial) { t_pay(g, p, WHEAT, 1); t_pay(g, p, SHEEP, 1); t_pay(g, p, WOOD, 1); t_pay(g, p, BRICK, 1); }
      g.corner(c) = static_cast<uint8_t>((pid << 2) | 1);
      t_prod_add(g, T, p, c, 1);
#pragma unroll
      for (int k = 0; k < 3; ++k) {                                  // the corner now cuts the other players' roads through it
        const int e = T.corner_neigh_edge[c][k];
        if (e < 0) continue;
        const int o = g.edge(e);
        if (o && o != pid) g.lr_dirty(o - 1) = 1;
      }
      const int slot = T.corner_harbour_slot[c];                     // board.py:182-183
      if (slot >= 0) {
        const int hres = T.harbour_res[g.harbour_perm(slot)];
        g.harbours(p) |= static_cast<uint8_t>(1u << hres);           // bit 0 = generic, bit Resource = 2:1
      }
      g.settlements_left(p) -= 1;
      g.vp(p) += 1;
      if (initial) {
        const int ns = g.init_settlements(p) + 1;
        g.init_settlements(p) = static_cast<uint8_t>(ns);
        if (ns == 2) {
          uint32_t gain = 0;                                         // 4 bits per resource
          for (int k = 0; k < 3; ++k) {
            const int tl = T.corner_tiles[c][k];
            if (tl < 0) continue;
            const int tr = g.tile_res(tl);
            if (tr == 0) continue;
            const int r = tr - 1;
            g.res(p, r) += 1; g.vis(p, r) += 1; g.bank(r) -= 1; gain += 1u << (4 * r);
          }
          EstReq& q = t_post_est(cx, tmp, pid, 0);
          for (int r = 0; r < 5; ++r) if ((gain >> (4 * r)) & 15) est_set(q, r, (gain >> (4 * r)) & 15);
          g.second_corner(p) = static_cast<int8_t>(c);
        }
      } else {
        EstReq& q = t_post_est(cx, tmp, pid, 0);
        est_set(q, BRICK, -1); est_set(q, WOOD, -1); est_set(q, WHEAT, -1); est_set(q, SHEEP, -1);
        if (g.lr_holder()) { tmp.lr_pid = g.lr_holder(); tmp.lr_kind = CATAN_LR_SETTLE; tmp.lr_loc = static_cast<uint8_t>(c); }   // game.py:552-553
      }
      break;
    }
    case CATAN_ACT_PLACE_ROAD: {                                     // game.py:556-597, :222-232
      bool final_init = false;
      const bool initial = g.initial_phase(), rb = g.rb_active();
      if (t.edge >= 0) {
        if (!initial && !rb) { t_pay(g, p, WOOD, 1); t_pay(g, p, BRICK, 1); }
        g.edge(t.edge) = static_cast<uint8_t>(pid);
        if (initial) {
          g.init_roads(p) += 1;
          int first = 0, second = 0;
          for (int q = 0; q < 4; ++q) { const int ns = g.init_settlements(q); first += ns >= 1; second += ns == 2; }
          if (first < 4) t_advance_seat(g, false);
          else if (second == 0) { /* last seat places twice in a row */ }
          else if (second < 4) t_advance_seat(g, true);
          else { g.initial_phase() = 0; final_init = true; }
        }
      }
      tmp.lr_pid = static_cast<uint8_t>(pid);                        // game.py:585 (also for the dummy edge)
      tmp.lr_kind = CATAN_LR_ROAD; tmp.lr_loc = static_cast<uint8_t>(t.edge >= 0 ? t.edge : 0xff);
      if (rb) {
        const int n = g.rb_count() + 1;
        if (n >= 2) { g.rb_active() = 0; g.rb_count() = 0; g.must_use_dev() = 0; }
        else g.rb_count() = static_cast<uint8_t>(n);
      } else if (!initial) {
        EstReq& q = t_post_est(cx, tmp, pid, 0);
        est_set(q, BRICK, -1); est_set(q, WOOD, -1);
      }
      break;
    }
    case CATAN_ACT_UPGRADE_CITY: {                                   // game.py:598-604, :240-251
      t_pay(g, p, WHEAT, 2); t_pay(g, p, ORE, 3);
      g.corner(t.corner) = static_cast<uint8_t>((pid << 2) | 2);
      t_prod_add(g, T, p, t.corner, 1);
      g.vp(p) += 1; g.cities_left(p) -= 1; g.settlements_left(p) += 1;
      EstReq& q = t_post_est(cx, tmp, pid, 0);
      est_set(q, ORE, -3); est_set(q, WHEAT, -2);
      break;
    }
    case CATAN_ACT_ROLL_DICE: {                                      // game.py:605-611, :138-150
      const int d1 = 1 + t_rng_bounded(cx, 6), d2 = 1 + t_rng_bounded(cx, 6);
      g.die1() = static_cast<uint8_t>(d1);
      g.die2() = static_cast<uint8_t>(d2);
      const int roll = d1 + d2;
      tmp.roll_info = static_cast<uint8_t>(roll);
      g.dice_rolled() = 1;
      if (roll == 7) {
        g.can_move_robber() = 1;
        int nd = g.n_discard();
        for (int i = 0; i < 4; ++i) {
          const int q = pid_at_seat(cx.s, i);
          if (t_hand_total(g, q) > 7) { g.need_discard() = 1; g.discard_queue(nd++) = static_cast<uint8_t>(q); }
        }
        g.n_discard() = static_cast<uint8_t>(nd);
      } else {
        tmp.dice_roll = static_cast<uint8_t>(roll);                  // payout + beliefs: t_dice_payout()
      }
      break;
    }
    case CATAN_ACT_END_TURN:                                         // game.py:612-622
      g.can_move_robber() = 0; g.dice_rolled() = 0; g.played_dev() = 0;
      t_advance_seat(g, false);
      g.turn() += 1;
      for (int c = 0; c < 5; ++c) g.bought(c) = 0;
      g.trades_this_turn() = 0; g.actions_this_turn() = 0;
      break;
    case CATAN_ACT_MOVE_ROBBER: {                                    // game.py:623-634
      g.robber_tile() = static_cast<uint8_t>(t.tile);
      g.can_move_robber() = 0;
      bool jm = false;
      for (int k = 0; k < 6; ++k) {
        const uint8_t b = g.corner(T.tile_corners[t.tile][k]);
        jm |= b && (b >> 2) != pid;
      }
      if (jm) g.just_moved_robber() = 1;
      break;
    }
    case CATAN_ACT_STEAL: {                                          // game.py:635-652
      const int v = t.target_pid;
      const int n = t_hand_total(g, v);
      if (n > 0) {
        int idx = t_rng_bounded(cx, n), r = BRICK;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
          const int ri = (0x23140 >> (4 * i)) & 15;            // Brick, Wheat, Wood, Sheep, Ore (game.py:638)
          const int cnt = g.res(v - 1, ri);
          if (idx >= 0 && idx < cnt) { r = ri; idx = -1; }
          else if (idx >= 0) idx -= cnt;
        }
        g.res(p, r) += 1; g.res(v - 1, r) -= 1;
        for (int q = 0; q < 5; ++q) if (g.vis(v - 1, q) > 0) g.vis(v - 1, q) -= 1;
        EstReq& rq = t_post_est(cx, tmp, v, pid);
        est_set(rq, r, -1);
      }
      g.just_moved_robber() = 0;
      break;
    }
    case CATAN_ACT_PLAY_DEV: {                                       // game.py:653-693
      const int n = g.n_hidden(p);
      int at = 0;
      while (at < n && t_hidden_at(g, p, at) != t.card) ++at;
      if (at < n) {                                                  // (always true for a validated action)
        for (int i = at; i + 1 < n; ++i) t_hidden_set(g, p, i, t_hidden_at(g, p, i + 1));
        t_hidden_set(g, p, n - 1, 0);
        g.n_hidden(p) = static_cast<uint8_t>(n - 1);
      }
      const int np = g.n_played(p);
      if (np < 25) { t_played_set(g, p, np, t.card); g.n_played(p) = static_cast<uint8_t>(np + 1); }
      g.played_dev() = 1;
      if (t.card == CATAN_DEV_VP) g.vp(p) += 1;
      else if (t.card == CATAN_DEV_KNIGHT) { g.can_move_robber() = 1; t_update_largest_army(g); }
      else if (t.card == CATAN_DEV_ROADBUILDING) { g.rb_active() = 1; g.rb_count() = 0; g.must_use_dev() = 1; }
      else if (t.card == CATAN_DEV_MONOPOLY) {
        const int r = t.res_a;
        tmp.est_special = EST_SPECIAL_MONOPOLY;
        tmp.mono_pid = static_cast<uint8_t>(pid); tmp.mono_res = static_cast<uint8_t>(r);
        for (int o = 0; o < 4; ++o) {
          tmp.mono_lost[o] = 0;
          if (o == p) continue;
          const int cnt = g.res(o, r);
          g.res(o, r) = 0; g.vis(o, r) = 0;
          g.res(p, r) = static_cast<uint8_t>(g.res(p, r) + cnt); g.vis(p, r) = static_cast<int16_t>(g.vis(p, r) + cnt);
          tmp.mono_lost[o] = static_cast<uint8_t>(cnt);
        }
        for (int o = 0; o < 4; ++o) tmp.mono_T[o] = static_cast<uint8_t>(t_hand_total(g, o + 1));
      } else {                                                       // Year of Plenty
        for (int i = 0; i < 2; ++i) {
          const int r = i == 0 ? t.res_a : t.res_b;
          if (g.bank(r) > 0) {
            g.bank(r) -= 1; g.res(p, r) += 1; g.vis(p, r) += 1;
            EstReq& q = t_post_est(cx, tmp, pid, 0);
            est_set(q, r, 1);
          }
        }
      }
      break;
    }
    case CATAN_ACT_BUY_DEV: {                                        // game.py:694-710
      t_pay(g, p, SHEEP, 1); t_pay(g, p, ORE, 1); t_pay(g, p, WHEAT, 1);
      EstReq& q = t_post_est(cx, tmp, pid, 0);
      est_set(q, SHEEP, -1); est_set(q, ORE, -1); est_set(q, WHEAT, -1);
      const int dn = g.deck_n(), nh = g.n_hidden(p);
      if (dn > 0 && nh < 25) {                                       // (always true for a validated action)
        const int card = g.deck(dn - 1);                             // deque.pop(): right end
        g.deck(dn - 1) = 0; g.deck_n() = static_cast<uint8_t>(dn - 1);
        t_hidden_set(g, p, nh, card); g.n_hidden(p) = static_cast<uint8_t>(nh + 1);
        g.bought(card) += 1;
      }
      break;
    }
    case CATAN_ACT_EXCHANGE: {                                       // game.py:711-734
      const int d = t.res_b, tr = t.res_a, rate = t.rate;
      g.res(p, d) += 1; g.vis(p, d) += 1;
      g.res(p, tr) = static_cast<uint8_t>(g.res(p, tr) - rate);
      const int v = g.vis(p, tr) - rate;
      g.vis(p, tr) = static_cast<int16_t>(v > 0 ? v : 0);
      g.bank(tr) = static_cast<uint8_t>(g.bank(tr) + rate); g.bank(d) -= 1;
      EstReq& q = t_post_est(cx, tmp, pid, 0);
      if (d == tr) est_set(q, d, 1 - rate);
      else { est_set(q, d, 1); est_set(q, tr, -rate); }
      break;
    }
    case CATAN_ACT_PROPOSE_TRADE:                                    // game.py:735-750
      g.must_respond() = 1;
      g.trade_proposer() = static_cast<uint8_t>(pid); g.trade_target() = static_cast<uint8_t>(t.target_pid);
      g.n_give() = static_cast<uint8_t>(t.n_give); g.n_recv() = static_cast<uint8_t>(t.n_recv);
      for (int k = 0; k < 4; ++k) {
        g.give(k) = static_cast<uint8_t>(k < t.n_give ? t.give[k] + 1 : 0);
        g.recv(k) = static_cast<uint8_t>(k < t.n_recv ? t.recv[k] + 1 : 0);
      }
      g.trades_this_turn() += 1;
      break;
    case CATAN_ACT_RESPOND: {                                        // game.py:751-784
      if (t.accept == 0) {
        const int p1 = g.trade_proposer() - 1, p2 = g.trade_target() - 1;
        const int ng = g.n_give(), nr = g.n_recv();
        int8_t d1[5] = {0, 0, 0, 0, 0};
        uint8_t touched = 0;
        for (int k = 0; k < ng; ++k) {
          const int r = g.give(k) - 1;
          g.res(p1, r) -= 1; if (g.vis(p1, r) > 0) g.vis(p1, r) -= 1;
          g.res(p2, r) += 1; g.vis(p2, r) += 1;
          d1[r] -= 1; touched |= static_cast<uint8_t>(1u << r);
        }
        for (int k = 0; k < nr; ++k) {
          const int r = g.recv(k) - 1;
          g.res(p1, r) += 1; g.vis(p1, r) += 1;
          g.res(p2, r) -= 1; if (g.vis(p2, r) > 0) g.vis(p2, r) -= 1;
          d1[r] += 1; touched |= static_cast<uint8_t>(1u << r);
        }
        EstReq& q1 = t_post_est(cx, tmp, p1 + 1, 0);
        EstReq& q2 = t_post_est(cx, tmp, p2 + 1, 0);
        for (int r = 0; r < 5; ++r) { q1.delta[r] = d1[r]; q2.delta[r] = static_cast<int8_t>(-d1[r]); }
        q1.touched = touched; q2.touched = touched;
      }
      g.must_respond() = 0;
      g.trade_proposer() = 0; g.trade_target() = 0; g.n_give() = 0; g.n_recv() = 0;
      for (int k = 0; k < 4; ++k) { g.give(k) = 0; g.recv(k) = 0; }
      break;
    }
    case CATAN_ACT_DISCARD: {                                        // game.py:785-807
      const int d = g.discard_queue(0), r = t.discard;
      g.res(d - 1, r) -= 1; g.bank(r) += 1;
      EstReq& q = t_post_est(cx, tmp, d, 0);
      est_set(q, r, -1);
      if (t_hand_total(g, d) <= 7) {
        for (int i = 0; i < 3; ++i) g.discard_queue(i) = g.discard_queue(i + 1);
        g.discard_queue(3) = 0;
        const int nd = g.n_discard() - 1;
        g.n_discard() = static_cast<uint8_t>(nd);
        if (nd == 0) g.need_discard() = 0;
      }
      break;
    }
  }
  if (t.type != CATAN_ACT_RESPOND && t.type != CATAN_ACT_END_TURN && t.type != CATAN_ACT_DISCARD)
    g.actions_this_turn() += 1;                                      // game.py:809-810
}

// ------------------------------------------------------------------------------------------------
// Follow-ups of a transition, executed by a GROUP of `nl` lanes on ONE game (a warp on the device, one lane in the host
// build): the dice payout (game.py:151-175) and the belief updates.  They are the data-parallel part of apply_action
// (19 tiles x 6 corners, 60 belief entries), so the thread that ran the scalar part only posts them to `tmp`, which
// lives in memory that the whole group sees (shared memory on the device).
//   generic  : update_player_resource_estimates (game.py:921-971)
//   dice     : the 20 calls of one roll folded into one pass (game.py:170-175; Q4)
//   monopoly : update_resource_estimates_monopoly (game.py:973-1010)
// ------------------------------------------------------------------------------------------------
#ifdef CATAN_DEVICE
#define CATAN_GROUP_SYNC() __syncwarp()
#else
#define CATAN_GROUP_SYNC() ((void)0)
#endif

CATAN_FN void t_dice_payout_group(const GameView& g, const Topo& T, StepTmp& tmp, int lane, int nl) {
  const int roll = tmp.dice_roll, robber = g.robber_tile();
  for (int i = lane; i < 20; i += nl) (&tmp.alloc[0][0])[i] = 0;
  CATAN_GROUP_SYNC();
  for (int t = lane; t < 19; t += nl) {
    if (g.tile_val(t) != roll || t == robber) continue;
    const int r = g.tile_res(t) - 1;
    uint8_t b[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) b[k] = g.corner(T.tile_corners[t][k]);
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (b[k]) sadd_u8(&tmp.alloc[r][(b[k] >> 2) - 1], b[k] & 3);   // settlement +1, city +2
  }
  CATAN_GROUP_SYNC();
  if (lane == 0) {
    int tot[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) tot[p] = t_hand_total(g, p + 1);
    uint8_t granted = 0;
#pragma unroll
    for (int ri = 0; ri < 5; ++ri) {
      const int r = (0x34021 >> (4 * ri)) & 15;                      // Wood, Ore, Brick, Wheat, Sheep (game.py:153-155)
      const int total = tmp.alloc[r][0] + tmp.alloc[r][1] + tmp.alloc[r][2] + tmp.alloc[r][3];
      if (total > g.bank(r)) continue;                               // all-or-nothing per resource (game.py:171)
      granted |= static_cast<uint8_t>(1u << r);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int a = tmp.alloc[r][p];
        if (a) { g.res(p, r) = static_cast<uint8_t>(g.res(p, r) + a); g.bank(r) = static_cast<uint8_t>(g.bank(r) - a); }
        tot[p] += a;
        tmp.dice_T[p][r] = static_cast<uint8_t>(tot[p]);             // owner's running total when (r, p) is re-clipped (Q4)
      }
    }
    tmp.granted = granted;
  }
  CATAN_GROUP_SYNC();
}

// one item = (observer o, resource r): the requests touch disjoint entries for different items
CATAN_FN void t_est_generic_group(const GameView& g, Seats s, const EstReq& rq, int lane, int nl) {
  const int owner = rq.owner, thief = rq.thief, T_o = rq.T_o, T_t = rq.T_t;
  for (int it = lane; it < 20; it += nl) {
    const int o = it / 5, r = it - 5 * o, observer = o + 1;
    const bool touched = (rq.touched >> r) & 1;
    if (!thief || observer == thief) {                               // game.py:936-954
      if (observer == owner || !touched) continue;
      const int l = label_of(s, observer, owner);
      g.est_max(o, l, r) = static_cast<int16_t>(clipi(g.est_max(o, l, r) + rq.delta[r], 0, T_o));
      g.est_min(o, l, r) = static_cast<int16_t>(clipi(g.est_min(o, l, r) + rq.delta[r], 0, T_o));
    } else if (observer == owner) {                                  // victim knows what was taken (game.py:929-933)
      if (!touched) continue;
      const int l = label_of(s, observer, thief);
      g.est_max(o, l, r) = static_cast<int16_t>(g.est_max(o, l, r) - rq.delta[r]);
      g.est_min(o, l, r) = static_cast<int16_t>(g.est_min(o, l, r) - rq.delta[r]);
    } else {                                                         // third party (game.py:955-971)
      const int lo = label_of(s, observer, owner), lt = label_of(s, observer, thief);
      const int m0 = g.est_max(o, lo, r);                            // the victim entry BEFORE its clip
      g.est_max(o, lo, r) = static_cast<int16_t>(clipi(m0, 0, T_o));
      g.est_min(o, lo, r) = static_cast<int16_t>(clipi(g.est_min(o, lo, r) - 1, 0, T_o));
      if (m0 > 0) {
        g.est_max(o, lt, r) = static_cast<int16_t>(clipi(g.est_max(o, lt, r) + 1, 0, T_t));
        g.est_min(o, lt, r) = static_cast<int16_t>(clipi(g.est_min(o, lt, r), 0, T_t));
      }
    }
  }
}

// one item = (observer o, label l, resource r)
CATAN_FN void t_est_special_group(const GameView& g, Seats s, const StepTmp& tmp, int special, int lane, int nl) {
  if (special == EST_SPECIAL_DICE) {
    for (int it = lane; it < 60; it += nl) {
      const int ol = it / 5, r = it - 5 * ol, o = ol / 3, l = ol - 3 * o;
      if (!((tmp.granted >> r) & 1)) continue;
      const int tp = pid_at_label(s, o + 1, l) - 1;
      const int gain = tmp.alloc[r][tp], T = tmp.dice_T[tp][r];
      g.est_max(o, l, r) = static_cast<int16_t>(clipi(g.est_max(o, l, r) + gain, 0, T));
      g.est_min(o, l, r) = static_cast<int16_t>(clipi(g.est_min(o, l, r) + gain, 0, T));
    }
  } else if (special == EST_SPECIAL_MONOPOLY) {
    const int tot = tmp.mono_lost[0] + tmp.mono_lost[1] + tmp.mono_lost[2] + tmp.mono_lost[3];
    const int mr = tmp.mono_res;
    for (int it = lane; it < 60; it += nl) {
      const int ol = it / 5, r = it - 5 * ol, o = ol / 3, l = ol - 3 * o;
      const int target = pid_at_label(s, o + 1, l);
      if (target == tmp.mono_pid) {                                  // game.py:984-991, unclipped
        if (r != mr) continue;
        g.est_min(o, l, mr) = static_cast<int16_t>(g.est_min(o, l, mr) + tot);
        g.est_max(o, l, mr) = static_cast<int16_t>(g.est_max(o, l, mr) + tot);
      } else {                                                       // game.py:993-1010
        const int T = tmp.mono_T[target - 1];
        const int lost = r == mr ? tmp.mono_lost[target - 1] : 0;
        g.est_max(o, l, r) = static_cast<int16_t>(clipi(g.est_max(o, l, r) - lost, 0, T));
        g.est_min(o, l, r) = static_cast<int16_t>(clipi(g.est_min(o, l, r) - lost, 0, T));
      }
    }
  }
}

CATAN_FN_NOINLINE void t_followups_group(const GameView& g_, const Topo& T, StepTmp& tmp, int lane, int nl) {
  const GameView g = g_;
  CATAN_IN_SMEM(&tmp);
  CATAN_STAGED_RULES_G(g); CATAN_IN_SMEM(&T);
  const int special = tmp.dice_roll ? EST_SPECIAL_DICE : tmp.est_special;   // (nobody writes tmp's flags here: the lanes only read them)
  if (tmp.dice_roll) t_dice_payout_group(g, T, tmp, lane, nl);
  for (int qi = 0; qi < tmp.n_est; ++qi) {
    t_est_generic_group(g, tmp.s, tmp.est[qi], lane, nl);
    CATAN_GROUP_SYNC();
  }
  if (special) t_est_special_group(g, tmp.s, tmp, special, lane, nl);
  CATAN_GROUP_SYNC();
}

// ------------------------------------------------------------------------------------------------
// the scalar part of one env step (wrapper.py:36-50, game.py:527-815), by the game's own thread.  `a`: the env's
// composite action row.  Leaves tmp.err / tmp.lr_pid / tmp.follow for the caller: when tmp.follow is set,
// t_followups_group() must run next, and the longest-road update (tmp.lr_pid) after that.
// ------------------------------------------------------------------------------------------------
CATAN_FN void t_step_scalar(TCx& cx, const int32_t* a, StepTmp& tmp) {
  tmp.n_est = 0; tmp.est_special = EST_SPECIAL_NONE; tmp.dice_roll = 0; tmp.lr_pid = 0; tmp.roll_info = 0; tmp.follow = 0;
  tmp.s = cx.s;
  tmp.acted_pid = static_cast<uint8_t>(t_current_actor(cx.g));
  tmp.act_type = static_cast<uint8_t>(a[CATAN_A_TYPE]);
  Act act;
  int err = t_translate_action(cx, a, act);
  if (!err && cx.cfg->validate_actions) err = t_validate_action(cx, act);
  tmp.err = static_cast<uint8_t>(err);
  if (err) return;
  t_apply_scalar(cx, tmp, act);
  tmp.follow = tmp.dice_roll || tmp.n_est || tmp.est_special;
}

// ------------------------------------------------------------------------------------------------
// Game.randomise_uncertainty (game.py:1207-1282; forward search, RL/forward_search_policy/worker.py:42-58): re-deal everything
// the controlling player `c` cannot see, consistently with what it knows.  (1) the deck and the opponents' hidden cards are
// pooled, shuffled and dealt back (players in dict order Blue, Red, Orange, White; cards popped from the right end);
// (2) every opponent's hand is set to c's MINIMUM belief, and the cards that are then unaccounted for (19 per resource minus
// bank, c's hand and the minima) are dealt one by one, in shuffled order, to the first opponent -- in a freshly shuffled player
// order -- that still has room (hand total below its true total) and whose MAXIMUM belief allows one more card of that
// resource; attempts repeat until every resource adds up to 19 (:1262-1272), which is also the reference's closing assert
// (:1276-1282).  One thread per game; draws come from the game stream in the reference's call order.  Returns the number of
// attempts, or 0 when `max_attempts` were not enough (the reference would loop forever on beliefs that admit no deal).
// ------------------------------------------------------------------------------------------------
CATAN_FN void t_rng_shuffle(TCx& cx, uint8_t* a, int n) {            // Fisher-Yates as pinned in catan_layout.h
  for (int i = n - 1; i > 0; --i) {
    const int j = t_rng_bounded(cx, i + 1);
    const uint8_t x = a[i]; a[i] = a[j]; a[j] = x;
  }
}
CATAN_FN_NOINLINE int t_randomise_uncertainty(TCx& cx, int c, int max_attempts) {
  const GameView& g = cx.g;
  const int dict_order[4] = {BLUE, RED, ORANGE, WHITE};              // game.py:17-22
  const int res_order[5] = {SHEEP, BRICK, ORE, WHEAT, WOOD};         // game.py:1231
  uint8_t pool[25];
  int n = g.deck_n();
  for (int i = 0; i < n; ++i) pool[i] = g.deck(i);
  for (int k = 0; k < 4; ++k) {
    const int q = dict_order[k];
    if (q == c) continue;
    const int nh = g.n_hidden(q - 1);
    for (int i = 0; i < nh; ++i) pool[n++] = static_cast<uint8_t>(t_hidden_at(g, q - 1, i));
  }
  t_rng_shuffle(cx, pool, n);
  for (int k = 0; k < 4; ++k) {
    const int q = dict_order[k];
    if (q == c) continue;
    const int nh = g.n_hidden(q - 1);
    for (int i = 0; i < nh; ++i) t_hidden_set(g, q - 1, i, pool[--n]);
  }
  for (int i = 0; i < 25; ++i) g.deck(i) = i < n ? pool[i] : 0;      // (n is the deck's length again)
  int before[4], unacc[5];
  for (int p = 0; p < 4; ++p) before[p] = t_hand_total(g, p + 1);
  for (int k = 0; k < 5; ++k) {
    const int r = res_order[k];
    int acc = g.bank(r);
    for (int q = 1; q <= 4; ++q) {
      if (q != c) {
        const int d = g.est_min(c - 1, label_of(cx.s, c, q), r);
        acc += d;
        g.res(q - 1, r) = static_cast<uint8_t>(d);
      } else {
        acc += g.res(q - 1, r);
      }
    }
    unacc[k] = 19 - acc;
  }
  uint8_t list[96], prop[4][5];
#ifdef CATAN_DEBUG_RANDOMISE
  if (cx.env_id == CATAN_DEBUG_RANDOMISE) {
    printf("DBG c %d before %d %d %d %d unacc %d %d %d %d %d ctr %u\n", c, before[0], before[1], before[2], before[3], unacc[0], unacc[1], unacc[2], unacc[3], unacc[4], (unsigned)g.rng_ctr());
    for (int q = 1; q <= 4; ++q) if (q != c) printf("DBG q %d label %d min %d %d %d %d %d max %d %d %d %d %d\n", q, label_of(cx.s, c, q),
      (int)g.est_min(c - 1, label_of(cx.s, c, q), 0), (int)g.est_min(c - 1, label_of(cx.s, c, q), 1), (int)g.est_min(c - 1, label_of(cx.s, c, q), 2), (int)g.est_min(c - 1, label_of(cx.s, c, q), 3), (int)g.est_min(c - 1, label_of(cx.s, c, q), 4),
      (int)g.est_max(c - 1, label_of(cx.s, c, q), 0), (int)g.est_max(c - 1, label_of(cx.s, c, q), 1), (int)g.est_max(c - 1, label_of(cx.s, c, q), 2), (int)g.est_max(c - 1, label_of(cx.s, c, q), 3), (int)g.est_max(c - 1, label_of(cx.s, c, q), 4));
  }
#endif
  for (int attempt = 1; attempt <= max_attempts; ++attempt) {
    for (int p = 0; p < 4; ++p) for (int r = 0; r < 5; ++r) prop[p][r] = g.res(p, r);
    int len = 0;
    for (int k = 0; k < 5; ++k) for (int j = 0; j < unacc[k] && len < 96; ++j) list[len++] = static_cast<uint8_t>(res_order[k]);
    t_rng_shuffle(cx, list, len);
    while (len > 0) {
      const int r = list[--len];
      uint8_t keys[4] = {BLUE, RED, ORANGE, WHITE};
      t_rng_shuffle(cx, keys, 4);
      for (int k = 0; k < 4; ++k) {
        const int q = keys[k];
        if (q == c) continue;
        const int tot = prop[q - 1][0] + prop[q - 1][1] + prop[q - 1][2] + prop[q - 1][3] + prop[q - 1][4];
#ifdef CATAN_DEBUG_RANDOMISE
        if (cx.env_id == CATAN_DEBUG_RANDOMISE) printf("DBG att %d len %d r %d k %d q %d tot %d before %d max %d prop %d\n", attempt, len, r, k, q, tot, before[q - 1], (int)g.est_max(c - 1, label_of(cx.s, c, q), r), (int)prop[q - 1][r]);
#endif
        if (tot < before[q - 1] && g.est_max(c - 1, label_of(cx.s, c, q), r) > prop[q - 1][r]) { prop[q - 1][r] += 1; break; }
      }
    }
    bool ok = true;
    for (int r = 0; r < 5; ++r) ok &= prop[0][r] + prop[1][r] + prop[2][r] + prop[3][r] + g.bank(r) == 19;
    if (ok) {
      for (int p = 0; p < 4; ++p) for (int r = 0; r < 5; ++r) g.res(p, r) = prop[p][r];
      return attempt;
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// longest road (game.py:843-919): adjacency masks from a view; the search itself is catan_core.cuh's lp_round
// ------------------------------------------------------------------------------------------------
CATAN_FN void t_lp_build_adj(const GameView& g, const Topo& T, int pid, uint64_t* adj, int lane, int nlanes) {
  for (int v = lane; v < 54; v += nlanes) {
    uint64_t a = 0;
    const uint8_t b = g.corner(v);
    if (!(b && (b >> 2) != pid)) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int e = T.corner_neigh_edge[v][k];
        if (e >= 0 && g.edge(e) == pid) a |= 1ull << T.corner_neigh[v][k];
      }
    }
    adj[v] = a;
  }
}
// scratch of a group-local search (host emulation; the block-cooperative callers are in catan_kernels.cu):
// adj 432 | adjb 432 | ctl | best | path stacks | task ring
#define CATAN_LP_WARP_TASKS 256
#define CATAN_LP_CTL_OFF (2 * CATAN_LP_ADJ_BYTES)
#define CATAN_LP_BEST_OFF (CATAN_LP_CTL_OFF + 4 * CATAN_LP_CTL_WORDS)
#define CATAN_LP_PATH_OFF (CATAN_LP_BEST_OFF + 16)
#define CATAN_LP_TASK_OFF ((CATAN_LP_PATH_OFF + 54 * CATAN_LANES + 15) & ~15)
#define CATAN_LP_SCRATCH_BYTES (CATAN_LP_TASK_OFF + CATAN_LP_WARP_TASKS * 16)
// longest path of one player searched by ONE group of lanes.  scratch: CATAN_LP_SCRATCH_BYTES, 16-byte aligned.
CATAN_FN int t_longest_path(const GameView& g, const Topo& T, int pid, uint8_t* scratch, int lane) {
  uint64_t* adj = reinterpret_cast<uint64_t*>(scratch);
  int32_t* ctl = reinterpret_cast<int32_t*>(scratch + CATAN_LP_CTL_OFF);
  int32_t* best = reinterpret_cast<int32_t*>(scratch + CATAN_LP_BEST_OFF);
  LpTask* ring = reinterpret_cast<LpTask*>(scratch + CATAN_LP_TASK_OFF);
  wsync();
  t_lp_build_adj(g, T, pid, adj, lane, CATAN_LANES);
  if (lane == 0) *best = 0;
  wsync();
  CATAN_LP_RUN(adj, adj, -1, -1, ctl, best, scratch + CATAN_LP_PATH_OFF, CATAN_LANES, lane, ring, CATAN_LP_WARP_TASKS, lane == 0, wsync());
  return *best;
}
// Tables of the THROUGH search for a new road a -> b (lp_round, through mode): adj = arcs out of a corner, adjb = arcs
// into it plus the jump bit b.  Group of nl lanes; the caller synchronises afterwards.
CATAN_FN void t_lp_build_adj2(const GameView& g, const Topo& T, int pid, uint64_t* adj, uint64_t* adjb, int b, int lane, int nl) {
  uint64_t blk = 0;
  for (int v = lane; v < 54; v += nl) {
    uint64_t und = 0;
    const uint8_t bd = g.corner(v);
    if (bd && (bd >> 2) != pid) blk |= 1ull << v;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int e = T.corner_neigh_edge[v][k];
      if (e >= 0 && g.edge(e) == pid) und |= 1ull << T.corner_neigh[v][k];
    }
    adjb[v] = und;
  }
  blk = group_or64(blk);
  for (int v = lane; v < 54; v += nl) {
    const uint64_t und = adjb[v];
    adj[v] = (blk >> v) & 1ull ? 0ull : und;
    adjb[v] = (und & ~blk) | (1ull << b);
  }
}
// longest path of PlayerId pid that contains the road `edge`, same group, same scratch
CATAN_FN int t_through_edge(const GameView& g, const Topo& T, int pid, int edge, uint8_t* scratch, int lane) {
  uint64_t* adj = reinterpret_cast<uint64_t*>(scratch);
  uint64_t* adjb = reinterpret_cast<uint64_t*>(scratch + CATAN_LP_ADJ_BYTES);
  int32_t* ctl = reinterpret_cast<int32_t*>(scratch + CATAN_LP_CTL_OFF);
  int32_t* best = reinterpret_cast<int32_t*>(scratch + CATAN_LP_BEST_OFF);
  LpTask* ring = reinterpret_cast<LpTask*>(scratch + CATAN_LP_TASK_OFF);
  int through = 0;
  for (int dir = 0; dir < 2; ++dir) {
    const int a = T.edge_corners[edge][dir], b = T.edge_corners[edge][dir ^ 1];
    const uint8_t bd = g.corner(a);
    if (bd && (bd >> 2) != pid) continue;                            // a is blocked: no arc a -> b
    wsync();
    t_lp_build_adj2(g, T, pid, adj, adjb, b, lane, CATAN_LANES);
    if (lane == 0) *best = 0;
    wsync();
    CATAN_LP_RUN(adj, adjb, b, a, ctl, best, scratch + CATAN_LP_PATH_OFF, CATAN_LANES, lane, ring, CATAN_LP_WARP_TASKS, lane == 0, wsync());
    if (*best > through) through = *best;
  }
  wsync();
  return through;
}
CATAN_FN bool t_lr_is_shrunk(const GameView& g, int pid, int len) { return g.lr_holder() == pid && g.lr_count() > len; }

// game.py:864-919 given the measured lengths: len of `pid`, and (only when shrunk) other_len[PlayerId] of the rest
CATAN_FN_NOINLINE void t_lr_apply(const GameView& g, int pid, int len, bool shrunk, const uint8_t* other_len) {
  const int holder = g.lr_holder(), count = g.lr_count();
  g.cur_longest_path(pid - 1) = static_cast<uint8_t>(len);
  g.has_path_key(pid - 1) = 1;
  if (!holder) {
    if (len >= 5) { g.lr_holder() = static_cast<uint8_t>(pid); g.lr_count() = static_cast<uint8_t>(len); g.vp(pid - 1) += 2; }
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
          if (player == pid) g.lr_count() = static_cast<uint8_t>(len);
          else { g.lr_holder() = 0; g.lr_count() = 0; g.vp(pid - 1) -= 2; }
        } else {
          g.lr_holder() = static_cast<uint8_t>(player); g.lr_count() = static_cast<uint8_t>(max_len);
          g.vp(player - 1) += 2; g.vp(pid - 1) -= 2;
        }
      } else { g.lr_holder() = 0; g.lr_count() = 0; g.vp(pid - 1) -= 2; }
    } else {
      g.lr_count() = static_cast<uint8_t>(len);
    }
  } else if (len > count) {
    g.vp(holder - 1) -= 2; g.vp(pid - 1) += 2;
    g.lr_holder() = static_cast<uint8_t>(pid); g.lr_count() = static_cast<uint8_t>(len);
  }
}

// ------------------------------------------------------------------------------------------------
// Incremental longest road, ONE THREAD per update.  The reference re-enumerates every simple path of the player's road
// graph each time (game.py:843-862).  The result can be had from the stored length in most cases:
//   * road placed (edge {u,v}): the new maximum is max(old, longest path THROUGH the new edge).  If the edge is a bridge
//     of the player's road graph (96 % of the placements in random play), the part behind u and the part beyond v cannot
//     share a corner, so that path is (longest path into u) + 1 + (longest path out of v) -- two small searches.  Known
//     cheaply when one end had no road before; otherwise the sum is only an upper bound: good enough when it does not
//     beat the stored length.
//   * settlement placed while somebody holds the longest road: the holder's paths are untouched unless the corner lies on
//     one of the holder's roads (and the holder is not the builder).
// The stored length is exact unless an opponent built on the player's network since it was measured (lr_dirty).  All other
// cases -- and searches that exceed CATAN_LR_FAST_ITERS -- are left to the full enumeration (return -1).
// Arc semantics as lp_build_adj: a corner holding an opponent's building has no outgoing arcs (game.py:851-858).
// ------------------------------------------------------------------------------------------------
#ifndef CATAN_LR_FAST_ITERS
#define CATAN_LR_FAST_ITERS 96
#endif
struct RoadBits { uint64_t em_lo, blk; uint32_t em_hi; };   // pid's roads (edge bit set), corners blocked for pid
CATAN_FN RoadBits t_load_road_bits(const GameView& g, int pid) {
  RoadBits R = {0ull, 0ull, 0u};
  CATAN_NO_UNROLL
  for (int e0 = 0; e0 < 72; e0 += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = e0 + j;
      const uint32_t own = g.edge(e) == static_cast<uint32_t>(pid);
      if (e0 < 64) R.em_lo |= static_cast<uint64_t>(own) << e; else R.em_hi |= own << (e - 64);
    }
  }
  CATAN_NO_UNROLL
  for (int c0 = 0; c0 < 54; c0 += 6) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const uint32_t b = g.corner(c0 + j);
      R.blk |= static_cast<uint64_t>(b != 0 && (b >> 2) != static_cast<uint32_t>(pid)) << (c0 + j);
    }
  }
  return R;
}
// new length of PlayerId pid's longest road, or -1 if the full enumeration is needed
// neighbours of corner x over pid's roads (ignoring buildings)
CATAN_FN uint64_t t_road_nb(const Topo& T, const RoadBits& R, int x) {
  uint64_t nb = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int e = T.corner_neigh_edge[x][k];
    if (e < 0) continue;
    const bool own = e < 64 ? (R.em_lo >> e) & 1ull : (R.em_hi >> (e - 64)) & 1u;
    if (own) nb |= 1ull << T.corner_neigh[x][k];
  }
  return nb;
}
// pre (may be null): the player's RoadBits, when the caller has already gathered them (the transition kernel does, one
// lane per corner / edge); und (may be null): likewise the table of t_road_nb() for all 54 corners
CATAN_FN_NOINLINE int t_lr_fast(const GameView& g_, const Topo& T, int pid, int kind, int loc, int placer, const RoadBits* pre = nullptr,
                                const uint64_t* und = nullptr) {
  const GameView g = g_;
  CATAN_STAGED_RULES_G(g); CATAN_IN_SMEM(&T);
  const int holder = g.lr_holder();
  const int old = holder == pid ? g.lr_count() : (g.has_path_key(pid - 1) ? g.cur_longest_path(pid - 1) : 0);
  if (kind == CATAN_LR_SETTLE) {                                     // pid == holder (game.py:552-553): its length is always current
    if (placer == pid) return old;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int e = T.corner_neigh_edge[loc][k];
      if (e >= 0 && g.edge(e) == pid) return -1;                     // the new building may cut the holder's road
    }
    return old;
  }
  if (g.lr_dirty(pid - 1)) return -1;
  if (loc == 0xff) return old;                                       // dummy edge of road building (game.py:585): nothing changed
  const RoadBits R = pre ? *pre : t_load_road_bits(g, pid);
  const int u = T.edge_corners[loc][0], v = T.edge_corners[loc][1];
  // one end had no road before: the new edge is certainly a bridge
  const bool leaf = !((und ? und[u] : t_road_nb(T, R, u)) & ~(1ull << v)) || !((und ? und[v] : t_road_nb(T, R, v)) & ~(1ull << u));
  // four searches in ONE loop (the lanes of a warp are in different searches of different games): into u, out of v, into
  // v, out of u; arcs out of a corner blocked by an opponent's building do not exist
  int through = 0, iters = CATAN_LR_FAST_ITERS, phase = 0, acc = 0;
  int depth = 0, best = 0, node = 0;
  uint64_t visited = 0, above = 0;
  uint8_t st[54];
  bool fresh = true;
  CATAN_NO_UNROLL
  for (;;) {
    if (fresh) {
      if (phase >= 4) break;
      const int a = phase < 2 ? u : v, b = phase < 2 ? v : u;
      if ((R.blk >> a) & 1ull) { phase += 2; continue; }             // no arc a -> b
      node = (phase & 1) ? b : a;
      depth = 0; best = 0; visited = (1ull << a) | (1ull << b); above = ~0ull;
      st[0] = static_cast<uint8_t>(node);
      fresh = false;
    }
    if (--iters < 0) return -1;
    const bool fwd = phase & 1;
    uint64_t nb = und ? und[node] : t_road_nb(T, R, node);
    nb = fwd ? (((R.blk >> node) & 1ull) ? 0ull : nb) : (nb & ~R.blk);
    const uint64_t cand = nb & ~visited & above;
    if (cand) {
      const int t = ctz64(cand);
      st[++depth] = static_cast<uint8_t>(t);
      visited |= 1ull << t;
      node = t; above = ~0ull;
      if (depth > best) best = depth;
    } else if (depth == 0) {
      if (!fwd) acc = best;
      else if (acc + 1 + best > through) through = acc + 1 + best;
      ++phase;
      fresh = true;
    } else {
      visited &= ~(1ull << node);
      above = ~((2ull << node) - 1ull);
      node = st[--depth];
    }
  }
  if (leaf) return through > old ? through : old;
  return through <= old ? old : -1;                                  // the sum is only an upper bound on a cycle
}

// ------------------------------------------------------------------------------------------------
// done / reward / info (wrapper.py:85-112).  Returns true when the game ended and must be reset (cfg.auto_reset); the
// reset itself is done by reset_game_group(), which then patches the two info bytes that describe the new game.
// ------------------------------------------------------------------------------------------------
CATAN_FN bool t_step_finish_inl(TCx& cx_, const StepTmp& tmp, float* reward_out, uint8_t* info_out) {
  TCx cx = cx_;   // (a private copy: the address-space hints below then hold for every access, whatever the stores in between)
  CATAN_IN_LOCAL(&tmp);
  CATAN_STAGED_ENCODE(cx);
  const GameView& g = cx.g;
  struct alignas(16) V16 { uint32_t w[4]; };
  struct alignas(16) F4 { float v[4]; };
  F4 rew = {{0.f, 0.f, 0.f, 0.f}};
  const int err = tmp.err;
  int done = 0, winner = g.winner();
  int vp[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) vp[p] = g.vp(p);
  if (!err) {
    g.episode_steps() += 1;
    // game.py:18-23: dict order Blue, Red, Orange, White; the LAST player with >= 10 VP becomes env.winner
    if (vp[BLUE - 1] >= 10) { done = 1; winner = BLUE; }
    if (vp[RED - 1] >= 10) { done = 1; winner = RED; }
    if (vp[ORANGE - 1] >= 10) { done = 1; winner = ORANGE; }
    if (vp[WHITE - 1] >= 10) { done = 1; winner = WHITE; }
    if (done) g.winner() = static_cast<uint8_t>(winner);
    if (cx.cfg->dense_reward) {                                      // wrapper.py:95-106
      const int ty = tmp.act_type;
      const double bonus = (ty == CATAN_ACT_PLAY_DEV ? 5.0 : 0.0) + (ty == CATAN_ACT_MOVE_ROBBER ? 1.0 : 0.0) -
                           (ty == CATAN_ACT_DISCARD ? 0.3 : 0.0) + (ty == CATAN_ACT_UPGRADE_CITY ? 2.5 : 0.0);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        // python: ((5*dvp + 5) + 1 - 0.3 + 2.5) * factor with at most one bonus non-zero => same double value
        const double r = (5.0 * static_cast<double>(vp[p] - g.curr_vps(p)) + bonus) * static_cast<double>(cx.cfg->reward_annealing_factor);
        rew.v[p] = static_cast<float>(r);
      }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) g.curr_vps(p) = static_cast<int8_t>(vp[p]);
    if (done) {
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (p == winner - 1) rew.v[p] = static_cast<float>(static_cast<double>(rew.v[p]) + static_cast<double>(cx.cfg->win_reward));
    }
  }
  const uint32_t actor = static_cast<uint32_t>(t_current_actor(g));   // game_manager.py:99 reads it before env.reset()
  V16 info;
  info.w[0] = static_cast<uint32_t>(done) | (static_cast<uint32_t>(winner) << 8) | (static_cast<uint32_t>(static_cast<uint8_t>(vp[0])) << 16) |
              (static_cast<uint32_t>(static_cast<uint8_t>(vp[1])) << 24);
  info.w[1] = static_cast<uint32_t>(static_cast<uint8_t>(vp[2])) | (static_cast<uint32_t>(static_cast<uint8_t>(vp[3])) << 8) | (actor << 16) |
              (static_cast<uint32_t>(tmp.acted_pid) << 24);
  info.w[2] = static_cast<uint32_t>(tmp.act_type) | (static_cast<uint32_t>(tmp.roll_info) << 8) | (static_cast<uint32_t>(err) << 16);
  info.w[3] = actor;
  *reinterpret_cast<F4*>(reward_out) = rew;
  *reinterpret_cast<V16*>(info_out) = info;
  return done && cx.cfg->auto_reset;
}
CATAN_FN_NOINLINE bool t_step_finish(TCx& cx, const StepTmp& tmp, float* reward_out, uint8_t* info_out) { CATAN_IN_LOCAL(&tmp); return t_step_finish_inl(cx, tmp, reward_out, info_out); }

// ------------------------------------------------------------------------------------------------
// reset: Board.reset (board.py:67-167) + Game.reset (game.py:39-136) + EnvWrapper.reset (wrapper.py:30-34), by a GROUP of
// `nl` lanes (a warp on the device, one lane in the host build).  The game stream is counter based, so the lanes first
// compute the next CATAN_RESET_WORDS draws in parallel; lane 0 then runs the (inherently serial) Fisher-Yates shuffles
// over small arrays in `arr` and only falls back to computing single draws when the 6/8 rejection loop of board.py:80-81
// needed more than that.  wbuf: CATAN_RESET_WORDS words, arr: 128 bytes, both private to the group (shared memory).
// info_patch (may be null): info row of the step that ended the previous game; gets the new actor and RESET = 1.
// ------------------------------------------------------------------------------------------------
#define CATAN_RESET_WORDS 512
#if defined(CATAN_PROFILE_PHASES) && defined(CATAN_DEVICE)
__device__ unsigned long long d_phase[64];     // [2 * k] = sum of cycles, [2 * k + 1] = count (see catan_kernels.cu)
#define CATAN_RESET_T(k_) do { const long long t_now = clock64(); atomicAdd(&d_phase[k_], static_cast<unsigned long long>(t_now - t_mark)); t_mark = t_now; } while (0)
#define CATAN_RESET_T0() long long t_mark = clock64()
#else
#define CATAN_RESET_T(k_) ((void)0)
#define CATAN_RESET_T0() ((void)0)
#endif
// One Fisher-Yates shuffle of a[0..n) with the draws d0, d0 + 1, ... of the game stream (catan_layout.h): the lanes turn the n - 1
// draws into swap partners in parallel (js), lane 0 then only swaps bytes.  (Round 2 measured ~700 cycles per swap when the one
// lane also fetched the draw and reduced it: a reset took 57 us on average and 350 us at worst, and a step waits for the slowest
// of the ~43 resets of a tick.)
CATAN_FN uint32_t reset_word(const uint32_t* wbuf, uint32_t d_base, uint32_t d, uint64_t seed, uint64_t env_id) {
  const uint32_t i = d - d_base;
  if (i < CATAN_RESET_WORDS) return wbuf[i];
  uint32_t w[4];
  philox4x32(d >> 2, CATAN_STREAM_GAME, static_cast<uint32_t>(env_id), static_cast<uint32_t>(env_id >> 32), static_cast<uint32_t>(seed),
             static_cast<uint32_t>(seed >> 32), w);
  return w[d & 3];
}
#define CATAN_RESET_SHUFFLE(a_, n_)                                                                                        \
  do {                                                                                                                     \
    for (int k_ = lane; k_ < (n_) - 1; k_ += nl)                                                                           \
      js[k_] = static_cast<uint8_t>(mulhi32(reset_word(wbuf, d_base, d + static_cast<uint32_t>(k_), seed, env_id), static_cast<uint32_t>((n_) - k_))); \
    CATAN_GROUP_SYNC();                                                                                                    \
    if (lane == 0)                                                                                                         \
      for (int k_ = 0; k_ < (n_) - 1; ++k_) {                                                                              \
        const int i_ = (n_) - 1 - k_, j_ = js[k_];                                                                         \
        const uint8_t x_ = (a_)[i_], y_ = (a_)[j_];                                                                        \
        (a_)[i_] = y_; (a_)[j_] = x_;                                                                                      \
      }                                                                                                                    \
    d += static_cast<uint32_t>((n_) - 1);                                                                                  \
    CATAN_GROUP_SYNC();                                                                                                    \
  } while (0)

CATAN_FN_NOINLINE void reset_game_group(const GameView& g, const Topo& T, uint64_t seed, uint64_t env_id, uint32_t* wbuf, uint8_t* arr,
                                        int lane, int nl, uint8_t* info_patch) {
  const uint32_t rng = g.rng_ctr(), dec = g.decision_ctr();
  CATAN_RESET_T0();
  CATAN_GROUP_SYNC();
  // clear the record field-size wise (16- and 32-bit fields interleave in units of their own size); the two stream counters survive
  for (int k = lane; k < static_cast<int>(offsetof(GameRec, rng_ctr) / 2); k += nl) g.raw<int16_t>(2 * k) = 0;
  if (lane == 0) { g.episode_steps() = 0; g.actions_this_turn() = 0; g.turn() = 0; }
  for (int k = static_cast<int>(offsetof(GameRec, robber_tile)) + lane; k < static_cast<int>(sizeof(GameRec)); k += nl) g.raw<uint8_t>(k) = 0;
  const uint32_t d_base = rng & ~3u;
  for (int b = lane; b < CATAN_RESET_WORDS / 4; b += nl)
    philox4x32((d_base >> 2) + b, CATAN_STREAM_GAME, static_cast<uint32_t>(env_id), static_cast<uint32_t>(env_id >> 32),
               static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), wbuf + 4 * b);
  uint8_t* terrain = arr;            // 19
  uint8_t* numbers = arr + 19;       // 18 + 1 (entry 18: the desert's "token", 7)
  uint8_t* harb = arr + 38;          // 9
  uint8_t* order = arr + 47;         // 4
  uint8_t* deck = arr + 51;          // 25
  uint8_t* tok = arr + 76;           // 19: index of tile t's number token (the desert: 18)
  uint8_t* js = arr + 96;            // 32: swap partners of the shuffle in progress
  for (int i = lane; i < 19; i += nl) terrain[i] = static_cast<uint8_t>(T.terrain_to_place[i]);
  for (int i = lane; i < 18; i += nl) numbers[i] = static_cast<uint8_t>(T.default_number_order[i]);
  for (int i = lane; i < 9; i += nl) harb[i] = static_cast<uint8_t>(i);
  for (int i = lane; i < 25; i += nl) deck[i] = static_cast<uint8_t>(T.deck_init[i]);
  if (lane == 0) { order[0] = WHITE; order[1] = BLUE; order[2] = ORANGE; order[3] = RED; numbers[18] = 7; }   // (numbers[18]: the desert's "token")
  uint32_t d = rng;                                                  // next draw of the game stream (the same in every lane)
  CATAN_GROUP_SYNC();
  CATAN_RESET_T(56);                                                  // clear + pre-drawn words
  CATAN_RESET_SHUFFLE(terrain, 19);                                  // board.py:71-72
  CATAN_RESET_SHUFFLE(numbers, 18);                                  // board.py:79
  // tile t takes the token at its rank in the placement order, not counting the desert (board.py:88-100)
  for (int t = lane; t < 19; t += nl) {
    int rank = 0, desert_rank = 0;
    for (int i = 0; i < 19; ++i) {
      const int pt = T.number_placement[i];
      if (pt == t) rank = i;
      if (terrain[pt] == 0) desert_rank = i;
    }
    tok[t] = static_cast<uint8_t>(rank == desert_rank ? 18 : rank - (desert_rank < rank ? 1 : 0));
  }
  CATAN_GROUP_SYNC();
  CATAN_RESET_T(57);                                                  // terrain + first number shuffle
  // board.py:80-81: reshuffle until no 6 / 8 touch (board.py:50-65); the check is one lane per tile
  for (;;) {
    uint32_t m68 = 0;
    for (int t = lane; t < 19; t += nl) { const int v = numbers[tok[t]]; if (v == 6 || v == 8) m68 |= 1u << t; }
    m68 = group_or32(m68);
    uint32_t bad = 0;
    for (int t = lane; t < 19; t += nl) {
      if (!((m68 >> t) & 1u)) continue;
      for (int k = 0; k < 6; ++k) { const int nb = T.tile_neigh[t][k]; if (nb >= 0 && ((m68 >> nb) & 1u)) bad = 1; }
    }
    if (!group_or32(bad)) break;
    CATAN_RESET_SHUFFLE(numbers, 18);
#if defined(CATAN_PROFILE_PHASES) && defined(CATAN_DEVICE)
    if (lane == 0) atomicAdd(&d_phase[60], 1ull);
#endif
  }
  CATAN_RESET_T(58);                                                  // the 6 / 8 rejection loop
  CATAN_RESET_SHUFFLE(harb, 9);                                      // board.py:83-84
  CATAN_RESET_SHUFFLE(order, 4);                                     // game.py:41-42
  CATAN_RESET_SHUFFLE(deck, 25);                                     // game.py:75-78
  for (int t = lane; t < 19; t += nl) {                              // board.py:88-100
    g.tile_res(t) = terrain[t];
    g.tile_val(t) = numbers[tok[t]];
    if (terrain[t] == 0) g.robber_tile() = static_cast<uint8_t>(t);
  }
  for (int i = lane; i < 25; i += nl) g.deck(i) = deck[i];
  for (int i = lane; i < 9; i += nl) g.harbour_perm(i) = harb[i];
  if (lane == 0) {
    g.rng_ctr() = d;
    g.decision_ctr() = dec;
    for (int i = 0; i < 4; ++i) g.player_order(i) = order[i];
    g.players_go() = order[0];
    for (int r = 0; r < 5; ++r) g.bank(r) = 19;                      // game.py:48-54
    for (int p = 0; p < 4; ++p) { g.settlements_left(p) = 5; g.cities_left(p) = 4; g.second_corner(p) = -1; }
    g.deck_n() = 25;
    g.initial_phase() = 1;
    if (info_patch) { info_patch[CATAN_INFO_ACTOR] = order[0]; info_patch[CATAN_INFO_RESET] = 1; }
    CATAN_RESET_T(59);                                                // the other shuffles + write-out
  }
  CATAN_GROUP_SYNC();
}

// info row written by catan_reset / catan_import_state (no step happened)
CATAN_FN void t_write_info_fresh(const GameView& g, uint8_t* info, bool was_reset) {
  struct alignas(16) V16 { uint32_t w[4]; };
  V16 v;
  v.w[0] = (static_cast<uint32_t>(g.winner()) << 8) | (static_cast<uint32_t>(static_cast<uint8_t>(g.vp(0))) << 16) |
           (static_cast<uint32_t>(static_cast<uint8_t>(g.vp(1))) << 24);
  v.w[1] = static_cast<uint32_t>(static_cast<uint8_t>(g.vp(2))) | (static_cast<uint32_t>(static_cast<uint8_t>(g.vp(3))) << 8) |
           (static_cast<uint32_t>(t_current_actor(g)) << 16);
  v.w[2] = was_reset ? (1u << 24) : 0u;
  v.w[3] = 0;
  *reinterpret_cast<V16*>(info) = v;
}

// ------------------------------------------------------------------------------------------------
// legal-action masks as BIT sets (wrapper.py:168-412): every head defaults to all ones except the
// type head (wrapper.py:172-185)
// ------------------------------------------------------------------------------------------------
struct MaskBits {
  uint64_t settle, city;     // rows 0 and 1 of the corner head (row 2 is never restricted)
  uint64_t edge_lo;          // edges 0..63
  uint32_t edge_hi;          // edges 64..71 and the dummy edge (bit 8)
  uint32_t type, tile, dev, accept, player, res_a, res_b, discard;
};
#define CATAN_ALL54 ((1ull << 54) - 1ull)

// Placement scan of ONE game by a GROUP of `nl` lanes (a warp on the device), one lane per corner / edge / tile: the
// occupancy bit boards seen by PlayerId pid, corner.py:24-39 for all corners and edge.py:23-42 for all edges.  Every lane
// returns the complete result.  solo (with lane 0, nl 1): the calling thread scans its own game alone -- cheaper for a
// warp when most of its 32 games need a scan.
struct Scan {
  uint64_t settle_free;      // distance rule only (corner.py:28-33)
  uint64_t road_at;          // an own road touches the corner
  uint64_t mine_settle;      // own settlements
  uint64_t road_lo;          // edges 0..63 where pid may build (edge.py:33-42)
  uint64_t e_any_lo;         // occupied edges 0..63
  uint32_t road_hi, e_any_hi;   // edges 64..71
  uint32_t tile_bld;         // tiles with any building on a corner (wrapper.py:308-320, Q1)
};
CATAN_FN_NOINLINE Scan t_scan_group(const GameView& g_, const Topo& T, const TopoX& X, int pid, int lane, int nl, bool solo = false) {
  const GameView g = g_;
  CATAN_STAGED_ENCODE_G(g); CATAN_IN_SMEM(&T); CATAN_IN_SMEM(&X);
  uint64_t bld = 0, mine = 0, ms = 0;
  for (int c = lane; c < 54; c += nl) {
    const uint32_t b = g.corner(c);
    bld |= static_cast<uint64_t>(b != 0) << c;
    mine |= static_cast<uint64_t>(b != 0 && (b >> 2) == static_cast<uint32_t>(pid)) << c;
    ms |= static_cast<uint64_t>(b == static_cast<uint32_t>((pid << 2) | 1)) << c;
  }
  uint64_t ea = 0, em = 0;
  uint32_t eah = 0, emh = 0;
  for (int e = lane; e < 72; e += nl) {
    const uint32_t b = g.edge(e);
    if (e < 64) { ea |= static_cast<uint64_t>(b != 0) << e; em |= static_cast<uint64_t>(b == static_cast<uint32_t>(pid)) << e; }
    else { eah |= static_cast<uint32_t>(b != 0) << (e - 64); emh |= static_cast<uint32_t>(b == static_cast<uint32_t>(pid)) << (e - 64); }
  }
  if (!solo) {
    bld = group_or64(bld); mine = group_or64(mine); ms = group_or64(ms);
    ea = group_or64(ea); em = group_or64(em); eah = group_or32(eah); emh = group_or32(emh);
  }
  uint64_t fre = 0, ra = 0;
  for (int c = lane; c < 54; c += nl) {
    const bool blocked = ((X.corner_nb[c] | (1ull << c)) & bld) != 0;
    const bool own_road = ((X.corner_edges_lo[c] & em) | static_cast<uint64_t>(X.corner_edges_hi[c] & emh)) != 0;
    fre |= static_cast<uint64_t>(!blocked) << c;
    ra |= static_cast<uint64_t>(own_road) << c;
  }
  if (!solo) { fre = group_or64(fre); ra = group_or64(ra); }
  const uint64_t reach = mine | (~bld & ra);                         // a road of pid may start here
  uint64_t lo = 0;
  uint32_t hi = 0, tl = 0;
  for (int e = lane; e < 72; e += nl) {
    const int c1 = T.edge_corners[e][0], c2 = T.edge_corners[e][1];
    const uint64_t ok = ((reach >> c1) | (reach >> c2)) & 1ull;
    if (e < 64) lo |= ok << e; else hi |= static_cast<uint32_t>(ok) << (e - 64);
  }
  for (int t = lane; t < 19; t += nl) tl |= static_cast<uint32_t>((X.tile_cmask[t] & bld) != 0) << t;
  Scan sc;
  sc.settle_free = fre; sc.road_at = ra; sc.mine_settle = ms;
  if (!solo) { lo = group_or64(lo); hi = group_or32(hi); tl = group_or32(tl); }
  sc.road_lo = lo & ~ea; sc.road_hi = hi & ~eah & 0xffu;
  sc.e_any_lo = ea; sc.e_any_hi = eah;
  sc.tile_bld = tl;
  return sc;
}

CATAN_FN void t_mask_play_dev(const GameView& g, int p, MaskBits& m) {   // wrapper.py:221-228 / :262-269 / :368-388
  if (g.n_hidden(p) == 0 || g.played_dev()) return;
  const uint32_t counts = t_hidden_counts(g, p);
  uint32_t bank_bits = 0;
  int bank_total = 0;
#pragma unroll
  for (int r = 0; r < 5; ++r) { const int b = g.bank(r); bank_total += b; bank_bits |= static_cast<uint32_t>(b > 0) << r; }
  uint32_t valid = 0;
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    const int k = (counts >> (6 * c)) & 63;
    if (k > 0 && g.bought(c) < k && (c != CATAN_DEV_YOP || bank_total > 0)) valid |= 1u << c;
  }
  if (!valid) return;
  m.type |= 1u << CATAN_ACT_PLAY_DEV;
  m.dev = valid;
  if (valid & (1u << CATAN_DEV_YOP)) {                               // Q11: bank mask lands on row 2 of head 9 and on head 10
    m.res_a = (m.res_a & ~(31u << 10)) | (bank_bits << 10);
    m.res_b = bank_bits;
  }
}

// The masks of one game in two halves around the placement scan: t_masks_pre() handles the phases that need no board
// scan and returns true when t_scan_group() must run for this game (PlayerId = players_go) before t_masks_post().
struct MaskPlan { uint8_t post, initial, rb, capped, init_settle, want_settle, want_city, want_road, want_tiles; };   // post: t_masks_post() must run
CATAN_FN bool t_masks_pre_inl(const TCx& cx_, MaskBits& m, MaskPlan& pl) {
  TCx cx = cx_;   // (a private copy: the address-space hints below then hold for every access, whatever the stores in between)
  CATAN_IN_LOCAL(&m); CATAN_IN_LOCAL(&pl);
  CATAN_STAGED_ENCODE(cx);
  const GameView& g = cx.g;
  const Topo& T = *cx.T;
  m.type = 0; m.settle = CATAN_ALL54; m.city = CATAN_ALL54; m.edge_lo = ~0ull; m.edge_hi = 0x1ffu; m.tile = (1u << 19) - 1u;
  m.dev = 31u; m.accept = 3u; m.player = 0x1ffu; m.res_a = 0xfffffu; m.res_b = 31u; m.discard = 31u;
  pl.post = 0; pl.initial = 0; pl.rb = 0; pl.capped = 0; pl.init_settle = 0; pl.want_settle = 0; pl.want_city = 0; pl.want_road = 0; pl.want_tiles = 0;
  const int pid = g.players_go(), p = pid - 1;
  if (g.need_discard()) {                                            // wrapper.py:186-192
    const int d = g.discard_queue(0);
    m.type = 1u << CATAN_ACT_DISCARD;
    uint32_t bits = 0;
#pragma unroll
    for (int r = 0; r < 5; ++r) bits |= static_cast<uint32_t>(g.res(d - 1, r) != 0) << r;
    m.discard = bits;
    return false;
  }
  const bool initial = g.initial_phase(), rb = g.rb_active();
  if (!initial && !rb) {
    if (g.just_moved_robber()) {                                     // wrapper.py:210-213, :341-351
      m.type = 1u << CATAN_ACT_STEAL;
      uint32_t row = 0;
      const int rt = g.robber_tile();
      for (int k = 0; k < 6; ++k) {
        const uint8_t b = g.corner(T.tile_corners[rt][k]);
        if (b && (b >> 2) != pid) row |= 1u << label_of(cx.s, pid, b >> 2);
      }
      m.player = (m.player & ~(7u << 3)) | (row << 3);
      return false;
    }
    if (g.must_respond()) {                                          // wrapper.py:214-218, :353-365
      m.type = 1u << CATAN_ACT_RESPOND;
      uint32_t cnt = 0;
      const int nr = g.n_recv(), tt = g.trade_target() - 1;
      for (int k = 0; k < nr; ++k) cnt += 1u << (4 * (g.recv(k) - 1));
      bool ok = true;
#pragma unroll
      for (int r = 0; r < 5; ++r) ok &= g.res(tt, r) >= static_cast<int>((cnt >> (4 * r)) & 15);
      m.accept = 2u | static_cast<uint32_t>(ok);
      return false;
    }
    if (!g.dice_rolled()) {                                          // wrapper.py:219-229
      m.type = 1u << CATAN_ACT_ROLL_DICE;
      t_mask_play_dev(g, p, m);
      return false;
    }
  }
  // the placement phases: initial (wrapper.py:195-204), road building (:206-209) and the main phase (:232-290)
  int h[5];
#pragma unroll
  for (int r = 0; r < 5; ++r) h[r] = g.res(p, r);
  const bool capped = !initial && !rb && cx.cfg->max_actions_per_turn >= 0 && g.actions_this_turn() > cx.cfg->max_actions_per_turn;
  const bool init_settle = initial && (g.init_settlements(p) == 0 || (g.init_settlements(p) == 1 && g.init_roads(p) == 1));
  const bool want_settle = init_settle || (!initial && !rb && !capped && h[WHEAT] && h[SHEEP] && h[WOOD] && h[BRICK]);
  const bool want_city = !initial && !rb && !capped && h[WHEAT] >= 2 && h[ORE] >= 3 && g.cities_left(p) > 0;
  const bool want_road = (initial && !init_settle) || rb || (!initial && !capped && h[WOOD] && h[BRICK]);
  const bool want_tiles = !initial && !rb && !capped && g.can_move_robber();
  if (!initial && !rb) m.type = 1u << CATAN_ACT_END_TURN;            // wrapper.py:232
  pl.post = 1; pl.initial = initial; pl.rb = rb; pl.capped = capped; pl.init_settle = init_settle;
  pl.want_settle = want_settle; pl.want_city = want_city; pl.want_road = want_road; pl.want_tiles = want_tiles;
  return want_settle || want_city || want_road || want_tiles;
}
CATAN_FN_NOINLINE bool t_masks_pre(const TCx& cx, MaskBits& m, MaskPlan& pl) { CATAN_IN_LOCAL(&m); CATAN_IN_LOCAL(&pl); return t_masks_pre_inl(cx, m, pl); }

CATAN_FN void t_masks_post_inl(const TCx& cx_, MaskBits& m, const MaskPlan& pl, const Scan& sc) {
  TCx cx = cx_;   // (a private copy: the address-space hints below then hold for every access, whatever the stores in between)
  CATAN_IN_LOCAL(&m); CATAN_IN_LOCAL(&pl); CATAN_IN_SMEM(&sc);
  CATAN_STAGED_ENCODE(cx);
  const GameView& g = cx.g;
  const int pid = g.players_go(), p = pid - 1;
  const bool initial = pl.initial, rb = pl.rb;
  if (pl.want_settle) {
    if (pl.init_settle) {
      m.type = 1u << CATAN_ACT_PLACE_SETTLEMENT;
      m.settle = sc.settle_free;
    } else {                                                         // wrapper.py:238-243
      const uint64_t ok = sc.settle_free & sc.road_at;
      if (ok && g.settlements_left(p) > 0) { m.type |= 1u << CATAN_ACT_PLACE_SETTLEMENT; m.settle = ok; }
    }
  }
  if (pl.want_city && sc.mine_settle) { m.type |= 1u << CATAN_ACT_UPGRADE_CITY; m.city = sc.mine_settle; }   // wrapper.py:245-250
  if (pl.want_road) {                                                // wrapper.py:322-339
    uint64_t lo = sc.road_lo;
    uint32_t hi = sc.road_hi;
    if (initial && g.init_settlements(p) == 2) {                     // the second road must touch the second settlement (edge.py:27-31)
      const int c2 = g.second_corner(p);
      lo = cx.X->corner_edges_lo[c2] & ~sc.e_any_lo;
      hi = cx.X->corner_edges_hi[c2] & ~sc.e_any_hi;
    }
    const bool placed = (lo | hi) != 0;
    if (initial) { m.type = 1u << CATAN_ACT_PLACE_ROAD; m.edge_lo = lo; m.edge_hi = hi; }
    else if (rb) { m.type = 1u << CATAN_ACT_PLACE_ROAD; m.edge_lo = lo; m.edge_hi = hi | (placed ? 0u : 0x100u); }
    else if (placed) { m.type |= 1u << CATAN_ACT_PLACE_ROAD; m.edge_lo = lo; m.edge_hi = hi; }
  }
  if (initial || rb) return;
  if (pl.want_tiles) m.tile = sc.tile_bld;                           // wrapper.py:278-281, :308-320 (Q1: any building)
  if (pl.capped) return;
  int h[5];
#pragma unroll
  for (int r = 0; r < 5; ++r) h[r] = g.res(p, r);
  if (h[WHEAT] && h[SHEEP] && h[ORE] && g.deck_n() > 0) m.type |= 1u << CATAN_ACT_BUY_DEV;
  t_mask_play_dev(g, p, m);                                          // wrapper.py:262-269
  {
    uint32_t give = 0, get = 0;                                      // wrapper.py:271-276, :390-412 (Q13)
    const int hb = g.harbours(p);
#pragma unroll
    for (int r = 0; r < 5; ++r) {
      const int rate = (hb >> (r + 1)) & 1 ? 2 : ((hb & 1) ? 3 : 4);
      give |= static_cast<uint32_t>(h[r] >= rate) << r;
      get |= static_cast<uint32_t>(g.bank(r) > 0) << r;
    }
    if (give && get) {
      m.type |= 1u << CATAN_ACT_EXCHANGE;
      m.res_a = (m.res_a & ~31u) | give;
      m.res_b = get;
    }
    if (h[0] + h[1] + h[2] + h[3] + h[4] > 0 &&                      // wrapper.py:283-289
        (cx.cfg->max_proposed_trades_per_turn < 0 || g.trades_this_turn() < cx.cfg->max_proposed_trades_per_turn))
      m.type |= 1u << CATAN_ACT_PROPOSE_TRADE;
    if (g.can_move_robber()) m.type |= 1u << CATAN_ACT_MOVE_ROBBER;
  }
}
CATAN_FN_NOINLINE void t_masks_post(const TCx& cx, MaskBits& m, const MaskPlan& pl, const Scan& sc) { CATAN_IN_LOCAL(&m); CATAN_IN_LOCAL(&pl); t_masks_post_inl(cx, m, pl, sc); }

// the 325 mask entries as one bit string (bit i = entry i of the packed row, catan_layout.h)
struct MaskFlat { uint32_t w[11]; };
template <int POS, int N>
CATAN_FN void flat_put(MaskFlat& F, uint64_t v) {
  constexpr int w0 = POS / 32, s = POS % 32;
  F.w[w0] |= static_cast<uint32_t>(v << s);
  if constexpr (s + N > 32) F.w[w0 + 1] |= static_cast<uint32_t>(v >> (32 - s));
  if constexpr (s + N > 64) F.w[w0 + 2] |= static_cast<uint32_t>(v >> (64 - s));
}
CATAN_FN void t_flatten_masks(const MaskBits& m, MaskFlat& F) {
#pragma unroll
  for (int i = 0; i < 11; ++i) F.w[i] = 0;
  flat_put<CATAN_MASK_TYPE, 13>(F, m.type);
  flat_put<CATAN_MASK_CORNER, 54>(F, m.settle);
  flat_put<CATAN_MASK_CORNER + 54, 54>(F, m.city);
  flat_put<CATAN_MASK_CORNER + 108, 54>(F, CATAN_ALL54);
  flat_put<CATAN_MASK_EDGE, 64>(F, m.edge_lo);
  flat_put<CATAN_MASK_EDGE + 64, 9>(F, m.edge_hi);
  flat_put<CATAN_MASK_TILE, 19>(F, m.tile);
  flat_put<CATAN_MASK_DEV, 5>(F, m.dev);
  flat_put<CATAN_MASK_ACCEPT, 2>(F, m.accept);
  flat_put<CATAN_MASK_PLAYER, 9>(F, m.player);
  flat_put<CATAN_MASK_GIVE, 12>(F, 0xfffu);
  flat_put<CATAN_MASK_RES_A, 20>(F, m.res_a);
  flat_put<CATAN_MASK_RES_B, 5>(F, m.res_b);
  flat_put<CATAN_MASK_DISCARD, 5>(F, m.discard);
}
// bits -> one byte per entry: 4 bits -> the 4 bytes of a word
CATAN_FN uint32_t spread4(uint32_t x) { return (x * 0x00204081u) & 0x01010101u; }
CATAN_FN void t_store_mask_row(const MaskFlat& F, uint8_t* row) {   // row: CATAN_MASK_STRIDE bytes, 16-byte aligned
  struct alignas(16) V16 { uint32_t a, b, c, d; };
  V16* out = reinterpret_cast<V16*>(row);
#pragma unroll
  for (int j = 0; j < CATAN_MASK_STRIDE / 16; ++j) {
    const uint32_t b = (F.w[j >> 1] >> (16 * (j & 1))) & 0xffffu;
    V16 v = {spread4(b & 15u), spread4((b >> 4) & 15u), spread4((b >> 8) & 15u), spread4(b >> 12)};
    out[j] = v;
  }
}
// packed row -> bit sets (stand-alone sampler reading a mask row back)
CATAN_FN void t_load_mask_row(const uint8_t* row, MaskBits& m) {
  auto bits = [&](int pos, int n) {
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) v |= static_cast<uint64_t>(row[pos + i] != 0) << i;
    return v;
  };
  m.type = static_cast<uint32_t>(bits(CATAN_MASK_TYPE, 13));
  m.settle = bits(CATAN_MASK_CORNER, 54); m.city = bits(CATAN_MASK_CORNER + 54, 54);
  m.edge_lo = bits(CATAN_MASK_EDGE, 64); m.edge_hi = static_cast<uint32_t>(bits(CATAN_MASK_EDGE + 64, 9));
  m.tile = static_cast<uint32_t>(bits(CATAN_MASK_TILE, 19)); m.dev = static_cast<uint32_t>(bits(CATAN_MASK_DEV, 5));
  m.accept = static_cast<uint32_t>(bits(CATAN_MASK_ACCEPT, 2)); m.player = static_cast<uint32_t>(bits(CATAN_MASK_PLAYER, 9));
  m.res_a = static_cast<uint32_t>(bits(CATAN_MASK_RES_A, 20)); m.res_b = static_cast<uint32_t>(bits(CATAN_MASK_RES_B, 5));
  m.discard = static_cast<uint32_t>(bits(CATAN_MASK_DISCARD, 5));
}

// ------------------------------------------------------------------------------------------------
// pinned random-legal sampler (twin of oracle/ref_harness.py:sample_action) on bit sets:
// pick = index of the floor(w*k/2^32)-th set entry (k = number of set entries), 0 if none
// ------------------------------------------------------------------------------------------------
CATAN_FN int popc32(uint32_t m) {
#ifdef CATAN_DEVICE
  return __popc(m);
#else
  return __builtin_popcount(m);
#endif
}
CATAN_FN int nth_set32(uint32_t m, int j) {   // index of the j-th (0-based) set bit of m; popcount descent (__fns is a slow loop)
  int pos = 0;
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const int c = popc32((m >> pos) & ((1u << w) - 1u));
    if (j >= c) { j -= c; pos += w; }
  }
  return pos;
}
CATAN_FN int pick96(uint32_t m0, uint32_t m1, uint32_t m2, uint32_t w) {
  const int k0 = popc32(m0), k1 = popc32(m1), k = k0 + k1 + popc32(m2);
  if (!k) return 0;
  const int j = static_cast<int>(mulhi32(w, static_cast<uint32_t>(k)));
  if (j < k0) return nth_set32(m0, j);
  if (j < k0 + k1) return 32 + nth_set32(m1, j - k0);
  return 64 + nth_set32(m2, j - k0 - k1);
}
CATAN_FN int pick32(uint32_t m, uint32_t w) { return pick96(m, 0u, 0u, w); }
CATAN_FN int pick64(uint64_t m, uint32_t w) { return pick96(static_cast<uint32_t>(m), static_cast<uint32_t>(m >> 32), 0u, w); }

// hand_bits: bit r set iff the acting player holds resource r (== obs current_resources[1..5] != 0, wrapper.py:70-71)
CATAN_FN void t_sample_action_inl(const MaskBits& m, uint32_t hand_bits, uint64_t seed, uint64_t env_id, uint32_t decision, int32_t* out) {
  uint32_t w[4];
  philox4x32(decision, CATAN_STREAM_SAMPLER, static_cast<uint32_t>(env_id), static_cast<uint32_t>(env_id >> 32),
             static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), w);
  const int t = pick32(m.type, w[0]);
  int corner = 0, edge = 0, tile = 0, card = 0, accept = 0, player = 0, give = 0, recv = 0, res_a = 0, res_b = 0, discard = 0;
  switch (t) {
    case CATAN_ACT_PLACE_SETTLEMENT: corner = pick64(m.settle, w[1]); break;
    case CATAN_ACT_UPGRADE_CITY: corner = pick64(m.city, w[1]); break;
    case CATAN_ACT_PLACE_ROAD: edge = pick96(static_cast<uint32_t>(m.edge_lo), static_cast<uint32_t>(m.edge_lo >> 32), m.edge_hi, w[1]); break;
    case CATAN_ACT_MOVE_ROBBER: tile = pick32(m.tile, w[1]); break;
    case CATAN_ACT_PLAY_DEV:
      card = pick32(m.dev, w[1]);
      if (card == CATAN_DEV_MONOPOLY) res_a = pick32((m.res_a >> 10) & 31u, w[2]);
      else if (card == CATAN_DEV_YOP) {
        res_a = pick32((m.res_a >> 15) & 31u, w[2]);
        res_b = pick32(m.res_b, w[3]);
      }
      break;
    case CATAN_ACT_EXCHANGE:
      res_a = pick32(m.res_a & 31u, w[1]);
      res_b = pick32(m.res_b, w[2]);
      break;
    case CATAN_ACT_PROPOSE_TRADE:
      player = pick32(m.player & 7u, w[1]);
      give = 1 + pick32(hand_bits, w[2]);                            // a resource the proposer holds
      recv = 1 + static_cast<int>(mulhi32(w[3], 5u));
      break;
    case CATAN_ACT_RESPOND: accept = pick32(m.accept, w[1]); break;
    case CATAN_ACT_STEAL: player = pick32((m.player >> 3) & 7u, w[1]); break;
    case CATAN_ACT_DISCARD: discard = pick32(m.discard, w[1]); break;
    default: break;
  }
  struct alignas(16) I4 { int32_t a, b, c, d; };
  I4* o = reinterpret_cast<I4*>(out);                                // 80-byte row, 16-byte aligned
  const I4 r0 = {t, corner, edge, tile}, r1 = {card, accept, player, give}, r2 = {0, 0, 0, recv}, r3 = {0, 0, 0, res_a}, r4 = {res_b, discard, 0, 0};
  o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3; o[4] = r4;
}
CATAN_FN_NOINLINE void t_sample_action(const MaskBits& m, uint32_t hand_bits, uint64_t seed, uint64_t env_id, uint32_t decision, int32_t* out) { CATAN_IN_LOCAL(&m); t_sample_action_inl(m, hand_bits, seed, env_id, decision, out); }
static_assert(CATAN_A_TYPE == 0 && CATAN_A_CORNER == 1 && CATAN_A_EDGE == 2 && CATAN_A_TILE == 3 && CATAN_A_CARD == 4 && CATAN_A_ACCEPT == 5 &&
              CATAN_A_PLAYER == 6 && CATAN_A_GIVE == 7 && CATAN_A_RECV == 11 && CATAN_A_RES_A == 15 && CATAN_A_RES_B == 16 &&
              CATAN_A_DISCARD == 17 && CATAN_ACTION_WORDS == 20, "t_sample_action packs the action row by hand");

// ------------------------------------------------------------------------------------------------
// observation (wrapper.py:52-83, :491-524, :526-709).
//
// A row is cut into CATAN_OBS_PARTS 16-byte aligned parts and every part is produced by ONE thread in registers, without
// any scratch memory: almost every byte of a row is 0 or 1, so a part is first built as a BIT IMAGE (bit i = byte i of the
// part) by a handful of shifts at compile-time positions, the few count-valued bytes (production table, road / army lengths,
// card lists, meta) are placed into a word image at compile-time offsets, and each 16-byte piece then leaves as
// spread(16 bits) | count words in one streaming vector store.  (Round 1 scattered bytes through a 128-byte shared-memory
// window per thread: 46 % of the encode kernel's instructions and 20 KB of shared memory per block.)
// ------------------------------------------------------------------------------------------------
CATAN_FN int t_bucket8(int n) { return n < 5 ? n : (n < 8 ? 5 : (n < 11 ? 6 : 7)); }                        // wrapper.py:554-561
CATAN_FN int t_bucket7(int n) { return n <= 2 ? n : (n <= 5 ? 3 : (n <= 7 ? 4 : (n <= 10 ? 5 : 6))); }       // wrapper.py:662-671
// 4 nibbles -> the 4 bytes of a word
CATAN_FN uint32_t nib4(uint32_t x) { x = (x | (x << 8)) & 0x00FF00FFu; return (x | (x << 4)) & 0x0F0F0F0Fu; }

template <int I, int N, class F>
CATAN_FN void static_for(F&& f) {
  if constexpr (I < N) { f(std::integral_constant<int, I>{}); static_for<I + 1, N>(f); }
}

// a part of up to 160 bytes: bit image (w0: bytes 0-63, w1: 64-127, w2: 128-159) + count bytes as words
struct PartImg {
  uint64_t w0, w1, w2;
  uint32_t pw[40];
};
// OR a word of 4 count bytes whose first byte lies at byte offset O of the part (compile time; bytes outside [0, 4 * NW) drop)
template <int O, int NW>
CATAN_FN void pw_put(uint32_t (&pw)[40], uint32_t v) {
  constexpr int idx = O >= 0 ? O / 4 : -((3 - O) / 4), s = O - idx * 4;
  if constexpr (s == 0) {
    if constexpr (idx >= 0 && idx < NW) pw[idx] |= v;
  } else {
    if constexpr (idx >= 0 && idx < NW) pw[idx] |= v << (8 * s);
    if constexpr (idx + 1 >= 0 && idx + 1 < NW) pw[idx + 1] |= v >> (32 - 8 * s);
  }
}
// OR a 160-bit block image (b0, b1, b2), shifted by SH bits (compile time, -8 < SH < 64), into the part image
template <int SH>
CATAN_FN void img_or(PartImg& P, uint64_t b0, uint64_t b1, uint64_t b2) {
  if constexpr (SH == 0) { P.w0 |= b0; P.w1 |= b1; P.w2 |= b2; }
  else if constexpr (SH > 0) { P.w0 |= b0 << SH; P.w1 |= (b1 << SH) | (b0 >> (64 - SH)); P.w2 |= (b2 << SH) | (b1 >> (64 - SH)); }
  else { P.w0 |= (b0 >> -SH) | (b1 << (64 + SH)); P.w1 |= (b1 >> -SH) | (b2 << (64 + SH)); P.w2 |= b2 >> -SH; }
}
// two adjacent 16-byte pieces of a row (dst 32-byte aligned) in ONE store: every lane of a row warp writes to its own row, so a
// store instruction costs one L1 wavefront per lane whatever its width -- the 256-bit form halves the wavefronts per byte
struct alignas(16) ObsPiece { uint32_t a, b, c, d; };
// WIDE: the 256-bit form (rows launch of a step only).  In a CALLED function nvcc 12.9 lowered the same statement to a 32-bit store
// of the first word (caught by the golden replay), so every other path writes the pair as two 128-bit stores.
template <bool WIDE>
CATAN_FN void obs_store32(uint8_t* dst, const ObsPiece& x, const ObsPiece& y) {
#ifdef CATAN_DEVICE
  if constexpr (WIDE) {
    asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(dst), "r"(x.a), "r"(x.b), "r"(x.c), "r"(x.d),
                 "r"(y.a), "r"(y.b), "r"(y.c), "r"(y.d) : "memory");                      // streamed: rows are not re-read here
  } else {
    __stcs(reinterpret_cast<uint4*>(dst), make_uint4(x.a, x.b, x.c, x.d));
    __stcs(reinterpret_cast<uint4*>(dst + 16), make_uint4(y.a, y.b, y.c, y.d));
  }
#else
  *reinterpret_cast<ObsPiece*>(dst) = x;
  *reinterpret_cast<ObsPiece*>(dst + 16) = y;
#endif
}
// the pieces [0, NP) of a part -> row + lo (NP even, the part 32-byte aligned)
template <int NP, bool WIDE>
CATAN_FN void img_store(const PartImg& P, uint8_t* dst) {
  static_assert(NP % 2 == 0, "pieces leave in pairs");
  static_for<0, NP / 2>([&](auto I) {
    constexpr int i0 = 2 * decltype(I)::value;
    ObsPiece v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = i0 + h;
      const uint64_t w = i < 4 ? P.w0 : (i < 8 ? P.w1 : P.w2);
      const uint32_t b16 = static_cast<uint32_t>(w >> (16 * (i & 3))) & 0xffffu;
      v[h] = ObsPiece{spread4(b16 & 15u) | P.pw[4 * i], spread4((b16 >> 4) & 15u) | P.pw[4 * i + 1], spread4((b16 >> 8) & 15u) | P.pw[4 * i + 2],
                      spread4(b16 >> 12) | P.pw[4 * i + 3]};
    }
    obs_store32<WIDE>(dst + 16 * i0, v[0], v[1]);
  });
}

// the production table of player index tp (GameRec::prod, 50 x 4 bits) as 50 count bytes starting at byte offset O of the part
template <int O>
CATAN_FN void t_put_production(const GameView& g, int tp, PartImg& P) {
  static_for<0, 13>([&](auto K) {
    constexpr int k = decltype(K)::value;
    uint32_t e = nib4(g.prod(tp, k));
    if constexpr (k == 12) e &= 0x0000ffffu;                        // entries 48, 49
    pw_put<O + 4 * k, 40>(P.pw, e);
  });
}

struct ObsCtx { int actor, ap, aseat, lr_holder, la_holder; uint32_t relpack; };
CATAN_FN ObsCtx t_obs_ctx(const TCx& cx) {
  const GameView& g = cx.g;
  ObsCtx C;
  C.actor = t_current_actor(g); C.ap = C.actor - 1; C.aseat = seat_of(cx.s, C.actor);
  C.lr_holder = g.lr_holder(); C.la_holder = g.la_holder();
  C.relpack = 0;
#pragma unroll
  for (int p = 1; p <= 4; ++p) C.relpack |= static_cast<uint32_t>((seat_of(cx.s, p) - C.aseat + 4) & 3) << (2 * p);
  return C;
}

// road / army / harbour entries shared by both block kinds (wrapper.py:613-637): returns the bits [holder?, -, holder?, -, h0..h5]
// relative to the "longest road" entry and places the two count bytes at part offsets O + 1 and O + 3
template <int O>
CATAN_FN uint32_t t_road_army_harbours(const GameView& g, const ObsCtx& C, int target, PartImg& P) {
  const int tp = target - 1;
  uint32_t bits = 0, len = 0;
  if (C.lr_holder) {                                                 // Q9
    if (C.lr_holder == target) { bits |= 1u; len = g.lr_count(); }
    else if (g.has_path_key(tp)) len = g.cur_longest_path(tp);
  }
  if (C.la_holder == target) bits |= 4u;                             // Q10
  pw_put<O + 1, 40>(P.pw, len);
  pw_put<O + 3, 40>(P.pw, static_cast<uint32_t>(g.cur_army(tp)));
  return bits | (static_cast<uint32_t>(g.harbours(tp) & 63u) << 4);
}

// the acting player's block (152 bytes, wrapper.py:526-562, :587-637, :657-686), its first byte at part offset ST
template <int ST>
CATAN_FN void t_current_block(const TCx& cx, const ObsCtx& C, PartImg& P) {
  const GameView& g = cx.g;
  uint64_t b0 = 0, b1 = 0, b2 = 0;
#pragma unroll
  for (int r = 0; r < 5; ++r) b0 |= 1ull << (CATAN_OBS_RES_SLOT(r) * 8 + t_bucket8(g.res(C.ap, r)));       // [0, 40)
  const int vps = g.vp(C.ap);
  b0 |= 1ull << (40 + (vps < 10 ? vps : 9));                                                              // [40, 50)
  t_put_production<ST + 50>(g, C.ap, P);                                                                  // [50, 100)
  b1 |= static_cast<uint64_t>(t_road_army_harbours<ST + 100>(g, C, C.actor, P)) << (100 - 64);            // [100, 110)
  uint64_t bank = 0;
#pragma unroll
  for (int r = 0; r < 5; ++r) bank |= 1ull << (CATAN_OBS_RES_SLOT(r) * 7 + t_bucket7(g.bank(r)));         // [110, 145)
  bank |= 1ull << (35 + t_bucket7(g.deck_n()));                                                           // [145, 152)
  b1 |= bank << (110 - 64);
  b2 |= bank >> (128 - 110);
  img_or<ST>(P, b0, b1, b2);
}

// an opponent's block (159 bytes, wrapper.py:563-637, :688-695), REL = 1..3 seats after the actor, first byte at part offset ST
template <int REL, int ST>
CATAN_FN void t_other_block(const TCx& cx, const ObsCtx& C, PartImg& P) {
  const GameView& g = cx.g;
  const int target = pid_at_seat(cx.s, C.aseat + REL), tp = target - 1;
  uint64_t b0 = 0, b1 = 0, b2 = 0;
#pragma unroll
  for (int r = 0; r < 5; ++r) {
    const int slot = CATAN_OBS_RES_SLOT(r);
    b0 |= 1ull << (slot * 8 + t_bucket8(g.est_min(C.ap, REL - 1, r)));                                    // [0, 40)
    const int mx = 40 + slot * 8 + t_bucket8(g.est_max(C.ap, REL - 1, r));                                // [40, 80)
    if (slot < 3) b0 |= 1ull << mx; else b1 |= 1ull << (mx - 64);
  }
  const int vps = g.vp(tp);
  b1 |= 1ull << (80 - 64 + (vps < 10 ? vps : 9));                                                         // [80, 90)
  t_put_production<ST + 90>(g, tp, P);                                                                    // [90, 140)
  uint64_t tail = t_road_army_harbours<ST + 140>(g, C, target, P);                                        // [140, 150)
  tail |= 1ull << (10 + REL - 1);                                                                         // [150, 153) wrapper.py:532-541
  const int nh = g.n_hidden(tp);
  tail |= 1ull << (13 + (nh <= 4 ? nh : 5));                                                              // [153, 159) wrapper.py:690-695
  b2 |= tail << (140 - 128);
  img_or<ST>(P, b0, b1, b2);
}
// the first bytes of the NEXT opponent's block that fall into this part (its min-estimate one-hot of the slot-0 resource, Wood)
template <int REL, int ST>
CATAN_FN void t_other_block_head(const TCx& cx, const ObsCtx& C, PartImg& P) {
  const int b = t_bucket8(cx.g.est_min(C.ap, REL - 1, WOOD));
  if (ST + b < 160) P.w2 |= 1ull << (ST + b - 128);
}

// word k (cards 4k .. 4k+3) of a card list as obs bytes: card + 1, 0 beyond the list (wrapper.py:642-655)
CATAN_FN uint32_t t_list_word(uint32_t nibbles16, int n, int k) {
  const uint32_t valid = n >= 32 ? ~0u : ((1u << n) - 1u);
  return nib4(nibbles16) + spread4((valid >> (4 * k)) & 15u);
}
// the card list LI (0: actor played, 1: actor hidden, 2..4: opponents' played), its first byte at part offset O; words [K0, K1)
template <int LI, int O, int K0, int K1>
CATAN_FN int t_put_list(const TCx& cx, const ObsCtx& C, PartImg& P) {
  const GameView& g = cx.g;
  const int tp = (LI < 2 ? C.actor : pid_at_seat(cx.s, C.aseat + LI - 1)) - 1;
  const int n = LI == 1 ? g.n_hidden(tp) : g.n_played(tp);
  static_for<K0, K1>([&](auto K) {
    constexpr int k = decltype(K)::value;
    uint32_t x = LI == 1 ? g.hidden(tp, 2 * k) : g.played(tp, 2 * k);
    if constexpr (k < 6) x |= static_cast<uint32_t>(LI == 1 ? g.hidden(tp, 2 * k + 1) : g.played(tp, 2 * k + 1)) << 8;
    uint32_t e = t_list_word(x, n, k);
    if constexpr (k == 6) e &= 0xffu;                               // card 24 is the last one
    pw_put<O + 4 * k, 40>(P.pw, e);
  });
  return n;
}

// ------------------------------------------------------------------------------------------------
// Bytes [lo, hi) of the row inside the header + tile region [0, 1152) -- 60 % of the row -- without the window: all of
// these features are 0/1 bytes, so a tile is first built as a 60-bit set in registers (bit j = byte j of the tile's
// block) and the stream of bits is then expanded 16 bits -> 16 bytes per vector store.  The five hand counts of the
// header are patched into the two pieces they fall in.
// ------------------------------------------------------------------------------------------------
CATAN_FN uint64_t t_tile_bits(const GameView& g, const Topo& T, int t, int robber, uint32_t relpack) {
  uint32_t b[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) b[k] = g.corner(T.tile_corners[t][k]);
  const int val = g.tile_val(t), tres = g.tile_res(t);
  uint32_t w0 = static_cast<uint32_t>(robber == t) | (1u << (val - 1)) | (1u << (12 + tres)), w1 = 0;   // wrapper.py:497-509
#pragma unroll
  for (int k = 0; k < 6; ++k) {                                      // wrapper.py:510-523: none/settlement/city, owner relative to the actor
    const uint32_t rel = (relpack >> (2 * (b[k] >> 2))) & 3u;
    const uint32_t f = (1u << (b[k] & 3u)) | (static_cast<uint32_t>(b[k] != 0) << (3 + rel));   // the corner's 7 bytes
    if (k < 2) w0 |= f << (18 + 7 * k); else w1 |= f << (7 * (k - 2));
  }
  return static_cast<uint64_t>(w0) | (static_cast<uint64_t>(w1) << 32);
}

template <bool WIDE>
CATAN_FN void t_encode_obs_tiles_inl(const TCx& cx_, uint8_t* row, int lo, int hi) {
  TCx cx = cx_;   // (a private copy: the address-space hints below then hold for every access, whatever the stores in between)
  CATAN_STAGED_ENCODE(cx);
  const GameView& g = cx.g;
  const Topo& T = *cx.T;
  const int actor = t_current_actor(g), ap = actor - 1, aseat = seat_of(cx.s, actor), robber = g.robber_tile();
  uint32_t relpack = 0;
#pragma unroll
  for (int p = 1; p <= 4; ++p) relpack |= static_cast<uint32_t>((seat_of(cx.s, p) - aseat + 4) & 3) << (2 * p);
  uint64_t acc_lo = 0, acc_hi = 0;       // pending bits: bit i = byte pos + i of the row
  int fill = 0, pos = lo, t = 0;
  uint32_t hand012 = 0, hand34 = 0;
  if (lo == 0) {                                                     // [0, 18): proposed trade (wrapper.py:61-69, Q15), hand (:70-71)
    uint32_t tr = 0;
    if (g.trade_proposer()) {
      const int ng = g.n_give(), nr = g.n_recv();
      for (int k = 0; k < ng; ++k) tr |= 1u << (CATAN_OBS_PROPOSED_TRADE + g.give(k));
      for (int k = 0; k < nr; ++k) tr |= 1u << (CATAN_OBS_PROPOSED_TRADE + g.recv(k) + 5);
    }
    acc_lo = tr; fill = CATAN_OBS_TILES;
    hand012 = (static_cast<uint32_t>(g.res(ap, 0)) << 8) | (static_cast<uint32_t>(g.res(ap, 1)) << 16) | (static_cast<uint32_t>(g.res(ap, 2)) << 24);
    hand34 = static_cast<uint32_t>(g.res(ap, 3)) | (static_cast<uint32_t>(g.res(ap, 4)) << 8);
  } else {                                                           // start inside tile t: drop its bytes below lo
    t = (lo - CATAN_OBS_TILES) / CATAN_OBS_TILE_DIM;
    const int skip = lo - (CATAN_OBS_TILES + t * CATAN_OBS_TILE_DIM);
    acc_lo = t_tile_bits(g, T, t, robber, relpack) >> skip;
    fill = CATAN_OBS_TILE_DIM - skip;
    ++t;
  }
  // (lo and hi are multiples of 32: two 16-byte pieces per iteration, one 32-byte store)
  CATAN_NO_UNROLL
  while (pos < hi) {
    if (fill < 32) {                                                 // (t < 19 here: the region ends inside tile 18)
      const uint64_t m = t_tile_bits(g, T, t, robber, relpack);
      ++t;
      acc_lo |= m << fill;
      acc_hi = fill > 4 ? m >> (64 - fill) : 0ull;
      fill += CATAN_OBS_TILE_DIM;
    }
    const uint32_t b32 = static_cast<uint32_t>(acc_lo);
    ObsPiece v0 = {spread4(b32 & 15u), spread4((b32 >> 4) & 15u), spread4((b32 >> 8) & 15u), spread4((b32 >> 12) & 15u)};
    ObsPiece v1 = {spread4((b32 >> 16) & 15u), spread4((b32 >> 20) & 15u), spread4((b32 >> 24) & 15u), spread4(b32 >> 28)};
    if (pos == 0) { v0.d |= hand012; v1.a |= hand34; }               // bytes 13..15 = CATAN_OBS_CURRENT_RES + 1 + r, then bytes 16, 17
    obs_store32<WIDE>(row + pos, v0, v1);
    acc_lo = (acc_lo >> 32) | (acc_hi << 32);
    acc_hi >>= 32;
    fill -= 32;
    pos += 32;
  }
}
// (the called form: the kernels that hold several phases keep their code small; the rows launch of a step inlines it -- a call
// forces the context through local memory, which misses the small L1 these kernels leave beside their shared memory)
CATAN_FN_NOINLINE void t_encode_obs_tiles(const TCx& cx, uint8_t* row, int lo, int hi) { t_encode_obs_tiles_inl<false>(cx, row, lo, hi); }
static_assert(CATAN_OBS_PROPOSED_TRADE == 0 && CATAN_OBS_CURRENT_RES == 12 && CATAN_OBS_TILES == 18 && CATAN_OBS_TILE_DIM == 60,
              "t_encode_obs_tiles packs the header by hand");

// The cuts used by the device encoder (multiples of 32 bytes: obs_store32): CATAN_OBS_PARTS threads share one row.  Parts 0 .. CATAN_OBS_TILE_PARTS-1 lie in
// the header + tile region (t_encode_obs_tiles); parts 4-7 hold one player block each (plus the few bytes of its neighbours
// that share its first / last 16-byte piece), part 8 the card lists and the meta bytes.
#define CATAN_OBS_PARTS 9
#define CATAN_OBS_TILE_PARTS 4
#define CATAN_OBS_TILE_END 1152
CATAN_FN int t_obs_part_lo(int part) {
  return part == 0 ? 0 : part == 1 ? 288 : part == 2 ? 576 : part == 3 ? 864 : part == 4 ? CATAN_OBS_TILE_END : part == 5 ? 1312 :
         part == 6 ? 1472 : part == 7 ? 1632 : part == 8 ? 1792 : CATAN_OBS_STRIDE;
}
template <bool WIDE>
CATAN_FN void t_encode_obs_players(const TCx& cx, uint8_t* row, int part) {
  const ObsCtx C = t_obs_ctx(cx);
  PartImg P = {};
  constexpr int OTH = CATAN_OBS_OTHER_MAIN, OD = CATAN_OBS_OTHER_MAIN_DIM, LISTS = CATAN_OBS_DEV_LISTS, PAD = CATAN_OBS_DEV_PAD;
  switch (part) {
    case 4: {                                                        // [1152, 1312): tail of tile 18, the actor's block, 2 bytes of the next
      constexpr int lo = 1152;
      P.w0 = t_tile_bits(cx.g, *cx.T, 18, cx.g.robber_tile(), C.relpack) >> (lo - (CATAN_OBS_TILES + 18 * CATAN_OBS_TILE_DIM));
      t_current_block<CATAN_OBS_CUR_MAIN - lo>(cx, C, P);
      t_other_block_head<1, OTH - lo>(cx, C, P);
      img_store<10, WIDE>(P, row + lo);
      break;
    }
    case 5: {
      constexpr int lo = 1312;
      t_other_block<1, OTH - lo>(cx, C, P);
      t_other_block_head<2, OTH + OD - lo>(cx, C, P);
      img_store<10, WIDE>(P, row + lo);
      break;
    }
    case 6: {
      constexpr int lo = 1472;
      t_other_block<2, OTH + OD - lo>(cx, C, P);
      t_other_block_head<3, OTH + 2 * OD - lo>(cx, C, P);
      img_store<10, WIDE>(P, row + lo);
      break;
    }
    case 7: {                                                        // ... and the first 5 cards of the actor's played list
      constexpr int lo = 1632;
      t_other_block<3, OTH + 2 * OD - lo>(cx, C, P);
      t_put_list<0, LISTS - lo, 0, 2>(cx, C, P);
      img_store<10, WIDE>(P, row + lo);
      break;
    }
    default: {                                                       // [1792, 1920): the card lists and the meta bytes
      constexpr int lo = 1792;
      const int n0 = t_put_list<0, LISTS - lo, 1, 7>(cx, C, P);
      const int n1 = t_put_list<1, LISTS + PAD - lo, 0, 7>(cx, C, P);
      const int n2 = t_put_list<2, LISTS + 2 * PAD - lo, 0, 7>(cx, C, P);
      const int n3 = t_put_list<3, LISTS + 3 * PAD - lo, 0, 7>(cx, C, P);
      const int n4 = t_put_list<4, LISTS + 4 * PAD - lo, 0, 7>(cx, C, P);
      P.pw[(CATAN_OBS_META - lo) / 4] = static_cast<uint32_t>(C.actor) | (static_cast<uint32_t>(n0) << 8) | (static_cast<uint32_t>(n1) << 16) |
                                        (static_cast<uint32_t>(n2) << 24);
      P.pw[(CATAN_OBS_META - lo) / 4 + 1] = static_cast<uint32_t>(n3) | (static_cast<uint32_t>(n4) << 8);
      img_store<8, WIDE>(P, row + lo);
      break;
    }
  }
}
static_assert(CATAN_OBS_CUR_MAIN == 1158 && CATAN_OBS_OTHER_MAIN == 1310 && CATAN_OBS_OTHER_MAIN_DIM == 159 && CATAN_OBS_DEV_LISTS == 1787 &&
              CATAN_OBS_META == 1912 && CATAN_OBS_STRIDE == 1920 && CATAN_OBS_DEV_PAD == 25, "the part cuts assume this row layout");

template <bool WIDE = false>
CATAN_FN void t_encode_obs_part(const TCx& cx, uint8_t* row, int part) {
  if (part < CATAN_OBS_TILE_PARTS) t_encode_obs_tiles(cx, row, t_obs_part_lo(part), t_obs_part_lo(part + 1));
  else t_encode_obs_players<WIDE>(cx, row, part);
}

}  // namespace catanb
