/*
 * TEST INFRASTRUCTURE — CPU oracle (see catan_oracle.h).  Scalar C restatement of the reference's
 * game/game.py + env/wrapper.py + game/components/ *.py.  Every function cites the lines it follows.
 * Parity status: PINNED — checked step by step against the real reference (state, obs, masks,
 * rewards, done) by tests/test_oracle_vs_reference.py and against tests/golden/ fixtures.
 */
#include "catan_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define CATAN_TOPO_CONST static const
#include "../include/catan_topology.h"

typedef catan_state_t S;

/* PlayerId constants (enums.py:8-12) */
enum { WHITE = 1, BLUE = 2, ORANGE = 3, RED = 4 };
/* resource indices r = Resource - 1 */
enum { BRICK = 0, WOOD = 1, ORE = 2, SHEEP = 3, WHEAT = 4 };

void catan_oracle_default_config(catan_config_t* c) {
  c->max_actions_per_turn = -1;         /* wrapper.py:14-17 (None -> inf) */
  c->max_proposed_trades_per_turn = 4;  /* wrapper.py:12 */
  c->validate_actions = 1;
  c->dense_reward = 0;
  c->auto_reset = 0;
  c->win_reward = 500.0f;
  c->reward_annealing_factor = 1.0f;    /* wrapper.py:28 */
}

int catan_oracle_state_words(void) { return CATAN_STATE_WORDS; }

/* ------------------------------------------------------------------------------------------
 * pinned RNG (catan_layout.h)
 * ------------------------------------------------------------------------------------------ */
void catan_oracle_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                         uint32_t out[4]) {
  for (int i = 0; i < 10; ++i) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct { uint64_t seed, env_id; S* s; } Rng;

static uint32_t rng_ctr_get(const S* s) { return (uint32_t)(uint16_t)s->rng_ctr_lo | ((uint32_t)(uint16_t)s->rng_ctr_hi << 16); }
static void rng_ctr_set(S* s, uint32_t c) { s->rng_ctr_lo = (int16_t)(c & 0xFFFF); s->rng_ctr_hi = (int16_t)(c >> 16); }

static uint32_t rng_next(Rng* g) {
  uint32_t d = rng_ctr_get(g->s);
  rng_ctr_set(g->s, d + 1);
  uint32_t w[4];
  catan_oracle_philox(d >> 2, CATAN_STREAM_GAME, (uint32_t)g->env_id, (uint32_t)(g->env_id >> 32),
                      (uint32_t)g->seed, (uint32_t)(g->seed >> 32), w);
  return w[d & 3];
}
static int rng_bounded(Rng* g, int n) { return (int)(((uint64_t)rng_next(g) * (uint64_t)n) >> 32); }
static void rng_shuffle(Rng* g, int16_t* a, int n) {
  for (int i = n - 1; i >= 1; --i) {
    int j = rng_bounded(g, i + 1);
    int16_t t = a[i]; a[i] = a[j]; a[j] = t;
  }
}

/* ------------------------------------------------------------------------------------------
 * seats (player.py:13-19)
 * ------------------------------------------------------------------------------------------ */
static int seat_of(const S* s, int pid) {
  for (int i = 0; i < 4; ++i) if (s->player_order[i] == pid) return i;
  return 0;
}
/* player_lookup[pid_b] as seen from pid_a: 0 next, 1 next_next, 2 next_next_next */
static int label_of(const S* s, int a, int b) { return (seat_of(s, b) - seat_of(s, a) + 4) % 4 - 1; }
/* inverse_player_lookup[label] as seen from pid_a */
static int pid_at_label(const S* s, int a, int label) { return s->player_order[(seat_of(s, a) + 1 + label) % 4]; }

static int hand_total(const S* s, int pid) {
  int t = 0;
  for (int r = 0; r < 5; ++r) t += s->res[pid - 1][r];
  return t;
}
static int clipi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* ------------------------------------------------------------------------------------------
 * Board.validate_number_order (board.py:50-65) and reset (board.py:67-167, game.py:39-136)
 * ------------------------------------------------------------------------------------------ */
static int validate_number_order(const int16_t* numbers, const int16_t* terrain) {
  int vals[19];
  int n = 0;
  for (int i = 0; i < 19; ++i) {
    int t = CATAN_NUMBER_PLACEMENT[i];
    if (terrain[t] == 0) vals[t] = 7; else vals[t] = numbers[n++];
  }
  for (int i = 0; i < 19; ++i) {
    if (vals[i] == 6 || vals[i] == 8) {
      for (int k = 0; k < 6; ++k) {
        int nb = CATAN_TILE_NEIGH[i][k];
        if (nb >= 0 && (vals[nb] == 6 || vals[nb] == 8)) return 0;
      }
    }
  }
  return 1;
}

void catan_oracle_reset(S* s, uint64_t seed, uint64_t env_id) {
  uint32_t ctr = rng_ctr_get(s);
  memset(s, 0, sizeof(*s));
  rng_ctr_set(s, ctr);
  Rng g = {seed, env_id, s};
  int16_t terrain[19], numbers[18], harbours[9];
  for (int i = 0; i < 19; ++i) terrain[i] = CATAN_TERRAIN_TO_PLACE[i];
  rng_shuffle(&g, terrain, 19);                                   /* board.py:71-72 */
  for (int i = 0; i < 18; ++i) numbers[i] = CATAN_DEFAULT_NUMBER_ORDER[i];
  rng_shuffle(&g, numbers, 18);                                   /* board.py:79 */
  while (!validate_number_order(numbers, terrain)) rng_shuffle(&g, numbers, 18); /* board.py:80-81 */
  for (int i = 0; i < 9; ++i) harbours[i] = (int16_t)i;
  rng_shuffle(&g, harbours, 9);                                   /* board.py:83-84 */
  for (int i = 0; i < 9; ++i) s->harbour_perm[i] = harbours[i];
  int n = 0;
  for (int i = 0; i < 19; ++i) {                                  /* board.py:88-100 */
    int t = CATAN_NUMBER_PLACEMENT[i];
    s->tile_res[t] = terrain[t];
    if (terrain[t] == 0) { s->tile_val[t] = 7; s->robber_tile = (int16_t)t; }
    else s->tile_val[t] = numbers[n++];
  }
  int16_t order[4] = {WHITE, BLUE, ORANGE, RED};                  /* game.py:41-42 */
  rng_shuffle(&g, order, 4);
  for (int i = 0; i < 4; ++i) s->player_order[i] = order[i];
  s->players_go = order[0];
  s->player_order_id = 0;
  for (int r = 0; r < 5; ++r) s->bank[r] = 19;                    /* game.py:48-54 */
  for (int p = 0; p < 4; ++p) {
    s->settlements_left[p] = 5; s->cities_left[p] = 4;            /* game.py:55-68 */
    s->second_corner[p] = -1;
  }
  for (int i = 0; i < 25; ++i) s->deck[i] = CATAN_DECK_INIT[i];   /* game.py:75-78 */
  rng_shuffle(&g, s->deck, 25);
  s->deck_n = 25;
  s->initial_phase = 1;                                           /* game.py:84 */
  /* wrapper.py:32-33: winner None, curr_vps 0 — already zero */
}

/* ------------------------------------------------------------------------------------------
 * placement predicates (corner.py:24-39, edge.py:23-42)
 * ------------------------------------------------------------------------------------------ */
static int can_place_settlement(const S* s, int c, int pid, int initial) {
  int roads = 0;
  if (s->corner_type[c]) return 0;
  for (int k = 0; k < 3; ++k) {
    int nb = CATAN_CORNER_NEIGH[c][k];
    if (nb < 0) continue;
    if (s->corner_type[nb]) return 0;
    if (s->edge_owner[CATAN_CORNER_NEIGH_EDGE[c][k]] == pid) roads++;
  }
  if (initial) return 1;
  return roads > 0;
}

static int can_place_road(const S* s, int e, int pid, int after_second, int second_corner) {
  if (s->edge_owner[e]) return 0;
  int c1 = CATAN_EDGE_CORNERS[e][0], c2 = CATAN_EDGE_CORNERS[e][1];
  if (after_second) return (c1 == second_corner || c2 == second_corner);
  if ((s->corner_type[c1] && s->corner_owner[c1] == pid) || (s->corner_type[c2] && s->corner_owner[c2] == pid)) return 1;
  int cs[2] = {c1, c2};
  for (int i = 0; i < 2; ++i) {
    int c = cs[i];
    for (int k = 0; k < 3; ++k) {
      int ne = CATAN_CORNER_NEIGH_EDGE[c][k];
      if (ne < 0) continue;
      if (s->edge_owner[ne] == pid && !s->corner_type[c]) return 1;
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * belief updates (game.py:921-971, :973-1010)
 * ------------------------------------------------------------------------------------------ */
static void est_update(S* s, const int* delta, const int* touched, int owner, int thief) {
  int T_o = hand_total(s, owner);
  int T_t = thief ? hand_total(s, thief) : 0;
  static const int observers[4] = {WHITE, RED, BLUE, ORANGE};     /* game.py:925 */
  for (int oi = 0; oi < 4; ++oi) {
    int o = observers[oi];
    if (o == owner) {
      if (!thief) continue;
      int l = label_of(s, o, thief);                              /* game.py:929-933 (unclipped) */
      for (int r = 0; r < 5; ++r) if (touched[r]) {
        s->est_max[o - 1][l][r] -= (int16_t)delta[r];
        s->est_min[o - 1][l][r] -= (int16_t)delta[r];
      }
    } else {
      int l = label_of(s, o, owner);
      if (!thief || o == thief) {                                 /* game.py:936-954 */
        for (int r = 0; r < 5; ++r) if (touched[r]) {
          s->est_max[o - 1][l][r] = (int16_t)clipi(s->est_max[o - 1][l][r] + delta[r], 0, T_o);
          s->est_min[o - 1][l][r] = (int16_t)clipi(s->est_min[o - 1][l][r] + delta[r], 0, T_o);
        }
      } else {                                                    /* game.py:955-971 */
        int lt = label_of(s, o, thief);
        for (int r = 0; r < 5; ++r) {
          int m0 = s->est_max[o - 1][l][r];
          s->est_max[o - 1][l][r] = (int16_t)clipi(m0, 0, T_o);
          s->est_min[o - 1][l][r] = (int16_t)clipi(s->est_min[o - 1][l][r] - 1, 0, T_o);
          if (m0 > 0) {
            s->est_max[o - 1][lt][r] = (int16_t)clipi(s->est_max[o - 1][lt][r] + 1, 0, T_t);
            s->est_min[o - 1][lt][r] = (int16_t)clipi(s->est_min[o - 1][lt][r], 0, T_t);
          }
        }
      }
    }
  }
}

static void est_single(S* s, int r, int d, int owner) {
  int delta[5] = {0, 0, 0, 0, 0}, touched[5] = {0, 0, 0, 0, 0};
  delta[r] = d; touched[r] = 1;
  est_update(s, delta, touched, owner, 0);
}

static void est_monopoly(S* s, int mono, int res, const int* lost /* by player index */) {
  int tot = 0;
  for (int p = 1; p <= 4; ++p) if (p != mono) tot += lost[p - 1];
  for (int p = 1; p <= 4; ++p) {                                  /* iteration order irrelevant: disjoint entries */
    if (p == mono) {
      for (int o = 1; o <= 4; ++o) {                              /* game.py:984-991 (unclipped) */
        if (o == p) continue;
        int l = label_of(s, o, p);
        s->est_min[o - 1][l][res] += (int16_t)tot;
        s->est_max[o - 1][l][res] += (int16_t)tot;
      }
    } else {
      int T_p = hand_total(s, p);                                 /* game.py:993-1010 */
      for (int o = 1; o <= 4; ++o) {
        if (o == p) continue;
        int l = label_of(s, o, p);
        for (int r = 0; r < 5; ++r) {
          int mx = s->est_max[o - 1][l][r], mn = s->est_min[o - 1][l][r];
          if (r == res) { mn -= lost[p - 1]; mx -= lost[p - 1]; }
          s->est_max[o - 1][l][r] = (int16_t)clipi(mx, 0, T_p);
          s->est_min[o - 1][l][r] = (int16_t)clipi(mn, 0, T_p);
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * longest road (game.py:843-919, utils.py:3-15)
 * ------------------------------------------------------------------------------------------ */
typedef struct { int n_out[54]; int out[54][3]; } Graph;

static int dfs(const Graph* G, int v, uint64_t seen, int depth) {
  /* utils.py:3-15: node-simple paths; returns the max number of arcs below v */
  seen |= (1ull << v);
  int best = depth;
  for (int k = 0; k < G->n_out[v]; ++k) {
    int t = G->out[v][k];
    if (seen & (1ull << t)) continue;
    int d = dfs(G, t, seen, depth + 1);
    if (d > best) best = d;
  }
  return best;
}

int catan_oracle_longest_path(const S* s, int pid) {
  Graph G;
  memset(&G, 0, sizeof(G));
  for (int e = 0; e < 72; ++e) {                                  /* game.py:845-858 */
    if (s->edge_owner[e] != pid) continue;
    int a = CATAN_EDGE_CORNERS[e][0], b = CATAN_EDGE_CORNERS[e][1];
    if (!(s->corner_type[a] && s->corner_owner[a] != pid)) G.out[a][G.n_out[a]++] = b;
    if (!(s->corner_type[b] && s->corner_owner[b] != pid)) G.out[b][G.n_out[b]++] = a;
  }
  int best = 0;
  for (int v = 0; v < 54; ++v) {
    if (!G.n_out[v]) continue;
    int d = dfs(&G, v, 0, 0);
    if (d > best) best = d;
  }
  return best;
}

static void update_longest_road(S* s, int pid) {                  /* game.py:864-919 */
  int len = catan_oracle_longest_path(s, pid);
  s->cur_longest_path[pid - 1] = (int16_t)len;
  s->has_path_key[pid - 1] = 1;
  if (!s->lr_holder) {
    if (len >= 5) { s->lr_holder = (int16_t)pid; s->lr_count = (int16_t)len; s->vp[pid - 1] += 2; }
    return;
  }
  if (s->lr_holder == pid) {
    if (s->lr_count > len) {
      int max_len = len, player = pid, tied = 0;
      static const int order[4] = {WHITE, BLUE, ORANGE, RED};     /* game.py:886 */
      for (int i = 0; i < 4; ++i) {
        int o = order[i];
        if (o == pid) continue;
        int pl = catan_oracle_longest_path(s, o);
        if (pl == max_len) tied = 1;
        else if (pl > max_len) { max_len = pl; tied = 0; player = o; }
      }
      if (max_len >= 5) {
        if (tied) {
          if (player == pid) { s->lr_count = (int16_t)len; }
          else { s->lr_holder = 0; s->lr_count = 0; s->vp[pid - 1] -= 2; }
        } else {
          s->lr_holder = (int16_t)player; s->lr_count = (int16_t)max_len;
          s->vp[player - 1] += 2; s->vp[pid - 1] -= 2;
        }
      } else { s->lr_holder = 0; s->lr_count = 0; s->vp[pid - 1] -= 2; }
    } else {
      s->lr_count = (int16_t)len;
    }
  } else if (len > s->lr_count) {
    s->vp[s->lr_holder - 1] -= 2; s->vp[pid - 1] += 2;
    s->lr_holder = (int16_t)pid; s->lr_count = (int16_t)len;
  }
}

static void update_largest_army(S* s) {                           /* game.py:817-841 */
  static const int order[4] = {BLUE, WHITE, RED, ORANGE};
  int max_count = 0, cp = 0;
  for (int i = 0; i < 4; ++i) {
    int p = order[i], k = 0;
    for (int j = 0; j < s->n_played[p - 1]; ++j) if (s->played[p - 1][j] == CATAN_DEV_KNIGHT) k++;
    s->cur_army[p - 1] = (int16_t)k;
    if (k >= 3 && k > max_count) { max_count = k; cp = p; }
  }
  if (!cp) return;
  if (!s->la_holder) { s->la_holder = (int16_t)cp; s->la_count = (int16_t)max_count; s->vp[cp - 1] += 2; }
  else if (s->la_holder == cp) s->la_count = (int16_t)max_count;
  else if (max_count > s->la_count) {
    s->vp[s->la_holder - 1] -= 2; s->la_holder = (int16_t)cp; s->la_count = (int16_t)max_count; s->vp[cp - 1] += 2;
  }
}

/* ------------------------------------------------------------------------------------------
 * helpers for apply_action
 * ------------------------------------------------------------------------------------------ */
static void pay(S* s, int pid, int r, int n) {                    /* hand -n, visible floor 0, bank +n */
  int p = pid - 1;
  s->res[p][r] -= (int16_t)n;
  s->vis[p][r] = (int16_t)(s->vis[p][r] - n > 0 ? s->vis[p][r] - n : 0);
  s->bank[r] += (int16_t)n;
}

static void update_players_go(S* s, int left) {                   /* game.py:253-262 */
  if (left) { s->player_order_id -= 1; if (s->player_order_id < 0) s->player_order_id = 3; }
  else { s->player_order_id += 1; if (s->player_order_id > 3) s->player_order_id = 0; }
  s->players_go = s->player_order[s->player_order_id];
}

static int best_exchange_rate(const S* s, int pid, int r) {       /* wrapper.py:428-438 */
  int h = s->harbours[pid - 1];
  if (h & (1 << (r + 1))) return 2;
  if (h & 1) return 3;
  return 4;
}

static int count_card(const int16_t* list, int n, int card) {
  int k = 0;
  for (int i = 0; i < n; ++i) if (list[i] == card) k++;
  return k;
}

static void roll_dice(S* s, Rng* g) {                             /* game.py:138-177 */
  s->die1 = (int16_t)(1 + rng_bounded(g, 6));
  s->die2 = (int16_t)(1 + rng_bounded(g, 6));
  int roll = s->die1 + s->die2;
  if (roll == 7) {
    for (int i = 0; i < 4; ++i) {
      int pid = s->player_order[i];
      if (hand_total(s, pid) > 7) { s->need_discard = 1; s->discard_queue[s->n_discard++] = (int16_t)pid; }
    }
    return;
  }
  int alloc[5][4];
  int total[5] = {0, 0, 0, 0, 0};
  memset(alloc, 0, sizeof(alloc));
  for (int t = 0; t < 19; ++t) {
    if (s->tile_val[t] != roll || t == s->robber_tile) continue;
    int r = s->tile_res[t] - 1;
    for (int k = 0; k < 6; ++k) {
      int c = CATAN_TILE_CORNERS[t][k];
      if (!s->corner_type[c]) continue;
      int inc = s->corner_type[c];                                /* settlement 1, city 2 */
      alloc[r][s->corner_owner[c] - 1] += inc;
      total[r] += inc;
    }
  }
  static const int res_order[5] = {WOOD, ORE, BRICK, WHEAT, SHEEP};   /* game.py:153-155 */
  static const int pl_order[4] = {BLUE, ORANGE, WHITE, RED};         /* game.py:172 */
  for (int ri = 0; ri < 5; ++ri) {
    int r = res_order[ri];
    if (total[r] <= s->bank[r]) {
      for (int pi = 0; pi < 4; ++pi) {
        int pid = pl_order[pi];
        s->res[pid - 1][r] += (int16_t)alloc[r][pid - 1];
        s->bank[r] -= (int16_t)alloc[r][pid - 1];
        est_single(s, r, alloc[r][pid - 1], pid);                 /* also when the gain is 0 (Q4) */
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * translated action (wrapper.py:114-166, :414-486)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int type, corner, edge /* -1 = None */, tile, card, accept, target_pid, res_a, res_b, rate;
  int n_give, give[4], n_recv, recv[4];   /* resource indices */
  int discard;
} Act;

static int translate(const S* s, const int32_t* a, Act* t) {
  memset(t, 0, sizeof(*t));
  t->type = a[CATAN_A_TYPE];
  int pg = s->players_go;
  switch (t->type) {
    case CATAN_ACT_PLACE_SETTLEMENT: case CATAN_ACT_UPGRADE_CITY:
      t->corner = a[CATAN_A_CORNER];
      if (t->corner < 0 || t->corner >= 54) return CATAN_ERR_BAD_HEAD_VALUE;
      break;
    case CATAN_ACT_PLACE_ROAD:
      if (a[CATAN_A_EDGE] < 0 || a[CATAN_A_EDGE] > 72) return CATAN_ERR_BAD_HEAD_VALUE;
      t->edge = a[CATAN_A_EDGE] == 72 ? -1 : a[CATAN_A_EDGE];
      break;
    case CATAN_ACT_MOVE_ROBBER:
      t->tile = a[CATAN_A_TILE];
      if (t->tile < 0 || t->tile >= 19) return CATAN_ERR_BAD_HEAD_VALUE;
      break;
    case CATAN_ACT_STEAL:
      if (a[CATAN_A_PLAYER] < 0 || a[CATAN_A_PLAYER] > 2) return CATAN_ERR_BAD_HEAD_VALUE;
      t->target_pid = pid_at_label(s, pg, a[CATAN_A_PLAYER]);
      break;
    case CATAN_ACT_PLAY_DEV:
      t->card = a[CATAN_A_CARD];
      if (t->card < 0 || t->card > 4) return CATAN_ERR_BAD_HEAD_VALUE;
      if (t->card == CATAN_DEV_MONOPOLY || t->card == CATAN_DEV_YOP) {
        t->res_a = a[CATAN_A_RES_A];
        if (t->res_a < 0 || t->res_a > 4) return CATAN_ERR_BAD_HEAD_VALUE;
      }
      if (t->card == CATAN_DEV_YOP) {
        t->res_b = a[CATAN_A_RES_B];
        if (t->res_b < 0 || t->res_b > 4) return CATAN_ERR_BAD_HEAD_VALUE;
      }
      break;
    case CATAN_ACT_EXCHANGE:
      t->res_a = a[CATAN_A_RES_A]; t->res_b = a[CATAN_A_RES_B];
      if (t->res_a < 0 || t->res_a > 4 || t->res_b < 0 || t->res_b > 4) return CATAN_ERR_BAD_HEAD_VALUE;
      t->rate = best_exchange_rate(s, pg, t->res_a);
      break;
    case CATAN_ACT_PROPOSE_TRADE:
      if (a[CATAN_A_PLAYER] < 0 || a[CATAN_A_PLAYER] > 2) return CATAN_ERR_BAD_HEAD_VALUE;
      t->target_pid = pid_at_label(s, pg, a[CATAN_A_PLAYER]);
      for (int k = 0; k < 4; ++k) {                               /* wrapper.py:451-466: 0 stops the list */
        int v = a[CATAN_A_GIVE + k];
        if (v == 0) break;
        if (v < 0 || v > 5) return CATAN_ERR_BAD_HEAD_VALUE;
        t->give[t->n_give++] = v - 1;
      }
      for (int k = 0; k < 4; ++k) {
        int v = a[CATAN_A_RECV + k];
        if (v == 0) break;
        if (v < 0 || v > 5) return CATAN_ERR_BAD_HEAD_VALUE;
        t->recv[t->n_recv++] = v - 1;
      }
      break;
    case CATAN_ACT_RESPOND:
      t->accept = a[CATAN_A_ACCEPT];
      if (t->accept < 0 || t->accept > 1) return CATAN_ERR_BAD_HEAD_VALUE;
      break;
    case CATAN_ACT_DISCARD:
      t->discard = a[CATAN_A_DISCARD];
      if (t->discard < 0 || t->discard > 4) return CATAN_ERR_BAD_HEAD_VALUE;
      break;
    case CATAN_ACT_BUY_DEV: case CATAN_ACT_ROLL_DICE: case CATAN_ACT_END_TURN:
      break;
    default:
      return CATAN_ERR_BAD_TYPE;
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Game.validate_action (game.py:264-525)
 * ------------------------------------------------------------------------------------------ */
static int validate(const S* s, const Act* t) {
  int pid = s->players_go, p = pid - 1;
  const int16_t* h = s->res[p];
  if (s->need_discard) {                                          /* game.py:279-300 */
    if (t->type != CATAN_ACT_DISCARD) return CATAN_ERR_PHASE;
    int cur = s->discard_queue[0];
    if (hand_total(s, cur) <= 7) return CATAN_ERR_PHASE;
    if (s->res[cur - 1][t->discard] <= 0) return CATAN_ERR_BAD_RESOURCE;
    return 0;
  } else if (t->type == CATAN_ACT_DISCARD) return CATAN_ERR_PHASE;

  switch (t->type) {
    case CATAN_ACT_PLACE_SETTLEMENT:                              /* game.py:305-323 */
      if (s->must_respond) return CATAN_ERR_PHASE;
      if (!s->dice_rolled && !s->initial_phase) return CATAN_ERR_PHASE;
      if (s->must_use_dev || s->just_moved_robber) return CATAN_ERR_PHASE;
      if (s->initial_phase || (s->settlements_left[p] > 0 && h[WHEAT] > 0 && h[WOOD] > 0 && h[BRICK] > 0 && h[SHEEP] > 0)) {
        if (can_place_settlement(s, t->corner, pid, s->initial_phase)) {
          if (s->initial_phase) {
            if (s->init_settlements[p] == 0 || (s->init_settlements[p] == 1 && s->init_roads[p] == 1)) return 0;
            return CATAN_ERR_BAD_LOCATION;
          }
          return 0;
        }
      }
      return CATAN_ERR_CANNOT_AFFORD;
    case CATAN_ACT_PLACE_ROAD:                                    /* game.py:324-357 */
      if (s->rb_active) {
        if (t->edge < 0) return 0;
        return can_place_road(s, t->edge, pid, 0, 0) ? 0 : CATAN_ERR_BAD_LOCATION;
      }
      if (s->must_respond) return CATAN_ERR_PHASE;
      if (!s->dice_rolled && !s->initial_phase) return CATAN_ERR_PHASE;
      if (s->must_use_dev || s->just_moved_robber) return CATAN_ERR_PHASE;
      if (s->initial_phase || (h[WOOD] > 0 && h[BRICK] > 0)) {
        if (t->edge < 0) return CATAN_ERR_BAD_LOCATION;           /* reference: TypeError on edges[None] */
        if (can_place_road(s, t->edge, pid, 0, 0)) {
          if (s->initial_phase) {
            if (s->init_settlements[p] == 1 && s->init_roads[p] == 0) return 0;
            if (s->init_settlements[p] == 2 && s->init_roads[p] == 1)
              return can_place_road(s, t->edge, pid, 1, s->second_corner[p]) ? 0 : CATAN_ERR_BAD_LOCATION;
            return CATAN_ERR_BAD_LOCATION;
          }
          return 0;
        }
        return CATAN_ERR_BAD_LOCATION;
      }
      return CATAN_ERR_CANNOT_AFFORD;
    case CATAN_ACT_UPGRADE_CITY:                                  /* game.py:358-376 */
      if (s->must_respond || s->initial_phase || !s->dice_rolled || s->must_use_dev || s->just_moved_robber) return CATAN_ERR_PHASE;
      if (s->cities_left[p] > 0 && h[WHEAT] > 1 && h[ORE] > 2) {
        if (s->corner_type[t->corner] == 1) {
          if (s->corner_owner[t->corner] == pid) return 0;
        } else return CATAN_ERR_BAD_LOCATION;
      }
      return CATAN_ERR_CANNOT_AFFORD;
    case CATAN_ACT_BUY_DEV:                                       /* game.py:377-393 */
      if (s->must_respond || s->initial_phase || !s->dice_rolled || s->must_use_dev || s->just_moved_robber) return CATAN_ERR_PHASE;
      if (h[WHEAT] > 0 && h[SHEEP] > 0 && h[ORE] > 0) return s->deck_n > 0 ? 0 : CATAN_ERR_BAD_CARD;
      return CATAN_ERR_CANNOT_AFFORD;
    case CATAN_ACT_PLAY_DEV: {                                    /* game.py:394-415 */
      if (s->must_respond || s->played_dev || s->initial_phase || s->just_moved_robber) return CATAN_ERR_PHASE;
      int k = count_card(s->hidden[p], s->n_hidden[p], t->card);
      if (k > 0) {
        if (k == s->bought[t->card]) return CATAN_ERR_BAD_CARD;
        return 0;
      }
      return CATAN_ERR_BAD_CARD;
    }
    case CATAN_ACT_EXCHANGE:                                      /* game.py:416-443 */
      if (s->must_respond || s->initial_phase || !s->dice_rolled || s->must_use_dev || s->just_moved_robber) return CATAN_ERR_PHASE;
      if (h[t->res_a] >= t->rate) return s->bank[t->res_b] > 0 ? 0 : CATAN_ERR_BAD_RESOURCE;
      return CATAN_ERR_CANNOT_AFFORD;
    case CATAN_ACT_PROPOSE_TRADE: {                               /* game.py:444-466 */
      if (s->must_respond || s->initial_phase || !s->dice_rolled || s->must_use_dev || s->just_moved_robber) return CATAN_ERR_PHASE;
      int cnt[5] = {0, 0, 0, 0, 0};
      for (int k = 0; k < t->n_give; ++k) cnt[t->give[k]]++;
      for (int r = 0; r < 5; ++r) if (h[r] < cnt[r]) return CATAN_ERR_CANNOT_AFFORD;
      return 0;
    }
    case CATAN_ACT_RESPOND: {                                     /* game.py:467-482 */
      if (!s->must_respond) return CATAN_ERR_PHASE;
      if (t->accept == 1) return 0;                               /* head value 1 = reject (wrapper.py:157-160) */
      int cnt[5] = {0, 0, 0, 0, 0};
      for (int k = 0; k < s->n_recv; ++k) cnt[s->recv[k] - 1]++;
      for (int r = 0; r < 5; ++r) if (s->res[s->trade_target - 1][r] < cnt[r]) return CATAN_ERR_CANNOT_AFFORD;
      return 0;
    }
    case CATAN_ACT_MOVE_ROBBER:                                   /* game.py:483-490 */
      if (s->must_respond || s->must_use_dev) return CATAN_ERR_PHASE;
      return s->can_move_robber ? 0 : CATAN_ERR_PHASE;
    case CATAN_ACT_ROLL_DICE:                                     /* game.py:491-500 */
      if (s->must_respond || s->initial_phase || s->dice_rolled || s->just_moved_robber) return CATAN_ERR_PHASE;
      return 0;
    case CATAN_ACT_END_TURN:                                      /* game.py:501-512 */
      if (s->must_respond || s->initial_phase || !s->dice_rolled || s->must_use_dev || s->just_moved_robber) return CATAN_ERR_PHASE;
      return 0;
    case CATAN_ACT_STEAL:                                         /* game.py:513-525 */
      if (s->must_respond) return CATAN_ERR_PHASE;
      if (!s->just_moved_robber) return CATAN_ERR_PHASE;
      for (int k = 0; k < 6; ++k) {
        int c = CATAN_TILE_CORNERS[s->robber_tile][k];
        if (s->corner_type[c] && s->corner_owner[c] == t->target_pid) return 0;
      }
      return CATAN_ERR_BAD_TARGET;
  }
  return CATAN_ERR_BAD_TYPE;
}

/* ------------------------------------------------------------------------------------------
 * Game.apply_action (game.py:527-815)
 * ------------------------------------------------------------------------------------------ */
static void apply(S* s, const Act* t, Rng* g, uint8_t* info) {
  int pid = s->players_go, p = pid - 1;
  switch (t->type) {
    case CATAN_ACT_PLACE_SETTLEMENT: {                            /* game.py:530-555, 195-212 */
      int c = t->corner;
      if (!s->initial_phase) { pay(s, pid, WHEAT, 1); pay(s, pid, SHEEP, 1); pay(s, pid, WOOD, 1); pay(s, pid, BRICK, 1); }
      s->corner_type[c] = 1; s->corner_owner[c] = (int16_t)pid;
      int slot = CATAN_CORNER_HARBOUR_SLOT[c];                    /* board.py:182-183 */
      if (slot >= 0) {
        int hres = CATAN_HARBOUR_RES[s->harbour_perm[slot]];
        s->harbours[p] |= (int16_t)(hres == 0 ? 1 : (1 << hres));
      }
      s->settlements_left[p] -= 1;
      s->vp[p] += 1;
      if (s->initial_phase) {
        s->init_settlements[p] += 1;
        if (s->init_settlements[p] == 2) {
          int delta[5] = {0, 0, 0, 0, 0}, touched[5] = {0, 0, 0, 0, 0};
          for (int k = 0; k < 3; ++k) {
            int tl = CATAN_CORNER_TILES[c][k];
            if (tl < 0 || s->tile_res[tl] == 0) continue;
            int r = s->tile_res[tl] - 1;
            s->res[p][r] += 1; s->vis[p][r] += 1; delta[r] += 1; touched[r] = 1; s->bank[r] -= 1;
          }
          est_update(s, delta, touched, pid, 0);
          s->second_corner[p] = (int16_t)c;
        }
      } else {
        int delta[5] = {-1, -1, 0, -1, -1}, touched[5] = {1, 1, 0, 1, 1};
        est_update(s, delta, touched, pid, 0);
        if (s->lr_holder) update_longest_road(s, s->lr_holder);
      }
      break;
    }
    case CATAN_ACT_PLACE_ROAD: {                                  /* game.py:556-597, 222-232 */
      int final_init = 0;
      if (t->edge >= 0) {
        if (!s->initial_phase && !s->rb_active) { pay(s, pid, WOOD, 1); pay(s, pid, BRICK, 1); }
        s->edge_owner[t->edge] = (int16_t)pid;
        if (s->initial_phase) {
          s->init_roads[p] += 1;
          int first = 0, second = 0;
          for (int q = 0; q < 4; ++q) {
            if (s->init_settlements[q] == 1) first++;
            else if (s->init_settlements[q] == 2) { first++; second++; }
          }
          if (first < 4) update_players_go(s, 0);
          else if (first == 4 && second == 0) { }
          else if (first == 4 && second < 4) update_players_go(s, 1);
          else { s->initial_phase = 0; final_init = 1; }
        }
      }
      update_longest_road(s, pid);
      if (s->rb_active) {
        s->rb_count += 1;
        if (s->rb_count >= 2) { s->rb_active = 0; s->rb_count = 0; s->must_use_dev = 0; }
      } else if (!s->initial_phase && !final_init) {
        int delta[5] = {-1, -1, 0, 0, 0}, touched[5] = {1, 1, 0, 0, 0};
        est_update(s, delta, touched, pid, 0);
      }
      break;
    }
    case CATAN_ACT_UPGRADE_CITY: {                                /* game.py:598-604, 240-251 */
      pay(s, pid, WHEAT, 2); pay(s, pid, ORE, 3);
      s->corner_type[t->corner] = 2; s->corner_owner[t->corner] = (int16_t)pid;
      s->vp[p] += 1; s->cities_left[p] -= 1; s->settlements_left[p] += 1;
      int delta[5] = {0, 0, -3, 0, -2}, touched[5] = {0, 0, 1, 0, 1};
      est_update(s, delta, touched, pid, 0);
      break;
    }
    case CATAN_ACT_ROLL_DICE:                                     /* game.py:605-611 */
      roll_dice(s, g);
      s->dice_rolled = 1;
      if (s->die1 + s->die2 == 7) s->can_move_robber = 1;
      info[CATAN_INFO_ROLL] = (uint8_t)(s->die1 + s->die2);
      break;
    case CATAN_ACT_END_TURN:                                      /* game.py:612-622 */
      s->can_move_robber = 0; s->dice_rolled = 0; s->played_dev = 0;
      update_players_go(s, 0);
      s->turn += 1;
      for (int c = 0; c < 5; ++c) s->bought[c] = 0;
      s->trades_this_turn = 0; s->actions_this_turn = 0;
      break;
    case CATAN_ACT_MOVE_ROBBER: {                                 /* game.py:623-634 */
      s->robber_tile = (int16_t)t->tile;
      s->can_move_robber = 0;
      for (int k = 0; k < 6; ++k) {
        int c = CATAN_TILE_CORNERS[t->tile][k];
        if (s->corner_type[c] && s->corner_owner[c] != pid) s->just_moved_robber = 1;
      }
      break;
    }
    case CATAN_ACT_STEAL: {                                       /* game.py:635-652 */
      int v = t->target_pid;
      static const int order[5] = {BRICK, WHEAT, WOOD, SHEEP, ORE};
      int n = hand_total(s, v);
      if (n > 0) {
        int idx = rng_bounded(g, n), r = BRICK;
        for (int i = 0; i < 5; ++i) {
          int cnt = s->res[v - 1][order[i]];
          if (idx < cnt) { r = order[i]; break; }
          idx -= cnt;
        }
        s->res[p][r] += 1; s->res[v - 1][r] -= 1;
        for (int q = 0; q < 5; ++q) s->vis[v - 1][q] = (int16_t)(s->vis[v - 1][q] - 1 > 0 ? s->vis[v - 1][q] - 1 : 0);
        int delta[5] = {0, 0, 0, 0, 0}, touched[5] = {0, 0, 0, 0, 0};
        delta[r] = -1; touched[r] = 1;
        est_update(s, delta, touched, v, pid);
      }
      s->just_moved_robber = 0;
      break;
    }
    case CATAN_ACT_PLAY_DEV: {                                    /* game.py:653-693 */
      int n = s->n_hidden[p], at = -1;
      for (int i = 0; i < n; ++i) if (s->hidden[p][i] == t->card) { at = i; break; }
      if (at >= 0) {
        for (int i = at; i < n - 1; ++i) s->hidden[p][i] = s->hidden[p][i + 1];
        s->hidden[p][n - 1] = 0; s->n_hidden[p] -= 1;
      }
      s->played[p][s->n_played[p]++] = (int16_t)t->card;
      s->played_dev = 1;
      if (t->card == CATAN_DEV_VP) s->vp[p] += 1;
      else if (t->card == CATAN_DEV_KNIGHT) { s->can_move_robber = 1; update_largest_army(s); }
      else if (t->card == CATAN_DEV_ROADBUILDING) { s->rb_active = 1; s->rb_count = 0; s->must_use_dev = 1; }
      else if (t->card == CATAN_DEV_MONOPOLY) {
        int r = t->res_a, lost[4] = {0, 0, 0, 0};
        for (int o = 0; o < 4; ++o) {
          if (o == p) continue;
          int cnt = s->res[o][r];
          s->res[o][r] = 0; s->vis[o][r] = 0;
          s->res[p][r] += (int16_t)cnt; s->vis[p][r] += (int16_t)cnt;
          lost[o] = cnt;
        }
        est_monopoly(s, pid, r, lost);
      } else if (t->card == CATAN_DEV_YOP) {
        int rr[2] = {t->res_a, t->res_b};
        for (int i = 0; i < 2; ++i) {
          int r = rr[i];
          if (s->bank[r] > 0) { s->bank[r] -= 1; s->res[p][r] += 1; s->vis[p][r] += 1; est_single(s, r, 1, pid); }
        }
      }
      break;
    }
    case CATAN_ACT_BUY_DEV: {                                     /* game.py:694-710 */
      pay(s, pid, SHEEP, 1); pay(s, pid, ORE, 1); pay(s, pid, WHEAT, 1);
      int delta[5] = {0, 0, -1, -1, -1}, touched[5] = {0, 0, 1, 1, 1};
      est_update(s, delta, touched, pid, 0);
      int card = s->deck[s->deck_n - 1];
      s->deck[s->deck_n - 1] = 0; s->deck_n -= 1;
      s->hidden[p][s->n_hidden[p]++] = (int16_t)card;
      s->bought[card] += 1;
      break;
    }
    case CATAN_ACT_EXCHANGE: {                                    /* game.py:711-734 */
      int d = t->res_b, tr = t->res_a, rate = t->rate;
      s->res[p][d] += 1; s->vis[p][d] += 1;
      s->res[p][tr] -= (int16_t)rate;
      s->vis[p][tr] = (int16_t)(s->vis[p][tr] - rate > 0 ? s->vis[p][tr] - rate : 0);
      s->bank[tr] += (int16_t)rate; s->bank[d] -= 1;
      int delta[5] = {0, 0, 0, 0, 0}, touched[5] = {0, 0, 0, 0, 0};
      delta[d] = 1; touched[d] = 1;
      if (d == tr) delta[d] -= rate; else { delta[tr] = -rate; touched[tr] = 1; }
      est_update(s, delta, touched, pid, 0);
      break;
    }
    case CATAN_ACT_PROPOSE_TRADE:                                 /* game.py:735-750 */
      s->must_respond = 1;
      s->trade_proposer = (int16_t)pid; s->trade_target = (int16_t)t->target_pid;
      s->n_give = (int16_t)t->n_give; s->n_recv = (int16_t)t->n_recv;
      for (int k = 0; k < 4; ++k) {
        s->give[k] = (int16_t)(k < t->n_give ? t->give[k] + 1 : 0);
        s->recv[k] = (int16_t)(k < t->n_recv ? t->recv[k] + 1 : 0);
      }
      s->trades_this_turn += 1;
      break;
    case CATAN_ACT_RESPOND: {                                     /* game.py:751-784 */
      if (t->accept == 0) {
        int p1 = s->trade_proposer, p2 = s->trade_target;
        int d1[5] = {0, 0, 0, 0, 0}, d2[5] = {0, 0, 0, 0, 0}, touched[5] = {0, 0, 0, 0, 0};
        for (int k = 0; k < s->n_give; ++k) {
          int r = s->give[k] - 1;
          s->res[p1 - 1][r] -= 1;
          s->vis[p1 - 1][r] = (int16_t)(s->vis[p1 - 1][r] - 1 > 0 ? s->vis[p1 - 1][r] - 1 : 0);
          d1[r] -= 1;
          s->res[p2 - 1][r] += 1; s->vis[p2 - 1][r] += 1; d2[r] += 1; touched[r] = 1;
        }
        for (int k = 0; k < s->n_recv; ++k) {
          int r = s->recv[k] - 1;
          s->res[p1 - 1][r] += 1; s->vis[p1 - 1][r] += 1; d1[r] += 1;
          s->res[p2 - 1][r] -= 1;
          s->vis[p2 - 1][r] = (int16_t)(s->vis[p2 - 1][r] - 1 > 0 ? s->vis[p2 - 1][r] - 1 : 0);
          d2[r] -= 1; touched[r] = 1;
        }
        est_update(s, d1, touched, p1, 0);
        est_update(s, d2, touched, p2, 0);
      }
      s->must_respond = 0;
      s->trade_proposer = 0; s->trade_target = 0; s->n_give = 0; s->n_recv = 0;
      for (int k = 0; k < 4; ++k) { s->give[k] = 0; s->recv[k] = 0; }
      break;
    }
    case CATAN_ACT_DISCARD: {                                     /* game.py:785-807 */
      int d = s->discard_queue[0], r = t->discard;
      s->res[d - 1][r] -= 1; s->bank[r] += 1;
      est_single(s, r, -1, d);
      if (hand_total(s, d) <= 7) {
        for (int i = 0; i < 3; ++i) s->discard_queue[i] = s->discard_queue[i + 1];
        s->discard_queue[3] = 0; s->n_discard -= 1;
        if (s->n_discard == 0) s->need_discard = 0;
      }
      break;
    }
  }
  if (t->type != CATAN_ACT_RESPOND && t->type != CATAN_ACT_END_TURN && t->type != CATAN_ACT_DISCARD)
    s->actions_this_turn += 1;                                    /* game.py:809-810 */
}

int catan_oracle_actor(const S* s) {                              /* game_manager.py:152-159, wrapper.py:53-58 */
  if (s->need_discard) return s->discard_queue[0];
  if (s->must_respond) return s->trade_target;
  return s->players_go;
}

int catan_oracle_step(S* s, const catan_config_t* cfg, const int32_t* action, uint64_t seed, uint64_t env_id,
                      float* reward, uint8_t* info) {
  Act t;
  memset(info, 0, CATAN_INFO_STRIDE);
  for (int p = 0; p < 4; ++p) reward[p] = 0.0f;
  info[CATAN_INFO_ACTED] = (uint8_t)catan_oracle_actor(s);
  info[CATAN_INFO_ACT_TYPE] = (uint8_t)action[CATAN_A_TYPE];
  int err = translate(s, action, &t);
  if (!err && cfg->validate_actions) err = validate(s, &t);
  if (err) {
    info[CATAN_INFO_ERR] = (uint8_t)err;
    info[CATAN_INFO_ACTOR] = (uint8_t)catan_oracle_actor(s);
    info[CATAN_INFO_ACTOR_PRE] = info[CATAN_INFO_ACTOR];
    info[CATAN_INFO_WINNER] = (uint8_t)s->winner;
    for (int p = 0; p < 4; ++p) info[CATAN_INFO_FINAL_VP + p] = (uint8_t)s->vp[p];
    return err;
  }
  Rng g = {seed, env_id, s};
  apply(s, &t, &g, info);
  /* wrapper.py:85-112 */
  int done = 0;
  static const int dict_order[4] = {BLUE, RED, ORANGE, WHITE};    /* game.py:18-23 */
  for (int i = 0; i < 4; ++i) if (s->vp[dict_order[i] - 1] >= 10) { done = 1; s->winner = (int16_t)dict_order[i]; }
  for (int p = 0; p < 4; ++p) {
    double r = 0.0;
    if (cfg->dense_reward) {
      r += 5.0 * (double)(s->vp[p] - s->curr_vps[p]);
      if (t.type == CATAN_ACT_PLAY_DEV) r += 5.0;
      if (t.type == CATAN_ACT_MOVE_ROBBER) r += 1.0;
      if (t.type == CATAN_ACT_DISCARD) r -= 0.3;
      if (t.type == CATAN_ACT_UPGRADE_CITY) r += 2.5;
      r *= (double)cfg->reward_annealing_factor;
    }
    s->curr_vps[p] = s->vp[p];
    if (done && s->winner == p + 1) r += (double)cfg->win_reward;
    reward[p] = (float)r;
  }
  info[CATAN_INFO_DONE] = (uint8_t)done;
  info[CATAN_INFO_WINNER] = (uint8_t)s->winner;
  for (int p = 0; p < 4; ++p) info[CATAN_INFO_FINAL_VP + p] = (uint8_t)s->vp[p];
  info[CATAN_INFO_ACTOR] = (uint8_t)catan_oracle_actor(s);
  info[CATAN_INFO_ACTOR_PRE] = info[CATAN_INFO_ACTOR];
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * EnvWrapper.get_action_masks (wrapper.py:168-412)
 * ------------------------------------------------------------------------------------------ */
static int playable_cards(const S* s, int pid, uint8_t* valid /* [5] */) {   /* wrapper.py:368-388 */
  int p = pid - 1, any = 0, bank_total = 0;
  for (int r = 0; r < 5; ++r) bank_total += s->bank[r];
  for (int c = 0; c < 5; ++c) {
    valid[c] = 0;
    int k = count_card(s->hidden[p], s->n_hidden[p], c);
    if (k > 0 && s->bought[c] < k) {
      if (c == CATAN_DEV_YOP) { if (bank_total > 0) valid[c] = 1; }
      else valid[c] = 1;
    }
    any |= valid[c];
  }
  return any;
}

static void mask_play_dev(const S* s, int pid, uint8_t* m) {      /* wrapper.py:221-228 / :262-269 */
  int p = pid - 1;
  if (s->n_hidden[p] > 0 && !s->played_dev) {
    uint8_t valid[5];
    if (playable_cards(s, pid, valid)) {
      m[CATAN_MASK_TYPE + CATAN_ACT_PLAY_DEV] = 1;
      for (int c = 0; c < 5; ++c) m[CATAN_MASK_DEV + c] = valid[c];
      if (valid[CATAN_DEV_YOP]) {
        for (int r = 0; r < 5; ++r) {                             /* Q11: written to row 2 of head 9 and to head 10 */
          uint8_t b = s->bank[r] > 0;
          m[CATAN_MASK_RES_A + 2 * 5 + r] = b;
          m[CATAN_MASK_RES_B + r] = b;
        }
      }
    }
  }
}

static int mask_roads(const S* s, int pid, int road_building, uint8_t* out /* [73] */) {   /* wrapper.py:322-339 */
  int after_second = 0, second = -1, placed = 0;
  if (s->initial_phase && s->init_settlements[s->players_go - 1] == 2) {
    after_second = 1; second = s->second_corner[s->players_go - 1];
  }
  for (int e = 0; e < 72; ++e) {
    out[e] = (uint8_t)can_place_road(s, e, pid, after_second, second);
    placed |= out[e];
  }
  out[72] = (uint8_t)(!placed && road_building);
  return placed;
}

void catan_oracle_masks(const S* s, const catan_config_t* cfg, uint8_t* m) {
  memset(m, 0, CATAN_MASK_STRIDE);
  for (int i = CATAN_MASK_CORNER; i < CATAN_MASK_ENTRIES; ++i) m[i] = 1;   /* wrapper.py:172-185 */
  int pid = s->players_go, p = pid - 1;
  if (s->need_discard) {                                          /* wrapper.py:186-192 */
    int d = s->discard_queue[0];
    m[CATAN_MASK_TYPE + CATAN_ACT_DISCARD] = 1;
    for (int r = 0; r < 5; ++r) if (s->res[d - 1][r] <= 0) m[CATAN_MASK_DISCARD + r] = 0;
    return;
  }
  if (s->initial_phase) {                                         /* wrapper.py:195-204 */
    if (s->init_settlements[p] == 0 || (s->init_settlements[p] == 1 && s->init_roads[p] == 1)) {
      m[CATAN_MASK_TYPE + CATAN_ACT_PLACE_SETTLEMENT] = 1;
      for (int c = 0; c < 54; ++c) m[CATAN_MASK_CORNER + c] = (uint8_t)can_place_settlement(s, c, pid, 1);
    } else {
      m[CATAN_MASK_TYPE + CATAN_ACT_PLACE_ROAD] = 1;
      mask_roads(s, pid, 0, m + CATAN_MASK_EDGE);
    }
    return;
  }
  if (s->rb_active) {                                             /* wrapper.py:206-209 */
    m[CATAN_MASK_TYPE + CATAN_ACT_PLACE_ROAD] = 1;
    mask_roads(s, pid, 1, m + CATAN_MASK_EDGE);
    return;
  }
  if (s->just_moved_robber) {                                     /* wrapper.py:210-213, :341-351 */
    m[CATAN_MASK_TYPE + CATAN_ACT_STEAL] = 1;
    for (int l = 0; l < 3; ++l) m[CATAN_MASK_PLAYER + 3 + l] = 0;
    for (int k = 0; k < 6; ++k) {
      int c = CATAN_TILE_CORNERS[s->robber_tile][k];
      if (s->corner_type[c] && s->corner_owner[c] != pid) m[CATAN_MASK_PLAYER + 3 + label_of(s, pid, s->corner_owner[c])] = 1;
    }
    return;
  }
  if (s->must_respond) {                                          /* wrapper.py:214-218, :353-365 */
    m[CATAN_MASK_TYPE + CATAN_ACT_RESPOND] = 1;
    int cnt[5] = {0, 0, 0, 0, 0}, ok = 1;
    for (int k = 0; k < s->n_recv; ++k) cnt[s->recv[k] - 1]++;
    for (int r = 0; r < 5; ++r) if (s->res[s->trade_target - 1][r] < cnt[r]) ok = 0;
    m[CATAN_MASK_ACCEPT] = (uint8_t)ok;
    return;
  }
  if (!s->dice_rolled) {                                          /* wrapper.py:219-229 */
    m[CATAN_MASK_TYPE + CATAN_ACT_ROLL_DICE] = 1;
    mask_play_dev(s, pid, m);
    return;
  }
  m[CATAN_MASK_TYPE + CATAN_ACT_END_TURN] = 1;                    /* wrapper.py:232-234 */
  if (cfg->max_actions_per_turn >= 0 && s->actions_this_turn > cfg->max_actions_per_turn) return;
  const int16_t* h = s->res[p];
  if (h[WHEAT] > 0 && h[SHEEP] > 0 && h[WOOD] > 0 && h[BRICK] > 0) {   /* wrapper.py:238-243 */
    uint8_t v[54]; int any = 0;
    for (int c = 0; c < 54; ++c) { v[c] = (uint8_t)can_place_settlement(s, c, pid, 0); any |= v[c]; }
    if (any && s->settlements_left[p] > 0) {
      m[CATAN_MASK_TYPE + CATAN_ACT_PLACE_SETTLEMENT] = 1;
      memcpy(m + CATAN_MASK_CORNER, v, 54);
    }
  }
  if (h[WHEAT] >= 2 && h[ORE] >= 3 && s->cities_left[p] > 0) {    /* wrapper.py:245-250 */
    uint8_t v[54]; int any = 0;
    for (int c = 0; c < 54; ++c) { v[c] = (uint8_t)(s->corner_type[c] == 1 && s->corner_owner[c] == pid); any |= v[c]; }
    if (any) { m[CATAN_MASK_TYPE + CATAN_ACT_UPGRADE_CITY] = 1; memcpy(m + CATAN_MASK_CORNER + 54, v, 54); }
  }
  if (h[WOOD] > 0 && h[BRICK] > 0) {                              /* wrapper.py:252-256 */
    uint8_t v[73];
    if (mask_roads(s, pid, 0, v)) { m[CATAN_MASK_TYPE + CATAN_ACT_PLACE_ROAD] = 1; memcpy(m + CATAN_MASK_EDGE, v, 73); }
  }
  if (h[WHEAT] > 0 && h[SHEEP] > 0 && h[ORE] > 0 && s->deck_n > 0) m[CATAN_MASK_TYPE + CATAN_ACT_BUY_DEV] = 1;
  mask_play_dev(s, pid, m);                                       /* wrapper.py:262-269 */
  {                                                               /* wrapper.py:271-276, :390-412 */
    uint8_t give[5], get[5]; int ag = 0, ar = 0;
    for (int r = 0; r < 5; ++r) {
      give[r] = (uint8_t)(h[r] >= best_exchange_rate(s, pid, r));
      get[r] = (uint8_t)(s->bank[r] > 0);
      ag |= give[r]; ar |= get[r];
    }
    if (ag && ar) {
      m[CATAN_MASK_TYPE + CATAN_ACT_EXCHANGE] = 1;
      memcpy(m + CATAN_MASK_RES_A, give, 5);
      memcpy(m + CATAN_MASK_RES_B, get, 5);
    }
  }
  if (s->can_move_robber) {                                       /* wrapper.py:278-281, :308-320 (Q1) */
    m[CATAN_MASK_TYPE + CATAN_ACT_MOVE_ROBBER] = 1;
    for (int t = 0; t < 19; ++t) {
      int any = 0;
      for (int k = 0; k < 6; ++k) any |= s->corner_type[CATAN_TILE_CORNERS[t][k]] != 0;
      m[CATAN_MASK_TILE + t] = (uint8_t)any;
    }
  }
  if (hand_total(s, pid) > 0 &&                                   /* wrapper.py:283-289 */
      (cfg->max_proposed_trades_per_turn < 0 || s->trades_this_turn < cfg->max_proposed_trades_per_turn))
    m[CATAN_MASK_TYPE + CATAN_ACT_PROPOSE_TRADE] = 1;
}

/* ------------------------------------------------------------------------------------------
 * EnvWrapper._get_obs (wrapper.py:52-83), _get_tile_features (:491-524), _get_player_inputs (:526-709)
 * ------------------------------------------------------------------------------------------ */
static int bucket8(int n) { return n < 5 ? n : (n < 8 ? 5 : (n < 11 ? 6 : 7)); }          /* wrapper.py:554-561 */
static int bucket7(int n) { return n <= 2 ? n : (n <= 5 ? 3 : (n <= 7 ? 4 : (n <= 10 ? 5 : 6))); } /* :662-671 */

static void player_common(const S* s, int target, uint8_t* o) {
  /* VP one-hot 10, production 50, longest road 2, largest army 2, harbours 6  (wrapper.py:587-637) */
  int tp = target - 1;
  int vps = s->vp[tp];
  o[vps < 10 ? vps : 9] = 1;
  o += 10;
  static const int res_order[5] = {WOOD, BRICK, WHEAT, ORE, SHEEP};   /* wrapper.py:597 */
  for (int c = 0; c < 54; ++c) {
    if (!s->corner_type[c] || s->corner_owner[c] != target) continue;
    for (int k = 0; k < 3; ++k) {
      int t = CATAN_CORNER_TILES[c][k];
      if (t < 0 || s->tile_val[t] == 7) continue;
      int v = s->tile_val[t];
      int ind = v <= 6 ? v - 2 : v - 3;
      int r = s->tile_res[t] - 1, slot = 0;
      for (int i = 0; i < 5; ++i) if (res_order[i] == r) slot = i;
      o[slot * 10 + ind] += (uint8_t)s->corner_type[c];
    }
  }
  o += 50;
  if (s->lr_holder) {                                             /* wrapper.py:613-620 */
    if (s->lr_holder == target) { o[0] = 1; o[1] = (uint8_t)s->lr_count; }
    else if (s->has_path_key[tp]) o[1] = (uint8_t)s->cur_longest_path[tp];
  }
  o += 2;
  if (s->la_holder == target) o[0] = 1;                           /* wrapper.py:623-627 */
  o[1] = (uint8_t)s->cur_army[tp];
  o += 2;
  for (int b = 0; b < 6; ++b) o[b] = (uint8_t)((s->harbours[tp] >> b) & 1);   /* wrapper.py:632-637 */
}

void catan_oracle_obs(const S* s, uint8_t* o) {
  memset(o, 0, CATAN_OBS_STRIDE);
  int actor = catan_oracle_actor(s), ap = actor - 1;
  if (s->trade_proposer) {                                        /* wrapper.py:65-69 */
    for (int k = 0; k < s->n_give; ++k) o[CATAN_OBS_PROPOSED_TRADE + s->give[k]] = 1;
    for (int k = 0; k < s->n_recv; ++k) o[CATAN_OBS_PROPOSED_TRADE + s->recv[k] + 5] = 1;
  }
  for (int r = 0; r < 5; ++r) o[CATAN_OBS_CURRENT_RES + r + 1] = (uint8_t)s->res[ap][r];   /* wrapper.py:70-71 */
  for (int t = 0; t < 19; ++t) {                                  /* wrapper.py:491-524 */
    uint8_t* f = o + CATAN_OBS_TILES + t * CATAN_OBS_TILE_DIM;
    f[0] = (uint8_t)(s->robber_tile == t);
    f[1 + s->tile_val[t] - 2] = 1;
    f[12 + s->tile_res[t]] = 1;
    for (int k = 0; k < 6; ++k) {
      int c = CATAN_TILE_CORNERS[t][k];
      uint8_t* cf = f + 18 + k * 7;
      cf[s->corner_type[c]] = 1;
      if (s->corner_type[c]) {
        int ow = s->corner_owner[c];
        if (ow == actor) cf[3] = 1; else cf[3 + 1 + label_of(s, actor, ow)] = 1;
      }
    }
  }
  static const int res_order[5] = {WOOD, BRICK, WHEAT, ORE, SHEEP};   /* wrapper.py:550 */
  {                                                               /* current player, wrapper.py:698-702 */
    uint8_t* m = o + CATAN_OBS_CUR_MAIN;
    for (int i = 0; i < 5; ++i) m[i * 8 + bucket8(s->res[ap][res_order[i]])] = 1;
    player_common(s, actor, m + 40);
    uint8_t* b = m + 40 + 70;
    for (int i = 0; i < 5; ++i) b[i * 7 + bucket7(s->bank[res_order[i]])] = 1;   /* wrapper.py:657-672 */
    b[35 + bucket7(s->deck_n)] = 1;                               /* wrapper.py:674-686 */
  }
  for (int l = 0; l < 3; ++l) {                                   /* other players, wrapper.py:703-707 */
    int target = pid_at_label(s, actor, l), tp = target - 1;
    uint8_t* m = o + CATAN_OBS_OTHER_MAIN + l * CATAN_OBS_OTHER_MAIN_DIM;
    for (int i = 0; i < 5; ++i) {
      m[i * 8 + bucket8(s->est_min[ap][l][res_order[i]])] = 1;
      m[40 + i * 8 + bucket8(s->est_max[ap][l][res_order[i]])] = 1;
    }
    player_common(s, target, m + 80);
    m[150 + l] = 1;                                               /* wrapper.py:532-541 */
    int nh = s->n_hidden[tp];
    m[153 + (nh <= 4 ? nh : 5)] = 1;                              /* wrapper.py:690-695 */
  }
  /* development-card lists (wrapper.py:642-655): value card+1 */
  uint8_t* d = o + CATAN_OBS_DEV_LISTS;
  for (int i = 0; i < s->n_played[ap]; ++i) d[i] = (uint8_t)(s->played[ap][i] + 1);
  for (int i = 0; i < s->n_hidden[ap]; ++i) d[25 + i] = (uint8_t)(s->hidden[ap][i] + 1);
  o[CATAN_OBS_META + 1] = (uint8_t)s->n_played[ap];
  o[CATAN_OBS_META + 2] = (uint8_t)s->n_hidden[ap];
  for (int l = 0; l < 3; ++l) {
    int tp = pid_at_label(s, actor, l) - 1;
    for (int i = 0; i < s->n_played[tp]; ++i) d[50 + l * 25 + i] = (uint8_t)(s->played[tp][i] + 1);
    o[CATAN_OBS_META + 3 + l] = (uint8_t)s->n_played[tp];
  }
  o[CATAN_OBS_META] = (uint8_t)actor;
}

/* ------------------------------------------------------------------------------------------
 * pinned random-legal sampler (twin of oracle/ref_harness.py:sample_action)
 * ------------------------------------------------------------------------------------------ */
static int pick(const uint8_t* bits, int n, uint32_t w) {
  int k = 0;
  for (int i = 0; i < n; ++i) k += bits[i] != 0;
  if (!k) return 0;
  int j = (int)(((uint64_t)w * (uint64_t)k) >> 32);
  for (int i = 0; i < n; ++i) if (bits[i]) { if (j == 0) return i; --j; }
  return 0;
}

void catan_oracle_sample(const uint8_t* m, const uint8_t* o, uint64_t seed, uint64_t env_id, uint64_t decision,
                         int32_t* a) {
  uint32_t w[4];
  catan_oracle_philox((uint32_t)decision, CATAN_STREAM_SAMPLER, (uint32_t)env_id, (uint32_t)(env_id >> 32),
                      (uint32_t)seed, (uint32_t)(seed >> 32), w);
  for (int i = 0; i < CATAN_ACTION_WORDS; ++i) a[i] = 0;
  int t = pick(m + CATAN_MASK_TYPE, 13, w[0]);
  a[CATAN_A_TYPE] = t;
  switch (t) {
    case CATAN_ACT_PLACE_SETTLEMENT: a[CATAN_A_CORNER] = pick(m + CATAN_MASK_CORNER, 54, w[1]); break;
    case CATAN_ACT_UPGRADE_CITY: a[CATAN_A_CORNER] = pick(m + CATAN_MASK_CORNER + 54, 54, w[1]); break;
    case CATAN_ACT_PLACE_ROAD: a[CATAN_A_EDGE] = pick(m + CATAN_MASK_EDGE, 73, w[1]); break;
    case CATAN_ACT_MOVE_ROBBER: a[CATAN_A_TILE] = pick(m + CATAN_MASK_TILE, 19, w[1]); break;
    case CATAN_ACT_PLAY_DEV: {
      int card = pick(m + CATAN_MASK_DEV, 5, w[1]);
      a[CATAN_A_CARD] = card;
      if (card == CATAN_DEV_MONOPOLY) a[CATAN_A_RES_A] = pick(m + CATAN_MASK_RES_A + 10, 5, w[2]);
      else if (card == CATAN_DEV_YOP) {
        a[CATAN_A_RES_A] = pick(m + CATAN_MASK_RES_A + 15, 5, w[2]);
        a[CATAN_A_RES_B] = pick(m + CATAN_MASK_RES_B, 5, w[3]);
      }
      break;
    }
    case CATAN_ACT_EXCHANGE:
      a[CATAN_A_RES_A] = pick(m + CATAN_MASK_RES_A, 5, w[1]);
      a[CATAN_A_RES_B] = pick(m + CATAN_MASK_RES_B, 5, w[2]);
      break;
    case CATAN_ACT_PROPOSE_TRADE: {
      a[CATAN_A_PLAYER] = pick(m + CATAN_MASK_PLAYER, 3, w[1]);
      uint8_t hand[5];
      for (int r = 0; r < 5; ++r) hand[r] = o[CATAN_OBS_CURRENT_RES + 1 + r] > 0;
      a[CATAN_A_GIVE] = 1 + pick(hand, 5, w[2]);
      a[CATAN_A_RECV] = 1 + (int)(((uint64_t)w[3] * 5u) >> 32);
      break;
    }
    case CATAN_ACT_RESPOND: a[CATAN_A_ACCEPT] = pick(m + CATAN_MASK_ACCEPT, 2, w[1]); break;
    case CATAN_ACT_STEAL: a[CATAN_A_PLAYER] = pick(m + CATAN_MASK_PLAYER + 3, 3, w[1]); break;
    case CATAN_ACT_DISCARD: a[CATAN_A_DISCARD] = pick(m + CATAN_MASK_DISCARD, 5, w[1]); break;
    default: break;
  }
}

/* ------------------------------------------------------------------------------------------
 * vector driver (parity at scale + CPU baseline)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int n_envs, n_steps, fresh;
  uint64_t seed, first_env_id;
  const catan_config_t* cfg;
  S* states; uint8_t* obs; uint8_t* masks; uint64_t* decisions;
  float* reward_sum; int32_t* games_done; uint8_t* info_last; int32_t* actions_last;
  int next;                 /* next chunk start, claimed with an atomic add */
} RolloutJob;

static void rollout_one(RolloutJob* J, int e) {
  S* s = &J->states[e];
  uint8_t* o = J->obs + (size_t)e * CATAN_OBS_STRIDE;
  uint8_t* m = J->masks + (size_t)e * CATAN_MASK_STRIDE;
  uint64_t env_id = J->first_env_id + (uint64_t)e;
  if (J->fresh) {
    memset(s, 0, sizeof(*s));
    J->decisions[e] = 0;
    catan_oracle_reset(s, J->seed, env_id);
    catan_oracle_masks(s, J->cfg, m);
    catan_oracle_obs(s, o);
  }
  int32_t a[CATAN_ACTION_WORDS];
  float r[4];
  uint8_t info[CATAN_INFO_STRIDE];
  for (int t = 0; t < J->n_steps; ++t) {
    catan_oracle_sample(m, o, J->seed, env_id, J->decisions[e]++, a);
    catan_oracle_step(s, J->cfg, a, J->seed, env_id, r, info);
    if (J->reward_sum) for (int p = 0; p < 4; ++p) J->reward_sum[e * 4 + p] += r[p];
    if (info[CATAN_INFO_DONE]) {
      if (J->games_done) J->games_done[e] += 1;
      if (J->cfg->auto_reset) {
        catan_oracle_reset(s, J->seed, env_id);
        info[CATAN_INFO_RESET] = 1;
        info[CATAN_INFO_ACTOR] = (uint8_t)catan_oracle_actor(s);
      }
    }
    catan_oracle_masks(s, J->cfg, m);
    catan_oracle_obs(s, o);
  }
  if (J->info_last && J->n_steps > 0) memcpy(J->info_last + (size_t)e * CATAN_INFO_STRIDE, info, CATAN_INFO_STRIDE);
  if (J->actions_last && J->n_steps > 0) memcpy(J->actions_last + (size_t)e * CATAN_ACTION_WORDS, a, sizeof(a));
}

static void* rollout_worker(void* arg) {
  RolloutJob* J = (RolloutJob*)arg;
  const int chunk = 8;
  for (;;) {
    int b = __atomic_fetch_add(&J->next, chunk, __ATOMIC_RELAXED);
    if (b >= J->n_envs) break;
    int e1 = b + chunk < J->n_envs ? b + chunk : J->n_envs;
    for (int e = b; e < e1; ++e) rollout_one(J, e);
  }
  return NULL;
}

int catan_oracle_rollout(int n_envs, uint64_t seed, uint64_t first_env_id, int n_steps, int fresh,
                         const catan_config_t* cfg, S* states, uint8_t* obs, uint8_t* masks, uint64_t* decisions,
                         float* reward_sum, int32_t* games_done, uint8_t* info_last, int32_t* actions_last,
                         int n_threads) {
  RolloutJob J = {n_envs, n_steps, fresh, seed, first_env_id, cfg, states, obs, masks, decisions,
                  reward_sum, games_done, info_last, actions_last, 0};
  if (n_threads <= 0) { long c = sysconf(_SC_NPROCESSORS_ONLN); n_threads = c > 0 ? (int)c : 1; }
  if (n_threads > 256) n_threads = 256;
  if (n_threads > (n_envs + 7) / 8) n_threads = (n_envs + 7) / 8 > 0 ? (n_envs + 7) / 8 : 1;
  if (n_threads == 1) { rollout_worker(&J); return 1; }
  pthread_t th[256];
  int started = 0;
  for (int i = 0; i < n_threads; ++i) if (pthread_create(&th[started], NULL, rollout_worker, &J) == 0) started++;
  if (!started) { rollout_worker(&J); return 1; }
  for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
  return started;
}

/* RL/ppo/process_batch.py:134-140 — GAE reverse scan in fp32, one column at a time.
 * (Normalisation, :141-142, is checked against torch in oracle/gae_ref.py.) */
void catan_oracle_gae(const float* rewards, const float* values, const float* masks, int T, int N,
                      double gamma_d, double lam_d, float* returns, float* advantages) {
  /* torch multiplies an fp32 tensor by the Python double rounded to fp32; gamma*gae_lambda is formed in
   * double first (process_batch.py:137). */
  const float gamma = (float)gamma_d, gl = (float)(gamma_d * lam_d);
  for (int n = 0; n < N; ++n) {
    float gae = 0.0f;
    for (int t = T - 1; t >= 0; --t) {
      float delta = rewards[(size_t)t * N + n] + gamma * values[(size_t)(t + 1) * N + n] * masks[(size_t)(t + 1) * N + n]
                    - values[(size_t)t * N + n];
      gae = delta + gl * masks[(size_t)(t + 1) * N + n] * gae;
      returns[(size_t)t * N + n] = gae + values[(size_t)t * N + n];
      if (advantages) advantages[(size_t)t * N + n] = returns[(size_t)t * N + n] - values[(size_t)t * N + n];
    }
  }
}
