// TEST INFRASTRUCTURE — compiles the PRODUCT game core (settlers_of_catan_rl_b200/csrc/catan_game.cuh, the thread-per-game
// engine) for the host with one game per "chunk" so the CPU test-suite can check the exact shipped logic against the
// oracle and the golden fixtures without a GPU.  Never loaded by the product package.
#define CATAN_HOST_EMU 1
#include "../../settlers_of_catan_rl_b200/csrc/catan_game.cuh"

#include <stdlib.h>

using namespace catanb;

static const Topo h_topo = CATAN_TOPO_INITIALIZER;
static TopoX h_topox;
static bool h_topox_ready = false;

struct EmuEnv {
  GameRec g;
  alignas(16) uint8_t obs[CATAN_OBS_STRIDE];
  alignas(16) uint8_t mask[CATAN_MASK_STRIDE];
  alignas(16) uint8_t scratch[CATAN_LP_SCRATCH_BYTES + 16];
  uint32_t wbuf[CATAN_RESET_WORDS];
  alignas(4) uint8_t arr[128];
  catan_config_t cfg;
  uint64_t seed, env_id;
};

static TCx make_ctx(EmuEnv* e) {
  if (!h_topox_ready) { build_topox(h_topo, h_topox, 0, 1); h_topox_ready = true; }
  TCx cx;
  cx.g = GameView(reinterpret_cast<uint8_t*>(&e->g), 0);
  cx.T = &h_topo; cx.X = &h_topox; cx.cfg = &e->cfg; cx.seed = e->seed; cx.env_id = e->env_id;
  cx.s = load_seats(cx.g);
  return cx;
}

static void encode(EmuEnv* e) {
  TCx cx = make_ctx(e);
  MaskBits m;
  MaskPlan pl;
  Scan sc;
  memset(&sc, 0, sizeof(sc));
  if (t_masks_pre(cx, m, pl)) sc = t_scan_group(cx.g, h_topo, h_topox, cx.g.players_go(), 0, 1);
  if (pl.post) t_masks_post(cx, m, pl, sc);
  MaskFlat F;
  t_flatten_masks(m, F);
  t_store_mask_row(F, e->mask);
  // the row is produced in the same pieces as on the device (one thread per piece there)
  for (int part = 0; part < CATAN_OBS_PARTS; ++part) t_encode_obs_part(cx, e->obs, part);
}

extern "C" {

EmuEnv* emu_create(uint64_t seed, uint64_t env_id, const catan_config_t* cfg) {
  EmuEnv* e = static_cast<EmuEnv*>(aligned_alloc(64, (sizeof(EmuEnv) + 63) / 64 * 64));
  memset(e, 0, sizeof(EmuEnv));
  e->seed = seed; e->env_id = env_id; e->cfg = *cfg;
  return e;
}
void emu_destroy(EmuEnv* e) { free(e); }
void emu_set_config(EmuEnv* e, const catan_config_t* cfg) { e->cfg = *cfg; }

void emu_reset(EmuEnv* e) {
  TCx cx = make_ctx(e);
  reset_game_group(cx.g, h_topo, e->seed, e->env_id, e->wbuf, e->arr, 0, 1, nullptr);
  encode(e);
}

int emu_step(EmuEnv* e, const int32_t* action, float* reward, uint8_t* info) {
  TCx cx = make_ctx(e);
  StepTmp tmp;
  t_step_scalar(cx, action, tmp);
  if (tmp.follow) t_followups_group(cx.g, h_topo, tmp, 0, 1);
  if (!tmp.err && tmp.lr_pid) {                                      // game.py:864-919
    const int pid = tmp.lr_pid;
    int len = t_lr_fast(cx.g, h_topo, pid, tmp.lr_kind, tmp.lr_loc, tmp.acted_pid);
    if (len >= 0) {
      t_lr_apply(cx.g, pid, len, false, nullptr);
    } else if (tmp.lr_kind == CATAN_LR_ROAD && tmp.lr_loc != 0xff && !cx.g.lr_dirty(pid - 1)) {
      // the stored length is exact: only the paths through the new road can beat it
      const int old = cx.g.lr_holder() == pid ? cx.g.lr_count() : (cx.g.has_path_key(pid - 1) ? cx.g.cur_longest_path(pid - 1) : 0);
      const int through = t_through_edge(cx.g, h_topo, pid, tmp.lr_loc, e->scratch, 0);
      t_lr_apply(cx.g, pid, through > old ? through : old, false, nullptr);
    } else {
      len = t_longest_path(cx.g, h_topo, pid, e->scratch, 0);
      const bool shrunk = t_lr_is_shrunk(cx.g, pid, len);
      uint8_t other_len[5] = {0, 0, 0, 0, 0};
      if (shrunk) for (int o = WHITE; o <= RED; ++o) if (o != pid) other_len[o] = static_cast<uint8_t>(t_longest_path(cx.g, h_topo, o, e->scratch, 0));
      t_lr_apply(cx.g, pid, len, shrunk, other_len);
      cx.g.lr_dirty(pid - 1) = 0;
    }
  }
  alignas(16) float rew[4];
  alignas(16) uint8_t inf[CATAN_INFO_STRIDE];
  if (t_step_finish(cx, tmp, rew, inf)) reset_game_group(cx.g, h_topo, e->seed, e->env_id, e->wbuf, e->arr, 0, 1, inf);
  memcpy(reward, rew, sizeof(rew));
  memcpy(info, inf, sizeof(inf));
  encode(e);
  return tmp.err;
}

void emu_sample(EmuEnv* e, int32_t* action) {
  TCx cx = make_ctx(e);
  MaskBits m;
  t_load_mask_row(e->mask, m);                                       // the stand-alone sampler reads the packed row back
  uint32_t hand = 0;
  for (int r = 0; r < 5; ++r) hand |= static_cast<uint32_t>(e->obs[CATAN_OBS_CURRENT_RES + 1 + r] != 0) << r;
  alignas(16) int32_t out[CATAN_ACTION_WORDS];
  t_sample_action(m, hand, e->seed, e->env_id, e->g.decision_ctr++, out);
  memcpy(action, out, sizeof(out));
}

const uint8_t* emu_obs(EmuEnv* e) { return e->obs; }
const uint8_t* emu_masks(EmuEnv* e) { return e->mask; }
void emu_export_state(EmuEnv* e, int16_t* out) { rec_to_state(e->g, *reinterpret_cast<catan_state_t*>(out)); }
void emu_import_state(EmuEnv* e, const int16_t* in) {
  state_to_rec(*reinterpret_cast<const catan_state_t*>(in), e->g);
  encode(e);
}
int emu_randomise_uncertainty(EmuEnv* e, int controlling_pid, int max_attempts) {
  TCx cx = make_ctx(e);
  const int n = t_randomise_uncertainty(cx, controlling_pid, max_attempts);
  encode(e);
  return n;
}
int emu_longest_path(EmuEnv* e, int pid) { TCx cx = make_ctx(e); return t_longest_path(cx.g, h_topo, pid, e->scratch, 0); }
int emu_rec_bytes(void) { return static_cast<int>(sizeof(GameRec)); }

// chunk <-> record round trip through a W-wide interleaved buffer (checks chunk_get / chunk_put, which the C ABI's
// export / import use on the host)
int emu_chunk_roundtrip(EmuEnv* e, int W, int lane) {
  uint8_t* chunk = static_cast<uint8_t*>(calloc(sizeof(GameRec), static_cast<size_t>(W)));
  GameRec back;
  memset(&back, 0xee, sizeof(back));
  chunk_put(chunk, lane, W, e->g);
  chunk_get(chunk, lane, W, back);
  int bad = memcmp(&back, &e->g, sizeof(GameRec)) != 0;
  // no other lane's bytes may have been touched
  GameRec other;
  for (int l = 0; l < W && !bad; ++l) {
    if (l == lane) continue;
    chunk_get(chunk, l, W, other);
    const uint8_t* o = reinterpret_cast<const uint8_t*>(&other);
    for (size_t i = 0; i < sizeof(GameRec); ++i) bad |= o[i] != 0;
  }
  free(chunk);
  return bad;
}

}  // extern "C"
