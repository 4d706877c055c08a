// TEST INFRASTRUCTURE — compiles the PRODUCT game core (settlers_of_catan_rl_b200/csrc/catan_core.cuh)
// for the host with CATAN_LANES == 1 so the CPU test-suite can check the exact shipped logic against
// the oracle and the golden fixtures without a GPU.  Never loaded by the product package.
#define CATAN_HOST_EMU 1
#ifndef CATAN_LP_BUDGET
#define CATAN_LP_BUDGET 24   // tiny on purpose: the host emulation must exercise the subtree hand-off of lp_round
#endif
#include "../../settlers_of_catan_rl_b200/csrc/catan_core.cuh"

#include <stdlib.h>

using namespace catanb;

static const Topo h_topo = CATAN_TOPO_INITIALIZER;

struct EmuEnv {
  GameRec g;
  WarpScratch ws;
  uint8_t obs[CATAN_OBS_STRIDE];
  uint8_t mask[CATAN_MASK_STRIDE];
  alignas(16) uint8_t scratch[CATAN_LP_SCRATCH_BYTES + 16];
  catan_config_t cfg;
  uint64_t seed, env_id;
};

static Ctx make_ctx(EmuEnv* e) {
  Ctx cx;
  cx.g = &e->g; cx.T = &h_topo; cx.ws = &e->ws; cx.obs = e->obs; cx.mask = e->mask; cx.scratch = e->scratch; cx.cfg = &e->cfg;
  cx.seed = e->seed; cx.env_id = e->env_id; cx.lane = 0;
  return cx;
}

extern "C" {

EmuEnv* emu_create(uint64_t seed, uint64_t env_id, const catan_config_t* cfg) {
  EmuEnv* e = static_cast<EmuEnv*>(calloc(1, sizeof(EmuEnv)));
  e->seed = seed; e->env_id = env_id; e->cfg = *cfg;
  return e;
}
void emu_destroy(EmuEnv* e) { free(e); }
void emu_set_config(EmuEnv* e, const catan_config_t* cfg) { e->cfg = *cfg; }

void emu_reset(EmuEnv* e) {
  Ctx cx = make_ctx(e);
  reset_game(cx);
  encode_masks(cx);
  encode_obs(cx);
}

int emu_step(EmuEnv* e, const int32_t* action, float* reward, uint8_t* info) {
  Ctx cx = make_ctx(e);
  for (int i = 0; i < CATAN_ACTION_WORDS; ++i) e->ws.action[i] = action[i];
  int err = step_game(cx, reward, info);
  encode_masks(cx);
  encode_obs(cx);
  return err;
}

void emu_sample(EmuEnv* e, int32_t* action) {
  sample_action(e->mask, e->obs + CATAN_OBS_CURRENT_RES + 1, e->seed, e->env_id, e->g.decision_ctr++, 0, action);
}

const uint8_t* emu_obs(EmuEnv* e) { return e->obs; }
const uint8_t* emu_masks(EmuEnv* e) { return e->mask; }
void emu_export_state(EmuEnv* e, int16_t* out) { rec_to_state(e->g, *reinterpret_cast<catan_state_t*>(out)); }
void emu_import_state(EmuEnv* e, const int16_t* in) {
  Ctx cx = make_ctx(e);
  state_to_rec(*reinterpret_cast<const catan_state_t*>(in), e->g);
  compute_seats(cx);
  encode_masks(cx);
  encode_obs(cx);
}
int emu_longest_path(EmuEnv* e, int pid) { Ctx cx = make_ctx(e); compute_seats(cx); return longest_path(cx, pid); }
int emu_rec_bytes(void) { return static_cast<int>(sizeof(GameRec)); }

}  // extern "C"
