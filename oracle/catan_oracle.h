/*
 * TEST INFRASTRUCTURE — CPU oracle for the Catan env-step hot path.
 *
 * A plain scalar C restatement of the reference's game/game.py + env/wrapper.py, written function
 * by function with the reference file:line each one follows.  It is pinned against the real
 * reference (tests/test_oracle_vs_reference.py runs the Python reference under the shared Philox
 * RNG when /root/reference is present; tests/test_oracle_golden.py replays the committed fixtures
 * under tests/golden/ that were generated from the reference by oracle/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Nothing under settlers_of_catan_rl_b200/ links, imports or calls it.
 */
#ifndef CATAN_ORACLE_H
#define CATAN_ORACLE_H

#include "../include/catan_layout.h"

#ifdef __cplusplus
extern "C" {
#endif

void catan_oracle_default_config(catan_config_t* cfg);

/* Philox4x32-10, the pinned RNG (catan_layout.h). out[4]. */
void catan_oracle_philox(uint32_t ctr0, uint32_t ctr1, uint32_t ctr2, uint32_t ctr3,
                         uint32_t key0, uint32_t key1, uint32_t out[4]);

/* Game.reset + Board.reset + EnvWrapper.reset (game.py:39-136, board.py:67-167, wrapper.py:30-34).
 * Keeps the state's game-stream draw counter. */
void catan_oracle_reset(catan_state_t* s, uint64_t seed, uint64_t env_id);

/* EnvWrapper.step without the observation (wrapper.py:36-50): translate (:114-166), validate
 * (game.py:264-525, if cfg->validate_actions), apply (game.py:527-815), done/reward (wrapper.py:85-112).
 * Returns 0 or a CATAN_ERR_* code (state untouched on error).  reward[4] by player index;
 * info[CATAN_INFO_STRIDE].  Never auto-resets. */
int catan_oracle_step(catan_state_t* s, const catan_config_t* cfg, const int32_t* action,
                      uint64_t seed, uint64_t env_id, float* reward, uint8_t* info);

/* PlayerId of the next decision maker (game_manager.py:152-159). */
int catan_oracle_actor(const catan_state_t* s);

/* EnvWrapper.get_action_masks (wrapper.py:168-412) -> uint8[CATAN_MASK_STRIDE]. */
void catan_oracle_masks(const catan_state_t* s, const catan_config_t* cfg, uint8_t* out);

/* EnvWrapper._get_obs (wrapper.py:52-83, :491-709) -> uint8[CATAN_OBS_STRIDE]. */
void catan_oracle_obs(const catan_state_t* s, uint8_t* out);

/* Game.get_longest_path (game.py:843-862) for PlayerId pid. */
int catan_oracle_longest_path(const catan_state_t* s, int pid);

/* pinned random-legal sampler: one Philox block of stream 1 per decision. */
void catan_oracle_sample(const uint8_t* mask_row, const uint8_t* obs_row, uint64_t seed, uint64_t env_id,
                         uint64_t decision, int32_t* action);

/* Vector driver used for large parity runs and as the CPU baseline: n_envs games with global ids
 * first_env_id.., each stepped n_steps times with the pinned sampler and auto-reset
 * (sample -> step -> reset-if-done -> masks -> obs).  states/obs/masks/decisions are caller
 * arrays ([n_envs] catan_state_t, [n_envs][OBS_STRIDE], [n_envs][MASK_STRIDE], [n_envs] u64) that carry
 * the situation between calls; if `fresh` they are initialised by a reset first.
 * Optional outputs (may be NULL): reward_sum[n_envs][4] accumulated rewards, games_done[n_envs]
 * accumulated finished games, info_last[n_envs][INFO_STRIDE], actions_last[n_envs][ACTION_WORDS].
 * n_threads <= 0 : all OpenMP threads.  Returns the number of threads used. */
int catan_oracle_rollout(int n_envs, uint64_t seed, uint64_t first_env_id, int n_steps, int fresh,
                         const catan_config_t* cfg, catan_state_t* states, uint8_t* obs, uint8_t* masks,
                         uint64_t* decisions, float* reward_sum, int32_t* games_done, uint8_t* info_last,
                         int32_t* actions_last, int n_threads);

/* RL/ppo/process_batch.py:134-140 restated in fp32 in torch's evaluation order: GAE reverse scan,
 * returns and un-normalised advantages, arrays [T(+1)][N].  (The normalisation, :141-142, is pinned
 * against torch itself in oracle/gae_ref.py.) */
void catan_oracle_gae(const float* rewards, const float* values, const float* masks, int T, int N,
                      double gamma, double lam, float* returns, float* advantages);

int catan_oracle_state_words(void);

#ifdef __cplusplus
}
#endif
#endif
