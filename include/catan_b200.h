/*
 * catan_b200.h — C ABI of the B200-native Catan self-play engine (libcatan_b200.so).
 *
 * This is the drop-in boundary for the reference's env-step + PPO-rollout hot path.  The reference
 * has no FFI of its own (it is pure Python); the boundary there is the Python class
 * env.wrapper.EnvWrapper (env/wrapper.py:11) and RL/ppo/process_batch.py's BatchProcessor.  Every
 * entry point below names the reference interface it replaces.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  `*_dev` pointers are device memory owned by
 *     the caller (e.g. tensor.data_ptr()); the library never allocates memory the caller sees.
 *   - every call returns 0 on success, < 0 on failure; catan_last_error() gives the message
 *     (thread-local).  Nothing throws across the ABI.  Illegal actions never abort a launch: they
 *     are reported per env in the info row (CATAN_INFO_ERR) and OR-ed into a sticky flag word
 *     (reference behaviour: RuntimeError from EnvWrapper.step, env/wrapper.py:38-41).
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream).  All launches are
 *     asynchronous on it; only the *_host / export / import / read_* calls synchronise.
 *   - a handle is bound to one device and is used from one host thread at a time.
 *   - layouts: catan_layout.h.
 */
#ifndef CATAN_B200_H
#define CATAN_B200_H

#include "catan_layout.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct catan_env catan_env_t;

/* library / layout introspection (used by the host side to check it agrees with the binary) */
int catan_abi_version(void);
int catan_obs_stride(void);
int catan_mask_stride(void);
int catan_info_stride(void);
int catan_action_words(void);
int catan_state_words(void);
int catan_record_bytes(void);          /* packed per-game record in HBM */
const char* catan_last_error(void);

/* reference defaults == EnvWrapper.__init__ kwargs (env/wrapper.py:12-28), auto_reset = 1 */
void catan_default_config(catan_config_t* cfg);

/* Replaces constructing n_envs EnvWrapper objects (RL/ppo/game_manager.py:16).  Games are numbered
 * first_env_id .. first_env_id + n_envs - 1; game i draws from Philox key (seed, i), so results do not
 * depend on how games are sharded over devices. */
int catan_create(int n_envs, int device, uint64_t seed, uint64_t first_env_id, const catan_config_t* cfg,
                 catan_env_t** out);
int catan_destroy(catan_env_t* env);
int catan_num_envs(const catan_env_t* env);

/* EnvWrapper attributes that callers mutate (RL/ppo/game_manager.py:164-166: reward_annealing_factor). */
int catan_set_config(catan_env_t* env, const catan_config_t* cfg);

/* Output buffers (device): obs uint8[n][CATAN_OBS_STRIDE], masks uint8[n][CATAN_MASK_STRIDE],
 * reward float[n][4] (by player index), info uint8[n][CATAN_INFO_STRIDE].  16-byte aligned; obs 32-byte aligned (its rows are
 * written with 256-bit stores). */
int catan_bind(catan_env_t* env, uint8_t* obs_dev, uint8_t* masks_dev, float* reward_dev, uint8_t* info_dev);

/* EnvWrapper.reset (env/wrapper.py:30-34) for every env (reset_mask_dev == NULL) or for envs whose
 * mask byte is non-zero; writes obs + masks (+ info actor). */
int catan_reset(catan_env_t* env, const uint8_t* reset_mask_dev, void* stream);

/* EnvWrapper.step + get_action_masks fused over all envs (env/wrapper.py:36-50, :168-290):
 * actions int32[n][CATAN_ACTION_WORDS].  Writes obs, masks, reward, info.  With cfg.auto_reset a game
 * that ends is reset inside the same launch (info RESET = 1; obs/masks describe the new game, reward /
 * DONE / WINNER / FINAL_VP describe the finished one). */
int catan_step(catan_env_t* env, const int32_t* actions_dev, void* stream);

/* catan_step for the envs whose byte in step_mask_dev[n] is non-zero; the others are frozen (state and outputs
 * untouched).  Used by the rollout collector: an env that has filled its quota of active-seat decisions waits for
 * the rest (RL/ppo/game_manager.py:78).  step_mask_dev == NULL: all envs. */
int catan_step_masked(catan_env_t* env, const int32_t* actions_dev, const uint8_t* step_mask_dev, void* stream);

/* Random-legal policy used for the env-only benchmark: samples one composite action
 * per env from the bound masks/obs into actions_out_dev. */
int catan_sample_random(catan_env_t* env, int32_t* actions_out_dev, void* stream);

/* catan_step followed by catan_sample_random in ONE launch: consumes actions_io_dev and overwrites it
 * with the next random-legal action of every env. */
int catan_step_sample(catan_env_t* env, int32_t* actions_io_dev, void* stream);

/* Same as catan_step but with HOST buffers: copies the actions host->device, steps, copies
 * obs/masks/reward/info device->host and synchronises.  Any output pointer may be NULL (not copied).
 * This is the call an EnvWrapper-shaped Python adapter makes. */
int catan_step_host(catan_env_t* env, const int32_t* actions_host, uint8_t* obs_host, uint8_t* masks_host,
                    float* reward_host, uint8_t* info_host, void* stream);
/* The same call without the final synchronisation (the copies are enqueued on `stream`; the host buffers must be pinned and
 * hold the step's result once the caller has synchronised the stream).  With the games split over two handles on two streams
 * (the double-buffered env groups of a sub-process manager, RL/ppo/vec_gather_experience.py), one half's PCIe copies overlap
 * the other half's kernels. */
int catan_step_host_async(catan_env_t* env, const int32_t* actions_host, uint8_t* obs_host, uint8_t* masks_host,
                          float* reward_host, uint8_t* info_host, void* stream);
/* catan_step_sample with host buffers, not synchronised (the env-only benchmark's random-legal policy, BASELINE.md §3, driven
 * from the host): actions_io_host (pinned) is copied in, applied, and overwritten with every env's next random-legal action. */
int catan_step_sample_host_async(catan_env_t* env, int32_t* actions_io_host, float* reward_host, uint8_t* info_host, void* stream);
int catan_reset_host(catan_env_t* env, uint8_t* obs_host, uint8_t* masks_host, uint8_t* info_host, void* stream);
/* The same call with the COMPACT host transport of action rows: one byte per word (uint8 [N, 20]; 255 stands for -1; every field of
 * an action row -- type, corner, edge, tile, card, player, resources -- is below 128).  A host-driven loop moves 80 + 112 bytes per
 * env step over PCIe with int32 rows and 20 + 52 with these; on an 8-GPU box the host's DMA bandwidth is what bounds such loops. */
int catan_step_sample_host_async_u8(catan_env_t* env, uint8_t* actions_io_host_u8, float* reward_host, uint8_t* info_host, void* stream);
/* One round of the host loop that keeps several env groups (handles) in flight -- the sub-process manager's pipelining,
 * RL/ppo/vec_gather_experience.py, without a Python round trip per group: for g = 0 .. n_groups-1, wait for group g's previous
 * step (cudaStreamSynchronize(streams[g])): its reward / info rows and next actions are then in its pinned host buffers; count the
 * `done` flags of its info rows into *done_seen (may be NULL; this is the host's read of the result); issue its next
 * catan_step_sample_host_async (action_format CATAN_ACTIONS_I32: int32 rows) or catan_step_sample_host_async_u8 (CATAN_ACTIONS_U8)
 * on streams[g].  Repeated `rounds` times.  The buffers of group g hold its last ISSUED step's result once streams[g] has been
 * synchronised by the caller. */
#define CATAN_ACTIONS_I32 0
#define CATAN_ACTIONS_U8 1
int catan_step_sample_host_groups(catan_env_t* const* envs, int n_groups, void* const* actions_io_host, int action_format,
                                  float* const* reward_host, uint8_t* const* info_host, void* const* streams, int rounds,
                                  long long* done_seen);

/* EnvWrapper.save_state / restore_state (env/wrapper.py:711-721; game/game.py:1013-1205) as the
 * canonical int16 state (catan_state_t) of `count` envs starting at `first`.  Host buffers;
 * synchronous.  import also re-encodes obs/masks of the touched envs. */
int catan_export_state(catan_env_t* env, int first, int count, int16_t* states_host);
int catan_import_state(catan_env_t* env, int first, int count, const int16_t* states_host);

/* sticky per-env error flags (bit c set = CATAN_ERR_* code c was raised since the last clear). */
int catan_read_err_flags(catan_env_t* env, uint32_t* flags_host, int clear);

/* Diagnostics of the longest-road path (game/game.py:843-919): eight counters since catan_create.
 * out_host[0] = updates triggered (road placed, or settlement placed while the card is held); [1] = updates that needed a
 * search over the road network by a whole block (the rest is settled incrementally by one thread); [2] = of those, full
 * enumerations (stored length not trusted); [3] = search tasks created; [4], [5] = GPU cycles per search, sum and max;
 * [6], [7] = walk steps per search, sum and max.  Synchronous. */
int catan_read_lr_stats(catan_env_t* env, unsigned long long* out_host);
/* Three log2 histograms of 24 bins each (bin b counts values in [2^b, 2^(b+1))), since construction: GPU cycles of a block-wide
 * search, walk steps of a search, GPU cycles of the LONGEST search of a step (what the step waits for).  out_host: 72 words.  Synchronous. */
int catan_read_lr_histograms(catan_env_t* env, unsigned long long* out_host);

/* Device-side timing of a step's two kernels on the caller's stream (CUDA events recorded by catan_step*): enable,
 * step, then read out_host[0] = steps timed, [1] = summed ms of transition_kernel, [2] = summed ms of the two encode launches
 * (observation rows, then masks + sampler; with the longest-road / reset streams running beside them, as in production),
 * [3] = summed ms of the rows launch alone.  While the hooks are on, the rows launch runs on the caller's stream in front of the
 * masks launch instead of beside it.  bench.py's roofline uses it.  Synchronous.  out_host: 9 doubles; [4..8] = summed ms, from
 * the end of the transition, until the search stream starts its search / of the search / of the rest of that stream (encode of the
 * searched games, copy-back) / until the reset stream is done / until the last of the step's streams is done. */
int catan_set_timing(catan_env_t* env, int enable);
int catan_read_timing(catan_env_t* env, double* out_host);

/* ---- PPO rollout path (RL/ppo/process_batch.py) -------------------------------------------------
 * All arrays are device fp32, time-major [T(+1)][N] like the reference's [T+1, N, 1] tensors. */

/* process_batch.py:134-140: GAE reverse scan.  rewards[T][N], values[T+1][N] (already denormalised),
 * masks[T+1][N]; writes returns[T][N] and advantages[T][N] = returns - values[:-1] (un-normalised).
 * Arithmetic is IEEE fp32 in the reference's evaluation order (no FMA contraction). */
int catan_gae(const float* rewards_dev, const float* values_dev, const float* masks_dev, int T, int N,
              double gamma, double gae_lambda, float* returns_dev, float* advantages_dev, void* stream);

/* process_batch.py:141-142: advantages <- (A - mean(A)) / (std_unbiased(A) + eps) over all elements,
 * in place.  stats_dev: double[3] = (count, sum, sum of squares).
 *   catan_adv_stats : overwrites stats_dev with the statistics of this device's `count` advantages
 *   catan_adv_apply : normalises with stats_dev.  When envs are sharded over GPUs the caller sum-all-reduces
 *                     the three doubles between the two calls to get the reference's global statistics
 *                     (SURVEY.md 8e); that 24-byte exchange is the only collective on this path. */
int catan_adv_stats(const float* advantages_dev, long long count, double* stats_dev, void* stream);
int catan_adv_apply(float* advantages_dev, long long count, const double* stats_dev, double eps, void* stream);

/* ---- rollout collector (RL/ppo/game_manager.py:69-140 + RL/ppo/process_batch.py:37-104) -----------
 * The reference records, per env, only the decisions of ONE "active" seat: an observation (+ masks) each time it
 * becomes that seat's turn, the action / log-prob it takes, the sum of its rewards until its next decision, and a
 * terminal mask that is 0 when the stored observation is the first of a new game.  Every env fills its own quota of
 * T+1 observations; time-major device buffers, all caller-owned: */
typedef struct catan_rollout {
  uint8_t* obs;               /* [T+1][N][CATAN_OBS_STRIDE]                     process_batch.py:40-51 */
  uint8_t* masks;             /* [T][N][CATAN_MASK_STRIDE]                      process_batch.py:77-91 */
  int32_t* actions;           /* [T][N][CATAN_ACTION_WORDS]                     process_batch.py:67-75 */
  float* logp;                /* [T][N]                                         process_batch.py:93-96 */
  float* rewards;             /* [T][N]                                         process_batch.py:61-65 */
  float* tmasks;              /* [T+1][N] terminal masks                        process_batch.py:98-103 */
  int32_t* cursors;           /* [N][4]  lengths of the env's obs / action / reward / terminal-mask lists */
  double* acc;                /* [N][4]  running reward sums per player, fp64 like the reference's Python floats (game_manager.py:94-95) */
  uint8_t* flags;             /* [N]     bit 0 = done_since_prev_turn (game_manager.py:77,134-136), bit 1 = last terminal mask */
  const uint8_t* active_pid;  /* [N]     the recorded seat's PlayerId (game_manager.py:26-27) */
  uint8_t* collecting;        /* [N] out (may be NULL): 1 while the env still needs observations -> catan_step_masked */
  int32_t T, N;
} catan_rollout_t;

/* One call per tick AFTER catan_step[_masked] (begin = 0): consumes the env's bound obs / masks / reward / info rows plus
 * the actions and log-probs of that tick; `stepped_dev` = the mask the step used (NULL = all).
 * begin = 1 starts a rollout without a step: fresh = 1 right after catan_reset (GamesAndPoliciesManager.reset,
 * game_manager.py:34-56), fresh = 0 carries the last observation / terminal mask over (_after_rollouts, :142-150). */
int catan_rollout_store(const catan_rollout_t* rollout, const uint8_t* env_obs_dev, const uint8_t* env_masks_dev,
                        const float* env_reward_dev, const uint8_t* env_info_dev, const int32_t* actions_dev,
                        const float* logp_dev, const uint8_t* stepped_dev, int begin, int fresh, void* stream);

/* ---- per-seat policy routing (RL/ppo/game_manager.py:21-31 policy_maps, :82-93 the per-env policy call) --------------
 * policy_map_dev uint8[N][4]: index (< n_policies) of the policy that plays PlayerId p + 1 in env n.  For every policy k:
 * lists_dev[k][0 .. counts_dev[k]) = ascending indices of the envs whose NEXT decision (info CATAN_INFO_ACTOR) is taken by
 * policy k, so that each policy runs one batched forward per tick.  active_dev uint8[N] or NULL: envs with a zero byte are
 * left out (those frozen by catan_step_masked).  counts_dev int32[n_policies], lists_dev int32[n_policies][N]. */
int catan_route_by_policy(const uint8_t* env_info_dev, const uint8_t* policy_map_dev, const uint8_t* active_dev, int N,
                          int n_policies, int32_t* counts_dev, int32_t* lists_dev, void* stream);

/* ---- policy inputs (RL/models/policy.py:168-190 obs_to_torch / act_masks_to_torch; batched form process_batch.py:43-51,
 * :80-84; consumer RL/models/observation_module.py, action_heads_module.py:62-80) ------------------------------------------
 * Expands B packed rows (the env's own obs / mask buffers, a routed subset copied out of them, or a gathered minibatch) into
 * the tensors the reference's policy network reads, in one launch:
 *   features_dev   [B][CATAN_POLICY_FEATURE_STRIDE] of dtype: columns 0..1786 = the numeric features in catan_layout.h order
 *                  (ratio features rescaled to len/8 and knights/4), columns 1787..1791 = 0;
 *   lists_dev      int64 [5][B][CATAN_OBS_DEV_PAD]: padded development-card lists (card + 1, 0 = pad), list order as in the row;
 *   head_masks_dev dtype [CATAN_MASK_ENTRIES * B]: head h starts at element CATAN_MASK_<h> * B and is [B][dim], except the
 *                  type-conditional heads 1 (corner), 6 (player), 9 (resource A), which are [types][B][dim].
 * mask_rows_dev and head_masks_dev may both be NULL (value-only pass, policy.py:108-110 get_value).  row_index_dev int32[B] or
 * NULL: batch row b is read from row row_index_dev[b] of the two row buffers — one policy's env list from
 * catan_route_by_policy, so that the per-seat policies of game_manager.py:82-93 each get their batch without a gather pass. */
#define CATAN_POLICY_FEATURE_STRIDE 1792
#define CATAN_DTYPE_F32 0
#define CATAN_DTYPE_BF16 1
int catan_policy_inputs(const uint8_t* obs_rows_dev, const uint8_t* mask_rows_dev, const int32_t* row_index_dev, int B, int dtype, void* features_dev,
                        int64_t* lists_dev, void* head_masks_dev, void* stream);

/* ---- masked categorical head (RL/distributions.py:11-40: Categorical.forward builds FixedCategorical(logits = x + log(mask));
 * the head module then calls sample() / mode(), log_probs() and entropy(), RL/models/action_heads_module.py:84-160) ----------
 * One launch for B rows of D logits (fp32, contiguous): mask_dev fp32 [B][D] (0 = illegal) or NULL.  The action is
 *   given_actions_dev[b]                      when given_actions_dev != NULL (evaluate_actions; actions_dev is not written),
 *   inverse CDF of uniforms_dev[b] in [0, 1)   when uniforms_dev != NULL (sample; never an illegal entry),
 *   the first maximum                          otherwise (mode, deterministic=True).
 * logp_dev[b] = log p(action), entropy_dev[b] (may be NULL) = -sum p log p over p > 0.  Forward only: the PPO update, which
 * differentiates through log_probs and entropy, keeps the reference's torch distribution.
 * The mask is BINARY: an entry != 0 keeps its logit (the reference adds log(mask), RL/distributions.py:36-38, which is the same
 * for the 0 / 1 masks EnvWrapper.get_action_masks produces); a given action outside [0, D) yields log-prob -inf. */
int catan_masked_categorical(const float* logits_dev, const float* mask_dev, const int64_t* given_actions_dev,
                             const float* uniforms_dev, int B, int D, int64_t* actions_dev, float* logp_dev,
                             float* entropy_dev, void* stream);

/* Game.randomise_uncertainty (game/game.py:1207-1282; consumer RL/forward_search_policy/worker.py:42-58): for every env n with
 * controlling_pid_dev[n] in 1..4, re-deal what that player cannot see (the deck + the opponents' hidden cards; the opponents' hands
 * between its minimum and maximum beliefs) until every resource adds up to 19 again; 0 leaves the env alone.  Draws come from the
 * game stream in the reference's order, so a reference game under the shared Philox stream is re-dealt identically.  The bound
 * observation / mask rows are refreshed.  An env whose beliefs admit no deal within max_attempts (the reference would loop forever)
 * gets bit CATAN_ERR_NO_DEAL in its sticky error word; its hands are then left at the minimum beliefs. */
/* CUDA-graph replay of the step calls: with enable != 0, catan_step / catan_step_masked / catan_step_sample and the two *_host_async
 * calls capture their work (6 launches on three streams, event calls, copies) ONCE per distinct set of buffer pointers and replay it
 * with one cudaGraphLaunch on the caller's stream afterwards.  Off by default; a call made while the caller's stream is itself being
 * captured (e.g. inside torch.cuda.graph) is issued directly and becomes part of that graph.  Host buffers must be pinned. */
int catan_set_graphs(catan_env_t* env, int enable);

int catan_randomise_uncertainty(catan_env_t* env, const uint8_t* controlling_pid_dev, int max_attempts, void* stream);

/* ---- the two small-dimension pieces of the policy network that the library kernels handle badly at rollout batch sizes ------
 * (the network itself stays PyTorch; these are autograd functions on its side, settlers_of_catan_rl_b200/policy_ops.py)
 * catan_tile_attention_*: the tile encoder's self-attention (RL/models/tile_encoder.py:43-57, multi_headed_attention.py:28-39):
 *   qkv fp32 [B][19][192] (q | k | v, 4 heads of 16) -> y fp32 [B][19][64] = concat_h softmax(q k^T / 4) v; the backward takes
 *   dy and returns dqkv.  16-byte aligned, contiguous.
 * catan_ln_small_*: LayerNorm over a last dimension dim <= 64 (tile_encoder.py:38, :75-76; player_modules.py:29-33) of `rows`
 *   contiguous fp32 rows; stats [rows][2] = (mean, 1 / std) is written by the forward (may be NULL) and read by the backward,
 *   which ACCUMULATES into dweight / dbias (zero them first). */
int catan_tile_attention_fwd(const float* qkv_dev, float* y_dev, int B, void* stream);
int catan_tile_attention_bwd(const float* qkv_dev, const float* dy_dev, float* dqkv_dev, int B, void* stream);
int catan_ln_small_fwd(const float* x_dev, const float* weight_dev, const float* bias_dev, float* y_dev, float* stats_dev, long long rows,
                       int dim, float eps, void* stream);
int catan_ln_small_bwd(const float* x_dev, const float* weight_dev, const float* stats_dev, const float* dy_dev, float* dx_dev,
                       float* dweight_dev, float* dbias_dev, long long rows, int dim, void* stream);

/* ---- minibatch generator (RL/ppo/process_batch.py:169-200, generator_standard) --------------------
 * The reference draws a random permutation of the T*N (time, env) pairs, cuts it into num_mini_batch index lists and, for
 * each, indexes every CPU buffer key by key and copies the pieces to the device.  Here the rollout buffers are already in
 * HBM: one launch gathers the B rows `indices_dev[b] = t * N + n` (t < T) of all arrays into contiguous minibatch buffers
 * (caller-owned, row buffers 16-byte aligned).  values [T+1][N], returns / advantages [T][N] as produced by catan_gae. */
typedef struct catan_minibatch {
  uint8_t* obs;               /* [B][CATAN_OBS_STRIDE]      obs_dict[key][:-1].view(-1, ...)[indices]   process_batch.py:179-181 */
  uint8_t* masks;             /* [B][CATAN_MASK_STRIDE]     action_masks[i]...[indices]                 :187-192 */
  int32_t* actions;           /* [B][CATAN_ACTION_WORDS]    actions[i].view(-1, ...)[indices]           :193 */
  float* logp;                /* [B]  old_action_log_probs_batch                                           :198 */
  float* values;              /* [B]  value_preds_batch = values[:-1]                                      :195 */
  float* returns;             /* [B]  returns_batch                                                        :196 */
  float* tmasks;              /* [B]  masks_batch = masks[:-1]                                             :197 */
  float* advantages;          /* [B]  adv_targets                                                          :199 */
} catan_minibatch_t;

int catan_minibatch_gather(const catan_rollout_t* rollout, const float* values_dev, const float* returns_dev,
                           const float* advantages_dev, const int32_t* indices_dev, int B, const catan_minibatch_t* out,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CATAN_B200_H */
