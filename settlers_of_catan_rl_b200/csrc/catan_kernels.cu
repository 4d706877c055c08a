// catan_kernels.cu — sm_100a kernels + C ABI of the vectorised Catan engine (see include/catan_b200.h).
//
// One env step is three launches on the caller's stream (catan_game.cuh holds the game logic):
//
//   transition_kernel   ONE THREAD PER GAME.  translate + validate + apply_action incl. dice payout and belief updates,
//                       straight on the lane-interleaved game records in HBM/L2 (every field access of a warp is one
//                       fully used 32-byte sector).  Games whose road network changed are appended to a device queue.
//   longest_road_kernel the queued longest-road re-evaluations (game.py:843-919), searched block-cooperatively: the
//                       node-simple path enumeration of all games of a batch is one pool of work units that 1024 lanes
//                       drain in rounds (lp_round), so a dense road network cannot pin a lane -- or a warp of 32 games.
//   encode_kernel       ONE THREAD PER GAME.  done / reward / info (+ auto-reset by the warp), legal-action masks as bit
//                       sets, the fused random-legal sampler, and the packed observation row streamed out through a
//                       128-byte sliding window per thread.
//
// Why not one fused kernel: the search is the only part of a step whose cost varies by four orders of magnitude between
// games; inside a thread-per-game kernel it would stall 31 other games per unit of imbalance (profiles/r1_notes.md).
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "../../include/catan_b200.h"
#include "catan_game.cuh"

namespace catanb {

__device__ const Topo d_topo = CATAN_TOPO_INITIALIZER;

// ---- launch shapes ------------------------------------------------------------------------------
constexpr int kGameThreads = 64;            // transition / encode: threads (= games) per block; small blocks keep the
                                            // 148 SMs evenly loaded at 65 536 games (1024 blocks ~ 7 per SM)
constexpr int kLrThreads = 1024;            // longest_road_kernel: one persistent block per SM
constexpr int kLrWarps = kLrThreads / 32;
constexpr int kLpBudget = 128;              // loop iterations per search round before unfinished subtrees are re-queued
constexpr int kMaxJobs = 39;                // road graphs searched per cooperative pass (13 games x 3 when re-measuring)
constexpr int kLpRingTasks = 2048;          // per-block queue of re-split subtrees (global memory, L2 resident)
constexpr int kSampleThreads = 128;         // stand-alone sampler kernel

enum { MODE_STEP = 0, MODE_RESET = 1, MODE_REFRESH = 2 };

struct LrCtl { int32_t count, ticket; };    // queue length written by transition_kernel, batch ticket of longest_road_kernel

struct EnvParams {
  uint8_t* recs;               // lane-interleaved chunks of 32 games (catan_game.cuh)
  int n_envs;
  uint64_t seed, first_env_id;
  catan_config_t cfg;
  const int32_t* actions;      // transition input
  int32_t* actions_out;        // fused sampler output (may alias `actions`), or nullptr
  uint8_t* obs;
  uint8_t* masks;
  float* reward;
  uint8_t* info;
  uint32_t* err_flags;
  const uint8_t* env_mask;     // envs whose byte is 0 are left untouched; nullptr = all envs
  int range_first, range_count;   // env range this launch covers
  uint32_t* side;              // [n] transition -> encode: err | acted_pid << 8 | act_type << 16 | roll << 24
  uint32_t* lr_queue;          // [n] env index | PlayerId << 28
  LrCtl* lr_ctl;
  LpTask* lp_ring;             // [gridDim.x][kLpRingTasks]
  int lp_budget;
};

struct GameSmem {
  Topo topo;
  TopoX topox;
};

__device__ __forceinline__ void stage_topology(GameSmem& S, int tid, int nthreads) {
  const int4* src = reinterpret_cast<const int4*>(&d_topo);
  int4* dst = reinterpret_cast<int4*>(&S.topo);
  for (int i = tid; i < static_cast<int>(sizeof(Topo) / 16); i += nthreads) dst[i] = src[i];
  build_topox(d_topo, S.topox, tid, nthreads);
  __syncthreads();
}

// ---- 1. transition ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGameThreads) transition_kernel(const __grid_constant__ EnvParams P) {
  __shared__ __align__(16) GameSmem S;
  stage_topology(S, threadIdx.x, kGameThreads);
  const int i = (P.range_first & ~31) + blockIdx.x * kGameThreads + threadIdx.x;
  if (i < P.range_first || i >= P.range_first + P.range_count) return;
  if (P.env_mask != nullptr && P.env_mask[i] == 0) return;
  TCx cx;
  cx.g = game_view(P.recs, static_cast<size_t>(i));
  cx.T = &S.topo; cx.X = &S.topox; cx.cfg = &P.cfg; cx.seed = P.seed; cx.env_id = P.first_env_id + static_cast<uint64_t>(i);
  cx.s = load_seats(cx.g);
  StepTmp tmp;
  t_step_transition(cx, P.actions + static_cast<size_t>(i) * CATAN_ACTION_WORDS, tmp);
  P.side[i] = static_cast<uint32_t>(tmp.err) | (static_cast<uint32_t>(tmp.acted_pid) << 8) | (static_cast<uint32_t>(tmp.act_type) << 16) |
              (static_cast<uint32_t>(tmp.roll_info) << 24);
  if (tmp.err) {
    P.err_flags[i] |= 1u << tmp.err;
  } else if (tmp.lr_pid) {
    const int slot = atomicAdd(&P.lr_ctl->count, 1);
    P.lr_queue[slot] = static_cast<uint32_t>(i) | (static_cast<uint32_t>(tmp.lr_pid) << 28);
  }
}

// ---- 2. longest road ----------------------------------------------------------------------------
struct alignas(16) LrSmem {
  Topo topo;
  uint64_t adj[kMaxJobs * 54];
  uint8_t paths[54 * kLrThreads];
  int32_t lp_best[kMaxJobs];
  int32_t lp_ctl[8];
  int32_t start, n_shrunk;
  uint32_t entry[kMaxJobs];
  uint8_t len[kMaxJobs], shrunk[kMaxJobs], shrunk_list[kMaxJobs];
  uint8_t other[kMaxJobs][5];
};

__global__ void __launch_bounds__(kLrThreads, 1) longest_road_kernel(const __grid_constant__ EnvParams P) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  LrSmem& S = *reinterpret_cast<LrSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int count = P.lr_ctl->count;
  if (count == 0) return;
  {
    const int4* src = reinterpret_cast<const int4*>(&d_topo);
    int4* dst = reinterpret_cast<int4*>(&S.topo);
    for (int i = tid; i < static_cast<int>(sizeof(Topo) / 16); i += kLrThreads) dst[i] = src[i];
  }
  // jobs per claim: spread the queue over the blocks, at most kMaxJobs at a time
  int per = (count + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  per = per < 1 ? 1 : (per > kMaxJobs ? kMaxJobs : per);
  LpTask* ring = P.lp_ring + static_cast<size_t>(blockIdx.x) * kLpRingTasks;
  for (;;) {
    __syncthreads();                                                 // previous batch retired; S.start reusable
    if (tid == 0) { S.start = atomicAdd(&P.lr_ctl->ticket, per); S.n_shrunk = 0; }
    __syncthreads();
    const int start = S.start;
    if (start >= count) break;
    const int nj = min(per, count - start);
    if (tid < nj) { S.entry[tid] = P.lr_queue[start + tid]; S.lp_best[tid] = 0; }
    __syncthreads();
    // pass A: the player whose road network changed
    for (int j = warp; j < nj; j += kLrWarps) {
      const uint32_t en = S.entry[j];
      t_lp_build_adj(game_view(P.recs, en & 0x0fffffffu), S.topo, static_cast<int>(en >> 28), S.adj + j * 54, lane, 32);
    }
    __syncthreads();
    CATAN_LP_RUN(S.adj, nj, S.lp_ctl, S.lp_best, S.paths, kLrThreads, tid, ring, kLpRingTasks, P.lp_budget, tid == 0, __syncthreads(), (void)0);
    __syncthreads();
    if (tid < nj) {
      const uint32_t en = S.entry[tid];
      const int len = S.lp_best[tid];
      const bool sh = t_lr_is_shrunk(game_view(P.recs, en & 0x0fffffffu), static_cast<int>(en >> 28), len);
      S.len[tid] = static_cast<uint8_t>(len);
      S.shrunk[tid] = sh;
      if (sh) S.shrunk_list[atomicAdd(&S.n_shrunk, 1)] = static_cast<uint8_t>(tid);
    }
    __syncthreads();
    // pass B (rare, game.py:880-881): the holder's path shrank -> re-measure the other three players
    const int n_sh = S.n_shrunk;
    for (int c0 = 0; c0 < n_sh; c0 += kMaxJobs / 3) {
      const int ng = min(kMaxJobs / 3, n_sh - c0), nb = 3 * ng;
      for (int j = warp; j < nb; j += kLrWarps) {
        const uint32_t en = S.entry[S.shrunk_list[c0 + j / 3]];
        const int pid = static_cast<int>(en >> 28);
        int o = j % 3 + 1;
        if (o >= pid) ++o;
        t_lp_build_adj(game_view(P.recs, en & 0x0fffffffu), S.topo, o, S.adj + j * 54, lane, 32);
      }
      if (tid < nb) S.lp_best[tid] = 0;
      __syncthreads();
      CATAN_LP_RUN(S.adj, nb, S.lp_ctl, S.lp_best, S.paths, kLrThreads, tid, ring, kLpRingTasks, P.lp_budget, tid == 0, __syncthreads(), (void)0);
      __syncthreads();
      if (tid < nb) {
        const int job = S.shrunk_list[c0 + tid / 3];
        const int pid = static_cast<int>(S.entry[job] >> 28);
        int o = tid % 3 + 1;
        if (o >= pid) ++o;
        S.other[job][o] = static_cast<uint8_t>(S.lp_best[tid]);
      }
      __syncthreads();
    }
    if (tid < nj) {
      const uint32_t en = S.entry[tid];
      t_lr_apply(game_view(P.recs, en & 0x0fffffffu), static_cast<int>(en >> 28), S.len[tid], S.shrunk[tid] != 0, S.other[tid]);
    }
  }
}

// ---- 3. finish + masks + sampler + observation --------------------------------------------------
struct alignas(16) EncSmem {
  GameSmem topo;
  uint32_t ring[(CATAN_RING_BYTES / 4) * kGameThreads];             // RowWriter windows, word-interleaved over the block
  uint32_t wbuf[kGameThreads / 32][CATAN_RESET_WORDS];              // reset: pre-drawn Philox words, per warp
  uint8_t arr[kGameThreads / 32][96];                               // reset: shuffle arrays, per warp
};

template <int MODE, bool SAMPLE>
__global__ void __launch_bounds__(kGameThreads) encode_kernel(const __grid_constant__ EnvParams P) {
  __shared__ EncSmem S;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  stage_topology(S.topo, tid, kGameThreads);
  if (blockIdx.x == 0 && tid == 0) { P.lr_ctl->count = 0; P.lr_ctl->ticket = 0; }   // the queue of this step has been consumed
  const int i = (P.range_first & ~31) + blockIdx.x * kGameThreads + tid;
  const bool valid = i >= P.range_first && i < P.range_first + P.range_count && !(MODE != MODE_REFRESH && P.env_mask != nullptr && P.env_mask[i] == 0);
  TCx cx;
  cx.g = game_view(P.recs, static_cast<size_t>(valid ? i : P.range_first));
  cx.T = &S.topo.topo; cx.X = &S.topo.topox; cx.cfg = &P.cfg; cx.seed = P.seed; cx.env_id = P.first_env_id + static_cast<uint64_t>(i);
  uint8_t* info = P.info + static_cast<size_t>(i) * CATAN_INFO_STRIDE;
  bool need_reset = false;
  if (MODE == MODE_STEP) {
    if (valid) {
      cx.s = load_seats(cx.g);
      const uint32_t sd = P.side[i];
      StepTmp tmp;
      tmp.err = static_cast<uint8_t>(sd); tmp.acted_pid = static_cast<uint8_t>(sd >> 8); tmp.act_type = static_cast<uint8_t>(sd >> 16);
      tmp.roll_info = static_cast<uint8_t>(sd >> 24);
      need_reset = t_step_finish(cx, tmp, P.reward + static_cast<size_t>(i) * 4, info);
    }
  } else if (MODE == MODE_RESET) {
    need_reset = valid;
  }
  if (MODE != MODE_REFRESH) {
    // Board.reset + Game.reset are a handful of serial shuffles: the warp does them game by game (lanes pre-draw the
    // Philox words in parallel).  Rare in a step (a game ends every ~1500 steps), everything in catan_reset.
    unsigned rb = __ballot_sync(0xffffffffu, need_reset);
    const int warp_first = i - lane;
    while (rb) {
      const int b = __ffs(static_cast<int>(rb)) - 1;
      rb &= rb - 1;
      const int e = warp_first + b;
      reset_game_group(game_view(P.recs, static_cast<size_t>(e)), S.topo.topo, P.seed, P.first_env_id + static_cast<uint64_t>(e), S.wbuf[warp], S.arr[warp],
                       lane, 32, MODE == MODE_STEP ? P.info + static_cast<size_t>(e) * CATAN_INFO_STRIDE : nullptr);
    }
  }
  if (!valid) return;
  if (MODE != MODE_STEP || need_reset) cx.s = load_seats(cx.g);
  if (MODE != MODE_STEP) t_write_info_fresh(cx.g, info, MODE == MODE_RESET);
  MaskBits m;
  t_build_masks(cx, m);
  {
    MaskFlat F;
    t_flatten_masks(m, F);
    t_store_mask_row(F, P.masks + static_cast<size_t>(i) * CATAN_MASK_STRIDE);
  }
  if (SAMPLE) {
    const int ap = t_current_actor(cx.g) - 1;
    uint32_t hand = 0;
#pragma unroll
    for (int r = 0; r < 5; ++r) hand |= static_cast<uint32_t>(cx.g.res(ap, r) != 0) << r;
    const uint32_t decision = cx.g.decision_ctr();
    cx.g.decision_ctr() = decision + 1;
    t_sample_action(m, hand, P.seed, cx.env_id, decision, P.actions_out + static_cast<size_t>(i) * CATAN_ACTION_WORDS);
  }
  t_encode_obs<kGameThreads>(cx, S.ring + tid, P.obs + static_cast<size_t>(i) * CATAN_OBS_STRIDE);
}

// stand-alone sampler: one thread per env, reads the bound mask / obs rows back from global memory
__global__ void __launch_bounds__(kSampleThreads) sample_kernel(uint8_t* recs, int n_envs, uint64_t seed, uint64_t first_env_id,
                                                                const uint8_t* masks, const uint8_t* obs, int32_t* actions_out) {
  const int e = blockIdx.x * kSampleThreads + threadIdx.x;
  if (e >= n_envs) return;
  const GameView g = game_view(recs, static_cast<size_t>(e));
  MaskBits m;
  t_load_mask_row(masks + static_cast<size_t>(e) * CATAN_MASK_STRIDE, m);
  const uint8_t* cur = obs + static_cast<size_t>(e) * CATAN_OBS_STRIDE + CATAN_OBS_CURRENT_RES + 1;
  uint32_t hand = 0;
#pragma unroll
  for (int r = 0; r < 5; ++r) hand |= static_cast<uint32_t>(cur[r] != 0) << r;
  const uint32_t dec = g.decision_ctr();
  g.decision_ctr() = dec + 1;
  t_sample_action(m, hand, seed, first_env_id + static_cast<uint64_t>(e), dec, actions_out + static_cast<size_t>(e) * CATAN_ACTION_WORDS);
}

}  // namespace catanb

// =================================================================================================
// host side: handle + C ABI
// =================================================================================================
using catanb::EnvParams;
using catanb::GameRec;

static thread_local std::string g_last_error;
static int fail(const std::string& msg) { g_last_error = msg; return -1; }
extern "C" void catan_set_last_error(const char* msg) { g_last_error = msg ? msg : ""; }   // used by ppo_kernels.cu
static int cuda_fail(cudaError_t e, const char* what) { return fail(std::string(what) + ": " + cudaGetErrorString(e)); }
#define CATAN_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

struct catan_env {
  int n = 0, device = 0, sm_count = 0;
  uint64_t seed = 0, first_env_id = 0;
  catan_config_t cfg{};
  uint8_t* recs = nullptr;            // ceil(n / 32) lane-interleaved chunks
  size_t rec_bytes = 0;
  uint32_t* err_flags = nullptr;
  uint32_t* side = nullptr;
  uint32_t* lr_queue = nullptr;
  catanb::LrCtl* lr_ctl = nullptr;
  int32_t* actions_stage = nullptr;   // device staging for catan_step_host
  catanb::LpTask* lp_ring = nullptr;  // per-block task rings of the longest-road search
  uint8_t* obs = nullptr;
  uint8_t* masks = nullptr;
  float* reward = nullptr;
  uint8_t* info = nullptr;
  int lr_grid = 0;
  int lp_budget = catanb::kLpBudget;
};

static int device_guard(const catan_env* env) {
  int cur = -1;
  CATAN_CUDA(cudaGetDevice(&cur));
  if (cur != env->device) CATAN_CUDA(cudaSetDevice(env->device));
  return 0;
}

static EnvParams make_params(const catan_env* env) {
  EnvParams P{};
  P.recs = env->recs; P.n_envs = env->n; P.seed = env->seed; P.first_env_id = env->first_env_id; P.cfg = env->cfg;
  P.obs = env->obs; P.masks = env->masks; P.reward = env->reward; P.info = env->info; P.err_flags = env->err_flags;
  P.side = env->side; P.lr_queue = env->lr_queue; P.lr_ctl = env->lr_ctl; P.lp_ring = env->lp_ring; P.lp_budget = env->lp_budget;
  return P;
}

static int game_blocks(int first, int count) {   // blocks of kGameThreads covering [first & ~31, first + count)
  const int span = first + count - (first & ~31);
  return (span + catanb::kGameThreads - 1) / catanb::kGameThreads;
}

template <int MODE, bool SAMPLE>
static int launch_encode(catan_env* env, EnvParams P, int first, int count, cudaStream_t stream) {
  P.range_first = first; P.range_count = count;
  if (count <= 0) return 0;
  catanb::encode_kernel<MODE, SAMPLE><<<game_blocks(first, count), catanb::kGameThreads, 0, stream>>>(P);
  CATAN_CUDA(cudaGetLastError());
  return 0;
}

template <bool SAMPLE>
static int launch_step(catan_env* env, EnvParams P, cudaStream_t stream) {
  P.range_first = 0; P.range_count = env->n;
  catanb::transition_kernel<<<game_blocks(0, env->n), catanb::kGameThreads, 0, stream>>>(P);
  CATAN_CUDA(cudaGetLastError());
  catanb::longest_road_kernel<<<env->lr_grid, catanb::kLrThreads, sizeof(catanb::LrSmem), stream>>>(P);
  CATAN_CUDA(cudaGetLastError());
  return launch_encode<catanb::MODE_STEP, SAMPLE>(env, P, 0, env->n, stream);
}

static int check_bound(const catan_env* env) {
  if (!env) return fail("null handle");
  if (!env->obs || !env->masks || !env->reward || !env->info) return fail("catan_bind has not been called");
  return 0;
}

static void free_env(catan_env* env) {
  cudaFree(env->recs); cudaFree(env->err_flags); cudaFree(env->side); cudaFree(env->lr_queue); cudaFree(env->lr_ctl);
  cudaFree(env->actions_stage); cudaFree(env->lp_ring);
  delete env;
}

extern "C" {

int catan_abi_version(void) { return 1; }
int catan_obs_stride(void) { return CATAN_OBS_STRIDE; }
int catan_mask_stride(void) { return CATAN_MASK_STRIDE; }
int catan_info_stride(void) { return CATAN_INFO_STRIDE; }
int catan_action_words(void) { return CATAN_ACTION_WORDS; }
int catan_state_words(void) { return CATAN_STATE_WORDS; }
int catan_record_bytes(void) { return static_cast<int>(sizeof(GameRec)); }
const char* catan_last_error(void) { return g_last_error.c_str(); }

void catan_default_config(catan_config_t* c) {
  c->max_actions_per_turn = -1;
  c->max_proposed_trades_per_turn = 4;
  c->validate_actions = 1;
  c->dense_reward = 0;
  c->auto_reset = 1;
  c->win_reward = 500.0f;
  c->reward_annealing_factor = 1.0f;
}

int catan_create(int n_envs, int device, uint64_t seed, uint64_t first_env_id, const catan_config_t* cfg, catan_env_t** out) {
  if (!out) return fail("out is null");
  *out = nullptr;
  if (n_envs <= 0) return fail("n_envs must be positive");
  if (n_envs >= (1 << 28)) return fail("n_envs must be below 2^28");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail("no CUDA device available: this library has no CPU path");
  if (device < 0 || device >= count) return fail("bad device index");
  CATAN_CUDA(cudaSetDevice(device));
  catan_env* env = new (std::nothrow) catan_env();
  if (!env) return fail("out of host memory");
  env->n = n_envs; env->device = device; env->seed = seed; env->first_env_id = first_env_id;
  if (cfg) env->cfg = *cfg; else catan_default_config(&env->cfg);
  cudaDeviceProp prop{};
  CATAN_CUDA(cudaGetDeviceProperties(&prop, device));
  env->sm_count = prop.multiProcessorCount;
  env->lr_grid = env->sm_count;                        // persistent search blocks: one per SM
  { const char* d = getenv("CATAN_LP_BUDGET"); if (d && atoi(d) > 0) env->lp_budget = atoi(d); }
  env->rec_bytes = CATAN_CHUNK_BYTES * ((static_cast<size_t>(n_envs) + 31) / 32);
  const size_t n = static_cast<size_t>(n_envs);
  e = cudaMalloc(&env->recs, env->rec_bytes);
  if (e == cudaSuccess) e = cudaMemset(env->recs, 0, env->rec_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&env->err_flags, sizeof(uint32_t) * n);
  if (e == cudaSuccess) e = cudaMemset(env->err_flags, 0, sizeof(uint32_t) * n);
  if (e == cudaSuccess) e = cudaMalloc(&env->side, sizeof(uint32_t) * n);
  if (e == cudaSuccess) e = cudaMemset(env->side, 0, sizeof(uint32_t) * n);
  if (e == cudaSuccess) e = cudaMalloc(&env->lr_queue, sizeof(uint32_t) * n);
  if (e == cudaSuccess) e = cudaMalloc(&env->lr_ctl, sizeof(catanb::LrCtl));
  if (e == cudaSuccess) e = cudaMemset(env->lr_ctl, 0, sizeof(catanb::LrCtl));
  if (e == cudaSuccess) e = cudaMalloc(&env->actions_stage, sizeof(int32_t) * CATAN_ACTION_WORDS * n);
  if (e == cudaSuccess) e = cudaMalloc(&env->lp_ring, sizeof(catanb::LpTask) * catanb::kLpRingTasks * static_cast<size_t>(env->lr_grid));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(catanb::longest_road_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(sizeof(catanb::LrSmem)));
  if (e != cudaSuccess) {
    free_env(env);
    return cuda_fail(e, "catan_create");
  }
  *out = env;
  return 0;
}

int catan_destroy(catan_env_t* env) {
  if (!env) return 0;
  free_env(env);
  return 0;
}

int catan_num_envs(const catan_env_t* env) { return env ? env->n : 0; }

int catan_set_config(catan_env_t* env, const catan_config_t* cfg) {
  if (!env || !cfg) return fail("null argument");
  env->cfg = *cfg;
  return 0;
}

int catan_bind(catan_env_t* env, uint8_t* obs_dev, uint8_t* masks_dev, float* reward_dev, uint8_t* info_dev) {
  if (!env) return fail("null handle");
  if (!obs_dev || !masks_dev || !reward_dev || !info_dev) return fail("null output buffer");
  if ((reinterpret_cast<uintptr_t>(obs_dev) | reinterpret_cast<uintptr_t>(masks_dev) | reinterpret_cast<uintptr_t>(reward_dev) |
       reinterpret_cast<uintptr_t>(info_dev)) & 15)
    return fail("output buffers must be 16-byte aligned");
  env->obs = obs_dev; env->masks = masks_dev; env->reward = reward_dev; env->info = info_dev;
  return 0;
}

int catan_reset(catan_env_t* env, const uint8_t* reset_mask_dev, void* stream) {
  if (check_bound(env)) return -1;
  if (device_guard(env)) return -1;
  EnvParams P = make_params(env);
  P.env_mask = reset_mask_dev;
  return launch_encode<catanb::MODE_RESET, false>(env, P, 0, env->n, static_cast<cudaStream_t>(stream));
}

int catan_step(catan_env_t* env, const int32_t* actions_dev, void* stream) { return catan_step_masked(env, actions_dev, nullptr, stream); }

int catan_step_masked(catan_env_t* env, const int32_t* actions_dev, const uint8_t* step_mask_dev, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_dev) return fail("actions_dev is null");
  if (device_guard(env)) return -1;
  EnvParams P = make_params(env);
  P.actions = actions_dev;
  P.env_mask = step_mask_dev;
  return launch_step<false>(env, P, static_cast<cudaStream_t>(stream));
}

int catan_step_sample(catan_env_t* env, int32_t* actions_io_dev, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_io_dev) return fail("actions_io_dev is null");
  if (device_guard(env)) return -1;
  EnvParams P = make_params(env);
  P.actions = actions_io_dev;
  P.actions_out = actions_io_dev;
  return launch_step<true>(env, P, static_cast<cudaStream_t>(stream));
}

int catan_sample_random(catan_env_t* env, int32_t* actions_out_dev, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_out_dev) return fail("actions_out_dev is null");
  if (device_guard(env)) return -1;
  const int blocks = (env->n + catanb::kSampleThreads - 1) / catanb::kSampleThreads;
  catanb::sample_kernel<<<blocks, catanb::kSampleThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      env->recs, env->n, env->seed, env->first_env_id, env->masks, env->obs, actions_out_dev);
  CATAN_CUDA(cudaGetLastError());
  return 0;
}

static int copy_outputs_to_host(catan_env_t* env, uint8_t* obs_host, uint8_t* masks_host, float* reward_host, uint8_t* info_host,
                                cudaStream_t s) {
  const size_t n = static_cast<size_t>(env->n);
  if (obs_host) CATAN_CUDA(cudaMemcpyAsync(obs_host, env->obs, n * CATAN_OBS_STRIDE, cudaMemcpyDeviceToHost, s));
  if (masks_host) CATAN_CUDA(cudaMemcpyAsync(masks_host, env->masks, n * CATAN_MASK_STRIDE, cudaMemcpyDeviceToHost, s));
  if (reward_host) CATAN_CUDA(cudaMemcpyAsync(reward_host, env->reward, n * 4 * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (info_host) CATAN_CUDA(cudaMemcpyAsync(info_host, env->info, n * CATAN_INFO_STRIDE, cudaMemcpyDeviceToHost, s));
  CATAN_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int catan_step_host(catan_env_t* env, const int32_t* actions_host, uint8_t* obs_host, uint8_t* masks_host, float* reward_host,
                    uint8_t* info_host, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_host) return fail("actions_host is null");
  if (device_guard(env)) return -1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CATAN_CUDA(cudaMemcpyAsync(env->actions_stage, actions_host, sizeof(int32_t) * CATAN_ACTION_WORDS * static_cast<size_t>(env->n),
                             cudaMemcpyHostToDevice, s));
  if (catan_step(env, env->actions_stage, stream)) return -1;
  return copy_outputs_to_host(env, obs_host, masks_host, reward_host, info_host, s);
}

int catan_reset_host(catan_env_t* env, uint8_t* obs_host, uint8_t* masks_host, uint8_t* info_host, void* stream) {
  if (catan_reset(env, nullptr, stream)) return -1;
  return copy_outputs_to_host(env, obs_host, masks_host, nullptr, info_host, static_cast<cudaStream_t>(stream));
}

// the chunks [first / 32, (first + count - 1) / 32] of the record array <-> host
static int chunk_range(const catan_env* env, int first, int count, size_t& off, size_t& bytes) {
  const size_t c0 = static_cast<size_t>(first) / 32, c1 = static_cast<size_t>(first + count - 1) / 32;
  off = c0 * CATAN_CHUNK_BYTES;
  bytes = (c1 - c0 + 1) * CATAN_CHUNK_BYTES;
  return off + bytes <= env->rec_bytes ? 0 : fail("env range out of bounds");
}

int catan_export_state(catan_env_t* env, int first, int count, int16_t* states_host) {
  if (!env || !states_host) return fail("null argument");
  if (first < 0 || count < 0 || first + count > env->n) return fail("env range out of bounds");
  if (count == 0) return 0;
  if (device_guard(env)) return -1;
  size_t off, bytes;
  if (chunk_range(env, first, count, off, bytes)) return -1;
  std::vector<uint8_t> tmp(bytes);
  CATAN_CUDA(cudaDeviceSynchronize());
  CATAN_CUDA(cudaMemcpy(tmp.data(), env->recs + off, bytes, cudaMemcpyDeviceToHost));
  catan_state_t* out = reinterpret_cast<catan_state_t*>(states_host);
  const int base = first & ~31;
  for (int i = 0; i < count; ++i) {
    const int e = first + i - base;
    GameRec rec;
    catanb::chunk_get(tmp.data() + static_cast<size_t>(e / 32) * CATAN_CHUNK_BYTES, e % 32, 32, rec);
    catanb::rec_to_state(rec, out[i]);
  }
  return 0;
}

int catan_import_state(catan_env_t* env, int first, int count, const int16_t* states_host) {
  if (check_bound(env)) return -1;
  if (!states_host) return fail("null argument");
  if (first < 0 || count < 0 || first + count > env->n) return fail("env range out of bounds");
  if (count == 0) return 0;
  if (device_guard(env)) return -1;
  size_t off, bytes;
  if (chunk_range(env, first, count, off, bytes)) return -1;
  std::vector<uint8_t> tmp(bytes);
  CATAN_CUDA(cudaDeviceSynchronize());
  CATAN_CUDA(cudaMemcpy(tmp.data(), env->recs + off, bytes, cudaMemcpyDeviceToHost));
  const catan_state_t* in = reinterpret_cast<const catan_state_t*>(states_host);
  const int base = first & ~31;
  for (int i = 0; i < count; ++i) {
    const int e = first + i - base;
    uint8_t* chunk = tmp.data() + static_cast<size_t>(e / 32) * CATAN_CHUNK_BYTES;
    GameRec rec;
    catanb::chunk_get(chunk, e % 32, 32, rec);   // keeps what the canonical state does not carry
    catanb::state_to_rec(in[i], rec);
    catanb::chunk_put(chunk, e % 32, 32, rec);
  }
  CATAN_CUDA(cudaMemcpy(env->recs + off, tmp.data(), bytes, cudaMemcpyHostToDevice));
  EnvParams P = make_params(env);
  if (launch_encode<catanb::MODE_REFRESH, false>(env, P, first, count, nullptr)) return -1;
  CATAN_CUDA(cudaDeviceSynchronize());
  return 0;
}

int catan_read_err_flags(catan_env_t* env, uint32_t* flags_host, int clear) {
  if (!env || !flags_host) return fail("null argument");
  if (device_guard(env)) return -1;
  CATAN_CUDA(cudaDeviceSynchronize());
  CATAN_CUDA(cudaMemcpy(flags_host, env->err_flags, sizeof(uint32_t) * static_cast<size_t>(env->n), cudaMemcpyDeviceToHost));
  if (clear) CATAN_CUDA(cudaMemset(env->err_flags, 0, sizeof(uint32_t) * static_cast<size_t>(env->n)));
  return 0;
}

}  // extern "C"
