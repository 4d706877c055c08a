// catan_kernels.cu — sm_100a kernels + C ABI of the vectorised Catan engine (see include/catan_b200.h).
//
// env_kernel: one warp per game.  Per game: coalesced 16-byte loads of the 832-byte packed record into
// shared memory -> transition (catan_core.cuh) -> legal-action masks and the packed observation are
// built in shared memory -> record, masks and observation leave through the TMA engine as 1-D bulk
// async copies (cp.async.bulk.global.shared::cta) so the warp can move on to its next game while the
// stores drain.  The 1 KB board topology is staged in shared memory once per block.
#include <cuda_runtime.h>

#include <new>
#include <string>
#include <vector>

#include "../../include/catan_b200.h"
#include "catan_core.cuh"

namespace catanb {

__device__ const Topo d_topo = CATAN_TOPO_INITIALIZER;

constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * 32;

enum { MODE_STEP = 0, MODE_RESET = 1, MODE_REFRESH = 2 };

struct EnvParams {
  GameRec* recs;
  int n_envs;
  uint64_t seed, first_env_id;
  catan_config_t cfg;
  const int32_t* actions;      // MODE_STEP input
  int32_t* actions_out;        // fused sampler output (may alias `actions`), or nullptr
  uint8_t* obs;
  uint8_t* masks;
  float* reward;
  uint8_t* info;
  uint32_t* err_flags;
  const uint8_t* reset_mask;   // MODE_RESET: nullptr = all envs
  int refresh_first, refresh_count;   // MODE_REFRESH range
};

struct alignas(16) WarpSmem {
  GameRec g;
  WarpScratch ws;
  uint8_t obs[CATAN_OBS_STRIDE];
  uint8_t mask[CATAN_MASK_STRIDE];
};
static_assert(sizeof(WarpSmem) % 16 == 0, "per-warp shared slab must be 16-byte granular");
struct alignas(16) BlockSmem {
  Topo topo;
  WarpSmem w[kWarpsPerBlock];
};

// ---- TMA 1-D bulk store helpers (SASS: UBLKCP) ---------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(static_cast<uint32_t>(__cvta_generic_to_shared(ssrc))), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int MODE, bool SAMPLE>
__global__ void __launch_bounds__(kThreads) env_kernel(const __grid_constant__ EnvParams P) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BlockSmem& S = *reinterpret_cast<BlockSmem*>(smem_raw);
  {
    const int4* src = reinterpret_cast<const int4*>(&d_topo);
    int4* dst = reinterpret_cast<int4*>(&S.topo);
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(Topo) / 16); i += kThreads) dst[i] = src[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpSmem& W = S.w[warp];
  Ctx cx;
  cx.g = &W.g; cx.T = &S.topo; cx.ws = &W.ws; cx.obs = W.obs; cx.mask = W.mask; cx.cfg = &P.cfg;
  cx.seed = P.seed; cx.lane = lane;
  const int e0 = MODE == MODE_REFRESH ? P.refresh_first : 0;
  const int e1 = MODE == MODE_REFRESH ? P.refresh_first + P.refresh_count : P.n_envs;
  bool stores_in_flight = false;
  for (int e = e0 + blockIdx.x * kWarpsPerBlock + warp; e < e1; e += gridDim.x * kWarpsPerBlock) {
    if (MODE == MODE_RESET && P.reset_mask != nullptr && P.reset_mask[e] == 0) continue;
    if (stores_in_flight) {                       // the TMA engine must be done READING this warp's slab
      if (lane == 0) bulk_wait_read_all();
      __syncwarp();
    }
    cx.env_id = P.first_env_id + static_cast<uint64_t>(e);
    GameRec* grec = P.recs + e;
    {
      const int4* src = reinterpret_cast<const int4*>(grec);
      int4* dst = reinterpret_cast<int4*>(&W.g);
#pragma unroll
      for (int i = lane; i < static_cast<int>(sizeof(GameRec) / 16); i += 32) dst[i] = src[i];
    }
    if (MODE == MODE_STEP && lane < CATAN_ACTION_WORDS) W.ws.action[lane] = P.actions[static_cast<size_t>(e) * CATAN_ACTION_WORDS + lane];
    __syncwarp();
    if (MODE == MODE_STEP) {
      const int err = step_game(cx, P.reward + static_cast<size_t>(e) * 4, P.info + static_cast<size_t>(e) * CATAN_INFO_STRIDE);
      if (err && lane == 0) P.err_flags[e] |= 1u << err;
    } else {
      if (lane == 0) {
        if (MODE == MODE_RESET) { reset_game(cx); W.g.episode_steps = 0; }
        else compute_seats(cx);
        uint8_t* info = P.info + static_cast<size_t>(e) * CATAN_INFO_STRIDE;
        for (int i = 0; i < CATAN_INFO_STRIDE; ++i) info[i] = 0;
        info[CATAN_INFO_ACTOR] = static_cast<uint8_t>(current_actor(W.g));
        info[CATAN_INFO_WINNER] = W.g.winner;
        for (int p = 0; p < 4; ++p) info[CATAN_INFO_FINAL_VP + p] = static_cast<uint8_t>(W.g.vp[p]);
        info[CATAN_INFO_RESET] = MODE == MODE_RESET;
      }
      __syncwarp();
    }
    encode_masks(cx);
    encode_obs(cx);
    if (SAMPLE) {
      const uint32_t decision = W.g.decision_ctr;
      __syncwarp();
      sample_action(W.mask, W.obs, P.seed, cx.env_id, decision, lane,
                    P.actions_out + static_cast<size_t>(e) * CATAN_ACTION_WORDS);
      if (lane == 0) W.g.decision_ctr = decision + 1;
    }
    // generic-proxy writes to shared memory -> visible to the async proxy, then one lane issues the stores
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      bulk_store(grec, &W.g, sizeof(GameRec));
      bulk_store(P.masks + static_cast<size_t>(e) * CATAN_MASK_STRIDE, W.mask, CATAN_MASK_STRIDE);
      bulk_store(P.obs + static_cast<size_t>(e) * CATAN_OBS_STRIDE, W.obs, CATAN_OBS_STRIDE);
      bulk_commit();
    }
    stores_in_flight = true;
  }
  if (stores_in_flight && lane == 0) bulk_wait_all();
}

// stand-alone sampler: one warp per env, reads the bound mask/obs rows from global memory
__global__ void __launch_bounds__(kThreads) sample_kernel(GameRec* recs, int n_envs, uint64_t seed, uint64_t first_env_id,
                                                          const uint8_t* masks, const uint8_t* obs, int32_t* actions_out) {
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (e >= n_envs) return;
  uint32_t dec = 0;
  if (lane == 0) { dec = recs[e].decision_ctr; recs[e].decision_ctr = dec + 1; }
  dec = __shfl_sync(0xffffffffu, dec, 0);
  sample_action(masks + static_cast<size_t>(e) * CATAN_MASK_STRIDE, obs + static_cast<size_t>(e) * CATAN_OBS_STRIDE, seed,
                first_env_id + static_cast<uint64_t>(e), dec, lane, actions_out + static_cast<size_t>(e) * CATAN_ACTION_WORDS);
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
  GameRec* recs = nullptr;
  uint32_t* err_flags = nullptr;
  int32_t* actions_stage = nullptr;   // device staging for catan_step_host
  uint8_t* obs = nullptr;
  uint8_t* masks = nullptr;
  float* reward = nullptr;
  uint8_t* info = nullptr;
  int grid = 0;
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
  return P;
}

template <int MODE, bool SAMPLE>
static int launch_env(const catan_env* env, const EnvParams& P, int n_items, cudaStream_t stream) {
  const size_t smem = sizeof(catanb::BlockSmem);
  static_assert(sizeof(catanb::BlockSmem) <= 48 * 1024, "stays under the default dynamic shared memory limit");
  int blocks = (n_items + catanb::kWarpsPerBlock - 1) / catanb::kWarpsPerBlock;
  if (blocks > env->grid) blocks = env->grid;
  if (blocks < 1) blocks = 1;
  catanb::env_kernel<MODE, SAMPLE><<<blocks, catanb::kThreads, smem, stream>>>(P);
  CATAN_CUDA(cudaGetLastError());
  return 0;
}

static int check_bound(const catan_env* env) {
  if (!env) return fail("null handle");
  if (!env->obs || !env->masks || !env->reward || !env->info) return fail("catan_bind has not been called");
  return 0;
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
  // persistent-style grid: a whole number of waves of resident blocks (shared memory bound: ~14.4 KB/block)
  env->grid = env->sm_count * 12;
  e = cudaMalloc(&env->recs, sizeof(GameRec) * static_cast<size_t>(n_envs));
  if (e == cudaSuccess) e = cudaMemset(env->recs, 0, sizeof(GameRec) * static_cast<size_t>(n_envs));
  if (e == cudaSuccess) e = cudaMalloc(&env->err_flags, sizeof(uint32_t) * static_cast<size_t>(n_envs));
  if (e == cudaSuccess) e = cudaMemset(env->err_flags, 0, sizeof(uint32_t) * static_cast<size_t>(n_envs));
  if (e == cudaSuccess) e = cudaMalloc(&env->actions_stage, sizeof(int32_t) * CATAN_ACTION_WORDS * static_cast<size_t>(n_envs));
  if (e != cudaSuccess) {
    cudaFree(env->recs); cudaFree(env->err_flags); cudaFree(env->actions_stage);
    delete env;
    return cuda_fail(e, "cudaMalloc(game records)");
  }
  *out = env;
  return 0;
}

int catan_destroy(catan_env_t* env) {
  if (!env) return 0;
  cudaFree(env->recs); cudaFree(env->err_flags); cudaFree(env->actions_stage);
  delete env;
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
  P.reset_mask = reset_mask_dev;
  return launch_env<catanb::MODE_RESET, false>(env, P, env->n, static_cast<cudaStream_t>(stream));
}

int catan_step(catan_env_t* env, const int32_t* actions_dev, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_dev) return fail("actions_dev is null");
  if (device_guard(env)) return -1;
  EnvParams P = make_params(env);
  P.actions = actions_dev;
  return launch_env<catanb::MODE_STEP, false>(env, P, env->n, static_cast<cudaStream_t>(stream));
}

int catan_step_sample(catan_env_t* env, int32_t* actions_io_dev, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_io_dev) return fail("actions_io_dev is null");
  if (device_guard(env)) return -1;
  EnvParams P = make_params(env);
  P.actions = actions_io_dev;
  P.actions_out = actions_io_dev;
  return launch_env<catanb::MODE_STEP, true>(env, P, env->n, static_cast<cudaStream_t>(stream));
}

int catan_sample_random(catan_env_t* env, int32_t* actions_out_dev, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_out_dev) return fail("actions_out_dev is null");
  if (device_guard(env)) return -1;
  const int blocks = (env->n + catanb::kWarpsPerBlock - 1) / catanb::kWarpsPerBlock;
  catanb::sample_kernel<<<blocks, catanb::kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
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

int catan_export_state(catan_env_t* env, int first, int count, int16_t* states_host) {
  if (!env || !states_host) return fail("null argument");
  if (first < 0 || count < 0 || first + count > env->n) return fail("env range out of bounds");
  if (device_guard(env)) return -1;
  std::vector<GameRec> tmp(static_cast<size_t>(count));
  CATAN_CUDA(cudaDeviceSynchronize());
  CATAN_CUDA(cudaMemcpy(tmp.data(), env->recs + first, sizeof(GameRec) * static_cast<size_t>(count), cudaMemcpyDeviceToHost));
  catan_state_t* out = reinterpret_cast<catan_state_t*>(states_host);
  for (int i = 0; i < count; ++i) catanb::rec_to_state(tmp[static_cast<size_t>(i)], out[i]);
  return 0;
}

int catan_import_state(catan_env_t* env, int first, int count, const int16_t* states_host) {
  if (check_bound(env)) return -1;
  if (!states_host) return fail("null argument");
  if (first < 0 || count < 0 || first + count > env->n) return fail("env range out of bounds");
  if (device_guard(env)) return -1;
  std::vector<GameRec> tmp(static_cast<size_t>(count));
  CATAN_CUDA(cudaDeviceSynchronize());
  CATAN_CUDA(cudaMemcpy(tmp.data(), env->recs + first, sizeof(GameRec) * static_cast<size_t>(count), cudaMemcpyDeviceToHost));
  const catan_state_t* in = reinterpret_cast<const catan_state_t*>(states_host);
  for (int i = 0; i < count; ++i) catanb::state_to_rec(in[i], tmp[static_cast<size_t>(i)]);
  CATAN_CUDA(cudaMemcpy(env->recs + first, tmp.data(), sizeof(GameRec) * static_cast<size_t>(count), cudaMemcpyHostToDevice));
  EnvParams P = make_params(env);
  P.refresh_first = first; P.refresh_count = count;
  if (launch_env<catanb::MODE_REFRESH, false>(env, P, count, nullptr)) return -1;
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
