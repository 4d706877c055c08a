// catan_kernels.cu — sm_100a kernels + C ABI of the vectorised Catan engine (see include/catan_b200.h).
//
// env_kernel: persistent, phase-major.  One 32-warp block per SM pulls a batch of 112 games: their
// 832-byte packed records are copied (contiguous, coalesced) into shared memory and stay there while the
// block walks through the phases of a step together -- translate/validate, scalar apply, dice payout,
// belief updates, longest road, done/reward, masks(+sampler), observation -- one warp per game inside a
// phase (catan_core.cuh).  Mask and observation rows are built in a per-warp staging row and leave
// through the TMA engine as 1-D bulk async copies (cp.async.bulk.global.shared::cta).  The 1 KB board
// topology is staged in shared memory once per block.
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "../../include/catan_b200.h"
#include "catan_core.cuh"

namespace catanb {

__device__ const Topo d_topo = CATAN_TOPO_INITIALIZER;

// ---- launch shape -------------------------------------------------------------------------------
// One persistent block of 32 warps per SM.  A block works on a BATCH of kBatch games whose packed
// records stay in shared memory while the block walks through the phases of a step TOGETHER
// (block barrier between phases).  At any time all 32 warps of an SM execute the same small phase
// function, so the hot instruction footprint stays within the 32 KB instruction cache
// (profiles/r1_notes.md: the fused warp-per-game pipeline spent 74% of its stall samples in
// stall_no_inst).  Inside a phase each warp still owns one game at a time (lanes cooperate on it).
constexpr int kWarps = 32;
constexpr int kThreads = kWarps * 32;
constexpr int kBlocksPerSM = 1;              // measured: 2 x 16 warps per SM is 8% slower (profiles/r1_notes.md)
constexpr int kBatch = 96;                  // games per block iteration: 3 per warp (measured: 64 / 96 / 112 / 128 within 3% of each other)
constexpr int kStageBytes = 2304;           // per-warp staging row: >= obs row, >= longest-road scratch
constexpr int kSampleWarpsPerBlock = 4;     // stand-alone sampler kernel
constexpr int kLpWarps = 32;                // warps that run the longest-road search (the rest wait at the phase barrier)
constexpr int kLpBudget = 128;              // loop iterations per round before unfinished subtrees are re-queued
constexpr int kMaxJobs = 39;                // longest-road graphs searched per cooperative pass (13 games x 3 when re-measuring)
static_assert(kStageBytes >= CATAN_OBS_STRIDE && kStageBytes % 16 == 0, "staging row too small");

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
  const uint8_t* reset_mask;   // MODE_RESET / MODE_STEP: envs whose byte is 0 are left untouched; nullptr = all envs
  int range_first, range_count;   // env range this launch covers
  LpTask* lp_ring;                // [gridDim.x][kLpRingTasks] queue of re-split longest-road subtrees
  unsigned int* ticket;           // device-wide batch ticket counter (never reset)
  unsigned int ticket_base;       // value of *ticket when this launch starts
  unsigned long long* prof;    // profiling build only
  int lp_budget;               // loop iterations per search round (default kLpBudget; CATAN_LP_BUDGET overrides for tuning)
  int debug_flags;             // timing experiments only (CATAN_DEBUG_FLAGS): 1 = skip longest-road search, 2 = skip obs encode, 4 = skip masks
};

struct alignas(16) BlockSmem {
  Topo topo;
  GameRec recs[kBatch];
  WarpScratch ws[kBatch];
  uint8_t stage[kWarps][kStageBytes];
  int32_t n_est, n_lr, n_shrunk, pad0_;
  int32_t lp_counter, batch, pad_[2];
  int32_t lp_best[kMaxJobs];   // block-cooperative longest-road search: result per job
  int32_t lp_ctl[8];           // two sets of lp_round control words: claim counter, ring cursor, ring limit, ring base
  uint8_t est_list[kBatch], lr_list[kBatch], shrunk_list[kBatch];
  uint8_t skip[kBatch];        // games of the batch the env mask excludes: left untouched
};
static_assert(kBatch <= 4 * kWarps, "scalar phases map the games of a batch onto lanes 0..3 of the 32 warps");
// the longest-road search borrows the whole staging area: 1024 path stacks, then kMaxJobs adjacency tables
// the search borrows the staging area: path stacks of the kLpWarps searching warps | adjacency tables.  The queue of
// re-split subtrees lives in a per-block ring in global memory (L2-resident, touched only when a unit is parked or
// resumed): a shared-memory ring was too small -- once it filled up, lanes could not park and one dense network again
// pinned a lane for milliseconds (profiles/r1_notes.md).
constexpr int kLpPathBytes = 54 * kLpWarps * 32;
constexpr int kLpAdjBytes = kMaxJobs * CATAN_LP_ADJ_BYTES;
constexpr int kLpRingTasks = 2048;
static_assert(kLpPathBytes % 8 == 0 && kLpPathBytes + kLpAdjBytes <= kWarps * kStageBytes, "longest-road scratch does not fit the staging area");
static_assert(sizeof(BlockSmem) * kBlocksPerSM <= 227 * 1024, "block shared memory exceeds what kBlocksPerSM blocks can get on one SM");

// ---- TMA 1-D bulk store helpers (SASS: UBLKCP) ---------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(static_cast<uint32_t>(__cvta_generic_to_shared(ssrc))), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// stage row (built by all lanes of the warp) -> global, through the TMA engine; the row may be reused
// once bulk_wait_read_all() has returned
__device__ __forceinline__ void stage_to_global(void* gdst, const void* ssrc, uint32_t bytes, int lane) {
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) { bulk_store(gdst, ssrc, bytes); bulk_commit(); }
}
__device__ __forceinline__ void stage_reuse_wait(int lane) {
  if (lane == 0) bulk_wait_read_all();
  __syncwarp();
}

template <int MODE, bool SAMPLE>
__global__ void __launch_bounds__(kThreads, kBlocksPerSM) env_kernel(const __grid_constant__ EnvParams P) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  BlockSmem& S = *reinterpret_cast<BlockSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    const int4* src = reinterpret_cast<const int4*>(&d_topo);
    int4* dst = reinterpret_cast<int4*>(&S.topo);
    for (int i = tid; i < static_cast<int>(sizeof(Topo) / 16); i += kThreads) dst[i] = src[i];
  }
  Ctx cx;
  cx.T = &S.topo; cx.obs = S.stage[warp]; cx.mask = S.stage[warp]; cx.scratch = S.stage[warp]; cx.cfg = &P.cfg;
  cx.seed = P.seed; cx.lane = lane;
  cx.g = nullptr; cx.ws = nullptr; cx.env_id = 0;
#ifdef CATAN_PROFILE_PHASES
  cx.prof = P.prof; cx.prof_t = clock64();
#endif
  bool stage_busy = false;
  const int n_batches = (P.range_count + kBatch - 1) / kBatch;
  for (;;) {
    // blocks claim batches from a device-wide ticket counter: an SM that drew an expensive batch simply takes
    // fewer of them.  The counter is never reset: every launch makes exactly n_batches + gridDim.x claims, so the
    // host advances ticket_base by that amount (launch_env).
    __syncthreads();                                                 // previous batch fully retired; S.batch reusable
    if (tid == 0) S.batch = static_cast<int32_t>(atomicAdd(P.ticket, 1u) - P.ticket_base);
    __syncthreads();
    const int batch = S.batch;
    if (batch >= n_batches) break;
    const int base = P.range_first + batch * kBatch;                 // first env of this batch
    const int nb = min(kBatch, P.range_first + P.range_count - base);
    if (stage_busy) { stage_reuse_wait(lane); stage_busy = false; }  // the longest-road scratch aliases the staging row
    // ---- phase 0: records -> shared memory (one contiguous, fully coalesced copy), actions -> scratch
    {
      const int4* src = reinterpret_cast<const int4*>(P.recs + base);
      int4* dst = reinterpret_cast<int4*>(S.recs);
      const int n16 = nb * static_cast<int>(sizeof(GameRec) / 16);
      for (int i = tid; i < n16; i += kThreads) dst[i] = src[i];
      if (MODE == MODE_STEP) {
        const int32_t* a = P.actions + static_cast<size_t>(base) * CATAN_ACTION_WORDS;
        for (int i = tid; i < nb * CATAN_ACTION_WORDS; i += kThreads) S.ws[i / CATAN_ACTION_WORDS].action[i % CATAN_ACTION_WORDS] = a[i];
      }
      for (int i = tid; i < nb; i += kThreads) S.skip[i] = MODE != MODE_REFRESH && P.reset_mask != nullptr && P.reset_mask[base + i] == 0;
      if (tid == 0) { S.n_est = 0; S.n_lr = 0; S.n_shrunk = 0; }
    }
    __syncthreads();
    CATAN_PROF(cx, PH_LOAD);
#define CATAN_BIND(gi) do { cx.g = &S.recs[(gi)]; cx.ws = &S.ws[(gi)]; cx.env_id = P.first_env_id + static_cast<uint64_t>(base + (gi)); } while (0)
    if (MODE == MODE_STEP) {
      // scalar phases: game gi of the batch is handled by lane (gi / 32) of warp (gi % 32), so all 32 warps
      // are busy and each warp instruction serves up to four games
      const int sgi = warp + kWarps * lane;
      const bool scalar_owner = lane < 4 && sgi < nb && !S.skip[sgi];
      // ---- phase 1: translate + validate
      // ---- phase 2: scalar part of apply_action; queue the lane-parallel follow-ups (same owner thread: no barrier)
      if (scalar_owner) {
        CATAN_BIND(sgi);
        step_begin(cx);
        WarpScratch& ws = *cx.ws;
        if (ws.err) {
          P.err_flags[base + sgi] |= 1u << ws.err;
        } else {
          apply_scalar(cx);
          if (ws.dice_roll || ws.n_est || ws.est_special) S.est_list[atomicAdd(&S.n_est, 1)] = static_cast<uint8_t>(sgi);
          if (ws.lr_pid) S.lr_list[atomicAdd(&S.n_lr, 1)] = static_cast<uint8_t>(sgi);
        }
      }
      __syncthreads();
      CATAN_PROF(cx, PH_SCALAR);
      // ---- phase 3+4: dice payout (game.py:151-175) and belief updates (game.py:921-1010), warp per queued game
      for (int i = warp; i < S.n_est; i += kWarps) {
        CATAN_BIND(S.est_list[i]);
        if (cx.ws->dice_roll) dice_payout(cx);
        est_apply(cx);
      }
      __syncthreads();
      CATAN_PROF(cx, PH_EST);
      int lp_rounds_total = 0;
      // ---- phase 5: longest road (game.py:843-919), searched by the WHOLE block: the work items of every
      // queued game go into one pool that the searching warps drain in rounds, so one dense road network cannot stall the SM
      {
        uint8_t* lp_paths = &S.stage[0][0];
        uint64_t* lp_adj = reinterpret_cast<uint64_t*>(&S.stage[0][0] + kLpPathBytes);
        LpTask* lp_ring = P.lp_ring + static_cast<size_t>(blockIdx.x) * kLpRingTasks;
        const int n_lr = (P.debug_flags & 1) ? 0 : S.n_lr;
        int lp_rounds = 0;
        for (int c0 = 0; c0 < n_lr; c0 += kMaxJobs) {                // pass A: the player whose road changed
          const int nj = min(kMaxJobs, n_lr - c0);
          for (int j = warp; j < nj; j += kWarps) {
            const int gi = S.lr_list[c0 + j];
            lp_build_adj(S.recs[gi], S.topo, S.ws[gi].lr_pid, lp_adj + j * 54, lane);
          }
          if (tid < nj) S.lp_best[tid] = 0;
          __syncthreads();
#ifdef CATAN_PROFILE_PHASES
          const long long t_search0 = clock64();
#endif
          if (warp < kLpWarps) {                                     // rounds: unfinished subtrees are re-queued and re-split
            CATAN_LP_RUN(lp_adj, nj, S.lp_ctl, S.lp_best, lp_paths, kLpWarps * 32, tid, lp_ring, kLpRingTasks, P.lp_budget, tid == 0,
                         asm volatile("bar.sync 1, %0;" :: "n"(kLpWarps * 32) : "memory"), ++lp_rounds);
          }
          __syncthreads();
#ifdef CATAN_PROFILE_PHASES
          if (tid == 0) {   // pure search time per pass goes into the (otherwise unused) "dice" slot
            const unsigned long long d = static_cast<unsigned long long>(clock64() - t_search0);
            atomicAdd(&P.prof[PH_DICE * 4 + 0], d); atomicMax(&P.prof[PH_DICE * 4 + 1], d); atomicAdd(&P.prof[PH_DICE * 4 + 2], 1ull);
          }
#endif
          if (tid < nj) {
            const int gi = S.lr_list[c0 + tid];
            WarpScratch& ws = S.ws[gi];
            const int len = S.lp_best[tid];
            ws.lr_len = static_cast<uint8_t>(len);
            ws.lr_shrunk = lr_is_shrunk(S.recs[gi], ws.lr_pid, len);
            if (ws.lr_shrunk) S.shrunk_list[atomicAdd(&S.n_shrunk, 1)] = static_cast<uint8_t>(gi);
          }
          __syncthreads();
        }
        const int n_sh = S.n_shrunk;
        for (int c0 = 0; c0 < n_sh; c0 += kMaxJobs / 3) {            // pass B (rare): holder's path shrank -> the other three
          const int nj = 3 * min(kMaxJobs / 3, n_sh - c0);
          for (int j = warp; j < nj; j += kWarps) {
            const int gi = S.shrunk_list[c0 + j / 3], pid = S.ws[gi].lr_pid;
            int o = j % 3 + 1;
            if (o >= pid) ++o;
            lp_build_adj(S.recs[gi], S.topo, o, lp_adj + j * 54, lane);
          }
          if (tid < nj) S.lp_best[tid] = 0;
          __syncthreads();
#ifdef CATAN_PROFILE_PHASES
          const long long t_search0 = clock64();
#endif
          if (warp < kLpWarps) {                                     // rounds: unfinished subtrees are re-queued and re-split
            CATAN_LP_RUN(lp_adj, nj, S.lp_ctl, S.lp_best, lp_paths, kLpWarps * 32, tid, lp_ring, kLpRingTasks, P.lp_budget, tid == 0,
                         asm volatile("bar.sync 1, %0;" :: "n"(kLpWarps * 32) : "memory"), ++lp_rounds);
          }
          __syncthreads();
#ifdef CATAN_PROFILE_PHASES
          if (tid == 0) {   // pure search time per pass goes into the (otherwise unused) "dice" slot
            const unsigned long long d = static_cast<unsigned long long>(clock64() - t_search0);
            atomicAdd(&P.prof[PH_DICE * 4 + 0], d); atomicMax(&P.prof[PH_DICE * 4 + 1], d); atomicAdd(&P.prof[PH_DICE * 4 + 2], 1ull);
          }
#endif
          if (tid < nj) {
            const int gi = S.shrunk_list[c0 + tid / 3], pid = S.ws[gi].lr_pid;
            int o = tid % 3 + 1;
            if (o >= pid) ++o;
            S.ws[gi].lr_other[o] = static_cast<uint8_t>(S.lp_best[tid]);
          }
          __syncthreads();
        }
        lp_rounds_total = lp_rounds;
        if (tid < n_lr) {
          const int gi = S.lr_list[tid];
          const WarpScratch& ws = S.ws[gi];
          lr_apply(S.recs[gi], ws.lr_pid, ws.lr_len, ws.lr_shrunk != 0, ws.lr_other);
        }
      }
      __syncthreads();
#ifdef CATAN_PROFILE_PHASES
      if (tid == 0 && lp_rounds_total > 0) {   // search rounds per batch go into the (otherwise unused) "sample" slot
        atomicAdd(&P.prof[PH_SAMPLE * 4 + 0], static_cast<unsigned long long>(lp_rounds_total));
        atomicMax(&P.prof[PH_SAMPLE * 4 + 1], static_cast<unsigned long long>(lp_rounds_total));
        atomicAdd(&P.prof[PH_SAMPLE * 4 + 2], 1ull);
      }
#endif
      CATAN_PROF(cx, PH_LROAD);
      // ---- phase 6: done / reward / info (+ auto-reset).  Game gi is finished by lane gi/32 of warp gi%32, the
      // warp that also encodes its masks and observation below, so a warp barrier is enough from here on.
      if (scalar_owner) {
        CATAN_BIND(sgi);
        step_finish(cx, P.reward + static_cast<size_t>(base + sgi) * 4, P.info + static_cast<size_t>(base + sgi) * CATAN_INFO_STRIDE);
      }
      __syncwarp();
      CATAN_PROF(cx, PH_FINISH);
    } else {
      const int gi = warp + kWarps * lane;
      if (lane < 4 && gi < nb && !S.skip[gi]) {
        CATAN_BIND(gi);
        if (MODE == MODE_RESET) { reset_game(cx); cx.g->episode_steps = 0; }
        else compute_seats(cx);
        uint8_t* info = P.info + static_cast<size_t>(base + gi) * CATAN_INFO_STRIDE;
        for (int i = 0; i < CATAN_INFO_STRIDE; ++i) info[i] = 0;
        info[CATAN_INFO_ACTOR] = static_cast<uint8_t>(current_actor(*cx.g));
        info[CATAN_INFO_WINNER] = cx.g->winner;
        for (int p = 0; p < 4; ++p) info[CATAN_INFO_FINAL_VP + p] = static_cast<uint8_t>(cx.g->vp[p]);
        info[CATAN_INFO_RESET] = MODE == MODE_RESET;
      }
      __syncwarp();
    }
    // ---- phase 7: legal-action masks (+ the next random-legal action), staged row -> TMA bulk store
    for (int gi = warp; gi < nb; gi += kWarps) {
      if (S.skip[gi]) continue;
      CATAN_BIND(gi);
      if (stage_busy) stage_reuse_wait(lane);
      if (!(P.debug_flags & 4)) encode_masks(cx);
      if (SAMPLE) {
        const uint32_t decision = cx.g->decision_ctr;
        __syncwarp();
        sample_action(cx.mask, cx.g->res[current_actor(*cx.g) - 1], P.seed, cx.env_id, decision, lane,
                      P.actions_out + static_cast<size_t>(base + gi) * CATAN_ACTION_WORDS);
        if (lane == 0) cx.g->decision_ctr = decision + 1;
      }
      stage_to_global(P.masks + static_cast<size_t>(base + gi) * CATAN_MASK_STRIDE, cx.mask, CATAN_MASK_STRIDE, lane);
      stage_busy = true;
    }
    CATAN_PROF(cx, PH_MASKS);
    // ---- phase 8: packed observation, staged row -> TMA bulk store
    for (int gi = warp; gi < nb; gi += kWarps) {
      if (S.skip[gi]) continue;
      CATAN_BIND(gi);
      if (stage_busy) { stage_reuse_wait(lane); stage_busy = false; }
      if (!(P.debug_flags & 2)) encode_obs(cx);
      if (P.debug_flags & 8) {                                       // experiment: plain coalesced stores instead of the TMA engine
        const int4* src = reinterpret_cast<const int4*>(cx.obs);
        int4* dst = reinterpret_cast<int4*>(P.obs + static_cast<size_t>(base + gi) * CATAN_OBS_STRIDE);
        for (int i = lane; i < CATAN_OBS_STRIDE / 16; i += 32) dst[i] = src[i];
        __syncwarp();
      } else {
        stage_to_global(P.obs + static_cast<size_t>(base + gi) * CATAN_OBS_STRIDE, cx.obs, CATAN_OBS_STRIDE, lane);
        stage_busy = true;
      }
    }
    __syncthreads();
    CATAN_PROF(cx, PH_OBS);
    // ---- phase 9: records -> global (contiguous, coalesced)
    {
      int4* dst = reinterpret_cast<int4*>(P.recs + base);
      const int4* src = reinterpret_cast<const int4*>(S.recs);
      const int n16 = nb * static_cast<int>(sizeof(GameRec) / 16);
      for (int i = tid; i < n16; i += kThreads) dst[i] = src[i];
    }
    CATAN_PROF(cx, PH_STORE);
#undef CATAN_BIND
  }
  if (stage_busy && lane == 0) bulk_wait_all();
}

// stand-alone sampler: one warp per env, reads the bound mask/obs rows from global memory
__global__ void __launch_bounds__(kSampleWarpsPerBlock * 32) sample_kernel(GameRec* recs, int n_envs, uint64_t seed, uint64_t first_env_id,
                                                          const uint8_t* masks, const uint8_t* obs, int32_t* actions_out) {
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * kSampleWarpsPerBlock + (threadIdx.x >> 5);
  if (e >= n_envs) return;
  uint32_t dec = 0;
  if (lane == 0) { dec = recs[e].decision_ctr; recs[e].decision_ctr = dec + 1; }
  dec = __shfl_sync(0xffffffffu, dec, 0);
  sample_action(masks + static_cast<size_t>(e) * CATAN_MASK_STRIDE, obs + static_cast<size_t>(e) * CATAN_OBS_STRIDE + CATAN_OBS_CURRENT_RES + 1, seed,
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
  catanb::LpTask* lp_ring = nullptr;  // per-block task rings of the longest-road search
  unsigned int* ticket = nullptr;     // batch ticket counter of the persistent kernel
  unsigned int ticket_base = 0;
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

#ifdef CATAN_PROFILE_PHASES
static unsigned long long* g_prof_dev = nullptr;
extern "C" int catan_prof_read(unsigned long long* out_host, int clear) {   // profiling build only; not part of the ABI
  if (!g_prof_dev) return -1;
  cudaDeviceSynchronize();
  cudaMemcpy(out_host, g_prof_dev, sizeof(unsigned long long) * catanb::PH_COUNT * 4, cudaMemcpyDeviceToHost);
  if (clear) cudaMemset(g_prof_dev, 0, sizeof(unsigned long long) * catanb::PH_COUNT * 4);
  return 0;
}
#endif

static EnvParams make_params(const catan_env* env) {
  EnvParams P{};
#ifdef CATAN_PROFILE_PHASES
  if (!g_prof_dev) { cudaMalloc(&g_prof_dev, sizeof(unsigned long long) * catanb::PH_COUNT * 4); cudaMemset(g_prof_dev, 0, sizeof(unsigned long long) * catanb::PH_COUNT * 4); }
  P.prof = g_prof_dev;
#endif
  P.recs = env->recs; P.n_envs = env->n; P.seed = env->seed; P.first_env_id = env->first_env_id; P.cfg = env->cfg;
  P.obs = env->obs; P.masks = env->masks; P.reward = env->reward; P.info = env->info; P.err_flags = env->err_flags;
  { const char* d = getenv("CATAN_DEBUG_FLAGS"); P.debug_flags = d ? atoi(d) : 0; }
  { const char* d = getenv("CATAN_LP_BUDGET"); P.lp_budget = d && atoi(d) > 0 ? atoi(d) : catanb::kLpBudget; }
  return P;
}

template <int MODE, bool SAMPLE>
static int launch_env(catan_env* env, EnvParams P, int first, int count, cudaStream_t stream) {
  const size_t smem = sizeof(catanb::BlockSmem);
  // opt in to > 48 KB of dynamic shared memory (per kernel instantiation and device; cheap, so done every time)
  CATAN_CUDA(cudaFuncSetAttribute(catanb::env_kernel<MODE, SAMPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  P.range_first = first; P.range_count = count;
  int blocks = (count + catanb::kBatch - 1) / catanb::kBatch;
  if (blocks > env->grid) blocks = env->grid;                       // persistent: one 1024-thread block per SM
  if (blocks < 1) blocks = 1;
  P.ticket = env->ticket;
  P.lp_ring = env->lp_ring;
  P.ticket_base = env->ticket_base;
  env->ticket_base += static_cast<unsigned int>((count + catanb::kBatch - 1) / catanb::kBatch + blocks);   // claims this launch makes
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
  env->grid = env->sm_count * catanb::kBlocksPerSM;   // persistent blocks (shared-memory bound)
  e = cudaMalloc(&env->recs, sizeof(GameRec) * static_cast<size_t>(n_envs));
  if (e == cudaSuccess) e = cudaMemset(env->recs, 0, sizeof(GameRec) * static_cast<size_t>(n_envs));
  if (e == cudaSuccess) e = cudaMalloc(&env->err_flags, sizeof(uint32_t) * static_cast<size_t>(n_envs));
  if (e == cudaSuccess) e = cudaMemset(env->err_flags, 0, sizeof(uint32_t) * static_cast<size_t>(n_envs));
  if (e == cudaSuccess) e = cudaMalloc(&env->actions_stage, sizeof(int32_t) * CATAN_ACTION_WORDS * static_cast<size_t>(n_envs));
  if (e == cudaSuccess) e = cudaMalloc(&env->lp_ring, sizeof(catanb::LpTask) * catanb::kLpRingTasks * static_cast<size_t>(env->grid));
  if (e == cudaSuccess) e = cudaMalloc(&env->ticket, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(env->ticket, 0, sizeof(unsigned int));
  if (e != cudaSuccess) {
    cudaFree(env->recs); cudaFree(env->err_flags); cudaFree(env->actions_stage); cudaFree(env->ticket); cudaFree(env->lp_ring);
    delete env;
    return cuda_fail(e, "cudaMalloc(game records)");
  }
  *out = env;
  return 0;
}

int catan_destroy(catan_env_t* env) {
  if (!env) return 0;
  cudaFree(env->recs); cudaFree(env->err_flags); cudaFree(env->actions_stage); cudaFree(env->ticket); cudaFree(env->lp_ring);
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
  return launch_env<catanb::MODE_RESET, false>(env, P, 0, env->n, static_cast<cudaStream_t>(stream));
}

int catan_step(catan_env_t* env, const int32_t* actions_dev, void* stream) { return catan_step_masked(env, actions_dev, nullptr, stream); }

int catan_step_masked(catan_env_t* env, const int32_t* actions_dev, const uint8_t* step_mask_dev, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_dev) return fail("actions_dev is null");
  if (device_guard(env)) return -1;
  EnvParams P = make_params(env);
  P.actions = actions_dev;
  P.reset_mask = step_mask_dev;
  return launch_env<catanb::MODE_STEP, false>(env, P, 0, env->n, static_cast<cudaStream_t>(stream));
}

int catan_step_sample(catan_env_t* env, int32_t* actions_io_dev, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_io_dev) return fail("actions_io_dev is null");
  if (device_guard(env)) return -1;
  EnvParams P = make_params(env);
  P.actions = actions_io_dev;
  P.actions_out = actions_io_dev;
  return launch_env<catanb::MODE_STEP, true>(env, P, 0, env->n, static_cast<cudaStream_t>(stream));
}

int catan_sample_random(catan_env_t* env, int32_t* actions_out_dev, void* stream) {
  if (check_bound(env)) return -1;
  if (!actions_out_dev) return fail("actions_out_dev is null");
  if (device_guard(env)) return -1;
  const int blocks = (env->n + catanb::kSampleWarpsPerBlock - 1) / catanb::kSampleWarpsPerBlock;
  catanb::sample_kernel<<<blocks, catanb::kSampleWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(
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
  if (launch_env<catanb::MODE_REFRESH, false>(env, P, first, count, nullptr)) return -1;
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
