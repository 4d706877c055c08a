// ppo_kernels.cu — rollout-side PPO kernels (RL/ppo/process_batch.py:134-142): GAE reverse scan and
// advantage normalisation.  HBM-bound streaming kernels; arithmetic is plain IEEE fp32 with explicit
// round-to-nearest intrinsics so that no FMA contraction changes the reference's results.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "device_scope.cuh"

#include <string>

#include "../../include/catan_b200.h"

namespace catanb {

// One thread per env column, time-major arrays => every load/store of a warp is one 128-byte line.
// The only loop-carried value is `gae`; loads of older rows are independent of it and are hoisted by
// the unroll, which keeps ~8 lines per thread in flight.
__global__ void __launch_bounds__(256) gae_kernel(const float* __restrict__ rewards, const float* __restrict__ values,
                                                  const float* __restrict__ masks, int T, int N, float gamma, float gamma_lambda,
                                                  float* __restrict__ returns, float* __restrict__ advantages) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float gae = 0.0f;
  float v_next = values[static_cast<size_t>(T) * N + n];
#pragma unroll 8
  for (int t = T - 1; t >= 0; --t) {
    const size_t i = static_cast<size_t>(t) * N + n;
    const float m = masks[i + N], v = values[i], r = rewards[i];
    // delta = r + gamma * V[t+1] * m[t+1] - V[t]                       (process_batch.py:136)
    const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(gamma, v_next), m)), v);
    // gae = delta + (gamma * lambda) * m[t+1] * gae                    (process_batch.py:137)
    gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_lambda, m), gae));
    const float ret = __fadd_rn(gae, v);                              // process_batch.py:138
    returns[i] = ret;
    advantages[i] = __fsub_rn(ret, v);                                // process_batch.py:140
    v_next = v;
  }
}

__global__ void __launch_bounds__(256) adv_stats_kernel(const float* __restrict__ a, long long count, double* __restrict__ stats) {
  double s = 0.0, ss = 0.0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n4 = count >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(a);
  for (long long i = tid; i < n4; i += stride) {
    const float4 v = a4[i];
    s += static_cast<double>(v.x) + static_cast<double>(v.y) + static_cast<double>(v.z) + static_cast<double>(v.w);
    ss += static_cast<double>(v.x) * v.x + static_cast<double>(v.y) * v.y + static_cast<double>(v.z) * v.z + static_cast<double>(v.w) * v.w;
  }
  for (long long i = (n4 << 2) + tid; i < count; i += stride) { const double v = a[i]; s += v; ss += v * v; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); ss += __shfl_xor_sync(0xffffffffu, ss, o); }
  __shared__ double sh[2][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double bs = 0.0, bss = 0.0;
    for (int w = 0; w < 8; ++w) { bs += sh[0][w]; bss += sh[1][w]; }
    atomicAdd(&stats[1], bs);
    atomicAdd(&stats[2], bss);
    if (blockIdx.x == 0) stats[0] = static_cast<double>(count);
  }
}

__global__ void __launch_bounds__(256) adv_apply_kernel(float* __restrict__ a, long long count, const double* __restrict__ stats,
                                                        float eps) {
  const double n = stats[0], sum = stats[1], sumsq = stats[2];
  const double mean_d = sum / n;
  double var = (sumsq - sum * mean_d) / (n - 1.0);                     // torch.Tensor.std(): unbiased (N-1)
  if (var < 0.0) var = 0.0;
  const float mean = static_cast<float>(mean_d);
  const float denom = __fadd_rn(static_cast<float>(sqrt(var)), eps);   // std + 1e-5 in fp32 (process_batch.py:142)
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n4 = count >> 2;
  float4* a4 = reinterpret_cast<float4*>(a);
  for (long long i = tid; i < n4; i += stride) {
    float4 v = a4[i];
    v.x = __fdiv_rn(__fsub_rn(v.x, mean), denom); v.y = __fdiv_rn(__fsub_rn(v.y, mean), denom);
    v.z = __fdiv_rn(__fsub_rn(v.z, mean), denom); v.w = __fdiv_rn(__fsub_rn(v.w, mean), denom);
    a4[i] = v;
  }
  for (long long i = (n4 << 2) + tid; i < count; i += stride) a[i] = __fdiv_rn(__fsub_rn(a[i], mean), denom);
}


// ---- policy inputs (RL/models/policy.py:168-190 obs_to_torch / act_masks_to_torch, process_batch.py:43-51, :80-84) -----
// The reference converts every observation dict and mask list to fp32 tensors on the host, env by env, and stacks them.  Here
// the packed uint8 rows are already in HBM and ONE launch expands a batch of them into the tensors the policy network reads:
//   features [B][1792] (fp32 or bf16): the 1787 numeric features in the reference's order (the two ratio features rescaled:
//            len / 8, knights / 4 — exact), 5 zero columns so that every row starts on a 16-byte boundary;
//   lists    [5][B][25] int64: the padded development-card lists (card + 1, 0 = pad; player_modules.py:49-53 tensor form);
//   head masks, one flat buffer: head h at element CATAN_MASK_<h> * B, shaped [B][dim], or [types][B][dim] for the
//            type-conditional heads 1, 6 and 9 (policy.py:188-189 transposes them that way).
// HBM-bound streaming: per row 1920 + 336 B read, 1792 * 4 + 125 * 8 + 325 * 4 = 9 468 B written (fp32).
struct PolicyInArgs {
  const uint8_t* obs;         // [B][CATAN_OBS_STRIDE]
  const uint8_t* masks;       // [B][CATAN_MASK_STRIDE] or null
  const int32_t* index;       // [B] or null: batch row b is row index[b] of obs / masks (the env list of one policy)
  void* features;             // [B][CATAN_POLICY_FEATURE_STRIDE]
  long long* lists;           // [5][B][CATAN_OBS_DEV_PAD]
  void* head_masks;           // [CATAN_MASK_ENTRIES * B] or null
  unsigned B;
  unsigned long long magic_B; // 2^64 / B + 1: n / B = umul64hi(n, magic_B) for every 32-bit n (B > 1)
};

#define CATAN_DIV_MAGIC(d) (0xFFFFFFFFFFFFFFFFull / (d) + 1ull)
__constant__ int c_head_off[13] = {CATAN_MASK_TYPE, CATAN_MASK_CORNER, CATAN_MASK_EDGE, CATAN_MASK_TILE, CATAN_MASK_DEV,
                                   CATAN_MASK_ACCEPT, CATAN_MASK_PLAYER, CATAN_MASK_GIVE, CATAN_MASK_RECV, CATAN_MASK_RES_A,
                                   CATAN_MASK_RES_B, CATAN_MASK_DISCARD, CATAN_MASK_ENTRIES};
// last dimension of every head; types = (off[h + 1] - off[h]) / dim (3 for the corner and player heads, 4 for resource A)
__constant__ int c_head_dim[13] = {13, 54, 73, 19, 5, 2, 3, 6, 6, 5, 5, 5, 1};
__constant__ unsigned long long c_head_magic[13] = {
    CATAN_DIV_MAGIC(13), CATAN_DIV_MAGIC(54), CATAN_DIV_MAGIC(73), CATAN_DIV_MAGIC(19), CATAN_DIV_MAGIC(5), CATAN_DIV_MAGIC(2),
    CATAN_DIV_MAGIC(3),  CATAN_DIV_MAGIC(6),  CATAN_DIV_MAGIC(6),  CATAN_DIV_MAGIC(5),  CATAN_DIV_MAGIC(5), CATAN_DIV_MAGIC(5), 0};

__device__ __forceinline__ unsigned div_magic(unsigned n, unsigned long long magic) {
  return static_cast<unsigned>(__umul64hi(static_cast<unsigned long long>(n), magic));
}
__device__ __forceinline__ unsigned div_B(unsigned n, const PolicyInArgs& A) { return A.B == 1 ? n : div_magic(n, A.magic_B); }
__device__ __forceinline__ size_t src_row(unsigned b, const PolicyInArgs& A) { return A.index == nullptr ? b : static_cast<size_t>(__ldg(A.index + b)); }

__device__ __forceinline__ float ratio_scale(int col) {
  if (col < CATAN_OBS_CUR_MAIN) return 1.0f;
  const int r = col < CATAN_OBS_OTHER_MAIN ? col - CATAN_OBS_CUR_MAIN - CATAN_OBS_CUR_LR_LEN
                                           : (col - CATAN_OBS_OTHER_MAIN) % CATAN_OBS_OTHER_MAIN_DIM - CATAN_OBS_OTH_LR_LEN;
  return r == 0 ? 0.125f : r == 2 ? 0.25f : 1.0f;                  // wrapper.py:613-627: len / 8.0, knights / 4.0
}

__device__ __forceinline__ void store4(float* p, size_t i4, const float* v) {
  __stcs(reinterpret_cast<float4*>(p) + i4, make_float4(v[0], v[1], v[2], v[3]));
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, size_t i4, const float* v) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
  uint2 u;
  u.x = *reinterpret_cast<const unsigned*>(&lo);
  u.y = *reinterpret_cast<const unsigned*>(&hi);
  __stcs(reinterpret_cast<uint2*>(p) + i4, u);
}
__device__ __forceinline__ void store1(float* p, size_t i, float a) { p[i] = a; }
__device__ __forceinline__ void store1(__nv_bfloat16* p, size_t i, float a) { p[i] = __float2bfloat16_rn(a); }

// Every thread keeps four independent loads in flight per loop trip (the kernel is a pure stream: without that it is bound
// by one DRAM latency per trip), all index arithmetic is 32-bit with multiply-high divisions, every store is a 16-byte
// (8-byte for bf16 quads) vector that the warp writes contiguously.
template <typename T>
__global__ void __launch_bounds__(256) policy_inputs_kernel(const __grid_constant__ PolicyInArgs A) {
  __shared__ uint4 s_rows[32 * (CATAN_MASK_STRIDE / 16)];          // a tile of 32 mask rows
  __shared__ uint2 s_col[CATAN_MASK_ENTRIES];                      // mask column -> (segment offset | dim << 16, division magic)
  const unsigned stride = gridDim.x * blockDim.x;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned B = A.B;
  {  // numeric features: one 4-byte load -> one 16-byte (fp32) / 8-byte (bf16) store.  The grid is a multiple of 7 blocks, so
     // the grid stride is a multiple of the 448 groups of a row: a thread keeps ITS four columns for the whole kernel and
     // their scale factors (1, 1/8, 1/4, or 0 for the pad columns) are computed once.
    constexpr unsigned G = CATAN_POLICY_FEATURE_STRIDE / 4;
    const unsigned* src = reinterpret_cast<const unsigned*>(A.obs);
    T* dst = static_cast<T*>(A.features);
    const unsigned row0 = tid / G, c4 = tid - row0 * G, row_step = stride / G;
    float scale[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) scale[k] = c4 * 4 + k < CATAN_OBS_FEATURES ? ratio_scale(static_cast<int>(c4 * 4) + k) : 0.0f;
    for (unsigned r0 = row0; r0 < B; r0 += 4 * row_step) {
      unsigned w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned row = r0 + u * row_step;
        if (row < B) w[u] = __ldcs(src + src_row(row, A) * (CATAN_OBS_STRIDE / 4) + c4);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned row = r0 + u * row_step;
        if (row < B) {
          float v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = static_cast<float>((w[u] >> (8 * k)) & 0xffu) * scale[k];
          store4(dst, static_cast<size_t>(row) * G + c4, v);
        }
      }
    }
  }
  {  // development-card lists: four bytes -> four int64 (two 16-byte stores); flat index i = (li * B + b) * 25 + j
    const unsigned total = 5 * CATAN_OBS_DEV_PAD * B;
    for (unsigned i0 = tid * 4; i0 < total; i0 += stride * 4) {
      const unsigned q = i0 / CATAN_OBS_DEV_PAD;
      unsigned j = i0 - q * CATAN_OBS_DEV_PAD, li = div_B(q, A), b = q - li * B;
      long long v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (i0 + k < total) v[k] = A.obs[src_row(b, A) * CATAN_OBS_STRIDE + CATAN_OBS_DEV_LISTS + li * CATAN_OBS_DEV_PAD + j];
        if (++j == CATAN_OBS_DEV_PAD) { j = 0; if (++b == B) { b = 0; ++li; } }
      }
      if (i0 + 3 < total) {
        __stcs(reinterpret_cast<longlong2*>(A.lists + i0), make_longlong2(v[0], v[1]));
        __stcs(reinterpret_cast<longlong2*>(A.lists + i0) + 1, make_longlong2(v[2], v[3]));
      } else {
        for (int k = 0; k < 4; ++k) if (i0 + k < total) A.lists[i0 + k] = v[k];
      }
    }
  }
  if (A.masks != nullptr && A.head_masks != nullptr) {
    // masks, head-major: head h owns output elements [off[h] * B, off[h + 1] * B); a (head, type) segment s of `dim` entries at
    // mask byte seg_off is, for 32 consecutive rows b0 .. b0 + 31, ONE contiguous run of 32 * dim output elements starting at
    // seg_off * B + b0 * dim.  A block stages the 32 mask rows of a tile in shared memory (16-byte loads) and streams the 19
    // runs out of it: consecutive threads write consecutive elements.
    T* dst = static_cast<T*>(A.head_masks);
    for (int c = threadIdx.x; c < CATAN_MASK_ENTRIES; c += blockDim.x) {
      int h = 0;
#pragma unroll
      for (int k = 1; k < 12; ++k) h += c >= c_head_off[k];
      const unsigned dim = c_head_dim[h], seg = c_head_off[h] + (c - c_head_off[h]) / dim * dim;
      s_col[c] = make_uint2(seg | (dim << 16), ((1u << 20) + dim - 1) / dim);   // e / dim = (e * magic) >> 20 for e < 32 * 73
    }
    const unsigned tiles = (B + 31) / 32;
    for (unsigned tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      const unsigned b0 = tile * 32, rows = min(32u, B - b0);
      __syncthreads();                                             // the previous tile has been read (and s_col written)
      for (unsigned q = threadIdx.x; q < rows * (CATAN_MASK_STRIDE / 16); q += blockDim.x) {
        const unsigned r = q / (CATAN_MASK_STRIDE / 16), w = q - r * (CATAN_MASK_STRIDE / 16);
        s_rows[q] = __ldcs(reinterpret_cast<const uint4*>(A.masks + src_row(b0 + r, A) * CATAN_MASK_STRIDE) + w);
      }
      __syncthreads();
      const uint8_t* rows_b = reinterpret_cast<const uint8_t*>(s_rows);
      for (unsigned p = threadIdx.x; p < rows * CATAN_MASK_ENTRIES; p += blockDim.x) {
        const unsigned c = rows == 32 ? p >> 5 : p / rows;         // p = seg_off * rows + e: any column of the segment finds it
        const uint2 sc = s_col[c];
        const unsigned seg = sc.x & 0xffffu, dim = sc.x >> 16;
        const unsigned e = p - seg * rows, bl = (e * sc.y) >> 20, j = e - bl * dim;
        store1(dst, static_cast<size_t>(seg) * B + static_cast<size_t>(b0) * dim + e,
               static_cast<float>(rows_b[bl * CATAN_MASK_STRIDE + seg + j]));
      }
    }
  }
}


// ---- masked categorical head (RL/distributions.py:11-40 FixedCategorical / Categorical.forward) ----------------------
// The tail of every action head in the reference: logits + log(mask) -> Categorical -> sample / mode, log_probs, entropy,
// ~10 small torch launches per head and 12 heads per decision.  One warp per row does all of it in one launch:
//   l_i = x_i (mask_i != 0) or -inf;  m = max l;  S = sum exp(l_i - m);  log p_i = l_i - m - log S;
//   entropy = -sum p_i log p_i over p_i > 0 (distributions.py:18-20) = log S - sum e_i (l_i - m) / S;
//   action: given (evaluate), argmax p (mode, first maximum), or inverse CDF of a caller-supplied uniform (sample).
// expf / logf are the IEEE-accurate library versions; results agree with torch to ~1e-6.
struct CategoricalArgs {
  const float* logits;        // [B][D]
  const float* mask;          // [B][D] or null (no mask)
  const long long* given;     // [B] or null
  const float* uniform;       // [B] or null
  long long* action;          // [B]
  float* logp;                // [B]
  float* entropy;             // [B] or null
  int B, D;
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) masked_categorical_kernel(const __grid_constant__ CategoricalArgs A) {
  const int lane = threadIdx.x & 31;
  const int row = static_cast<int>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= A.B) return;
  const int D = A.D;
  const float* x = A.logits + static_cast<size_t>(row) * D;
  const float* mk = A.mask == nullptr ? nullptr : A.mask + static_cast<size_t>(row) * D;
  const float NEG_INF = __int_as_float(0xff800000);
  auto logit = [&](int i) { return (mk == nullptr || mk[i] != 0.0f) ? x[i] : NEG_INF; };
  float m = NEG_INF;
  int arg = 0x7fffffff;
  for (int i = lane; i < D; i += 32) { const float l = logit(i); if (l > m) { m = l; arg = i; } }   // first maximum of this lane
  const float mx = warp_max(m);
  {  // mode: the first index holding the maximum
    int cand = (m == mx && arg != 0x7fffffff) ? arg : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    arg = cand == 0x7fffffff ? 0 : cand;
  }
  float s = 0.0f, sl = 0.0f;
  for (int i = lane; i < D; i += 32) {
    const float l = logit(i);
    if (l != NEG_INF) { const float d = l - mx, e = expf(d); s += e; sl += e * d; }
  }
  const float S = warp_sum(s), SL = warp_sum(sl), logS = logf(S);
  long long a;
  if (A.given != nullptr) {
    a = A.given[row];
  } else if (A.uniform != nullptr) {
    // inverse CDF in index order: the first i with e_0 + ... + e_i > u * S; masked entries add nothing and are never picked
    const float target = A.uniform[row] * S;
    float run = 0.0f;
    int pick = -1, last = -1;
    for (int base = 0; base < D && pick < 0; base += 32) {
      const int i = base + lane;
      const float l = i < D ? logit(i) : NEG_INF;
      const float e = l != NEG_INF ? expf(l - mx) : 0.0f;
      float c = e;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, c, o); if (lane >= o) c += t; }
      const unsigned hit = __ballot_sync(0xffffffffu, e > 0.0f && run + c > target);
      const unsigned legal = __ballot_sync(0xffffffffu, e > 0.0f);
      if (legal) last = base + 31 - __clz(legal);
      if (hit) pick = base + __ffs(hit) - 1;
      run += __shfl_sync(0xffffffffu, c, 31);
    }
    a = pick >= 0 ? pick : (last >= 0 ? last : arg);               // rounding at the top of the CDF: the last legal entry
  } else {
    a = arg;
  }
  if (lane == 0) {
    const int ai = static_cast<int>(a);
    const float la = (ai >= 0 && ai < D) ? logit(ai) : NEG_INF;
    if (A.given == nullptr) A.action[row] = a;
    A.logp[row] = (la - mx) - logS;
    if (A.entropy != nullptr) A.entropy[row] = logS - SL / S;
  }
}

}  // namespace catanb

static thread_local std::string g_ppo_error;
extern "C" const char* catan_last_error(void);
static int ppo_fail(cudaError_t e, const char* what);

extern "C" {

int catan_gae(const float* rewards_dev, const float* values_dev, const float* masks_dev, int T, int N, double gamma,
              double gae_lambda, float* returns_dev, float* advantages_dev, void* stream) {
  if (!rewards_dev || !values_dev || !masks_dev || !returns_dev || !advantages_dev || T <= 0 || N <= 0) return ppo_fail(cudaErrorInvalidValue, "catan_gae: bad argument");
  catanb::DeviceScope device_scope_;                                 // the kernel runs where its buffers live; the caller's device comes back
  if (device_scope_.enter_for(rewards_dev, static_cast<cudaStream_t>(stream))) return ppo_fail(cudaErrorInvalidDevice, "cannot make the device of the buffers current");
  const int threads = 256, blocks = (N + threads - 1) / threads;
  // torch multiplies fp32 tensors by Python doubles rounded to fp32; gamma*lambda is formed in double first
  catanb::gae_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      rewards_dev, values_dev, masks_dev, T, N, static_cast<float>(gamma), static_cast<float>(gamma * gae_lambda), returns_dev,
      advantages_dev);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : ppo_fail(e, "catan_gae launch");
}

static int stream_grid(long long count) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long want = (count / 4 + 255) / 256;
  long long cap = static_cast<long long>(sms) * 8;   // 8 resident 256-thread blocks per SM, whole waves
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}

int catan_adv_stats(const float* advantages_dev, long long count, double* stats_dev, void* stream) {
  if (!advantages_dev || !stats_dev || count <= 1) return ppo_fail(cudaErrorInvalidValue, "catan_adv_stats: bad argument");
  if (reinterpret_cast<uintptr_t>(advantages_dev) & 15) return ppo_fail(cudaErrorInvalidValue, "catan_adv_stats: advantages must be 16-byte aligned");
  catanb::DeviceScope device_scope_;                                 // the kernel runs where its buffers live; the caller's device comes back
  if (device_scope_.enter_for(advantages_dev, static_cast<cudaStream_t>(stream))) return ppo_fail(cudaErrorInvalidDevice, "cannot make the device of the buffers current");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(stats_dev, 0, 3 * sizeof(double), s);
  if (e != cudaSuccess) return ppo_fail(e, "catan_adv_stats memset");
  catanb::adv_stats_kernel<<<stream_grid(count), 256, 0, s>>>(advantages_dev, count, stats_dev);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : ppo_fail(e, "catan_adv_stats launch");
}

int catan_adv_apply(float* advantages_dev, long long count, const double* stats_dev, double eps, void* stream) {
  if (!advantages_dev || !stats_dev || count <= 1) return ppo_fail(cudaErrorInvalidValue, "catan_adv_apply: bad argument");
  if (reinterpret_cast<uintptr_t>(advantages_dev) & 15) return ppo_fail(cudaErrorInvalidValue, "catan_adv_apply: advantages must be 16-byte aligned");
  catanb::DeviceScope device_scope_;                                 // the kernel runs where its buffers live; the caller's device comes back
  if (device_scope_.enter_for(advantages_dev, static_cast<cudaStream_t>(stream))) return ppo_fail(cudaErrorInvalidDevice, "cannot make the device of the buffers current");
  catanb::adv_apply_kernel<<<stream_grid(count), 256, 0, static_cast<cudaStream_t>(stream)>>>(advantages_dev, count, stats_dev,
                                                                                           static_cast<float>(eps));
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : ppo_fail(e, "catan_adv_apply launch");
}

}  // extern "C"

// error text is shared with catan_kernels.cu through catan_set_last_error
extern "C" void catan_set_last_error(const char* msg);
static int ppo_fail(cudaError_t e, const char* what) {
  std::string m = std::string(what) + ": " + cudaGetErrorString(e);
  catan_set_last_error(m.c_str());
  return -1;
}

// =================================================================================================
// rollout collector: the per-env state machine of GamesAndPoliciesManager.gather_rollouts
// (RL/ppo/game_manager.py:69-140) and the buffer layout of BatchProcessor.process_rollouts
// (RL/ppo/process_batch.py:37-104), one warp per env per tick.
// =================================================================================================
namespace catanb {

struct RolloutArgs {
  catan_rollout_t r;
  const uint8_t* env_obs;      // [N][OBS_STRIDE]   post-step observation (the env's bound buffer)
  const uint8_t* env_masks;    // [N][MASK_STRIDE]
  const float* env_reward;     // [N][4]
  const uint8_t* env_info;     // [N][INFO_STRIDE]
  const int32_t* actions;      // [N][ACTION_WORDS] the actions applied this tick
  const float* logp;           // [N]
  const uint8_t* stepped;      // [N] envs that were stepped this tick (nullptr = all)
  int begin;                   // 1: (re)start of a rollout (GamesAndPoliciesManager.reset / _after_rollouts), no step happened
  int fresh;                   // with begin: 1 = after env.reset() (manager.reset), 0 = carry the last observation over
};

__device__ __forceinline__ void copy_row16(uint8_t* dst, const uint8_t* src, int bytes, int lane) {
  const int4* s = reinterpret_cast<const int4*>(src);
  int4* d = reinterpret_cast<int4*>(dst);
  for (int i = lane; i < bytes / 16; i += 32) d[i] = s[i];
}

__global__ void __launch_bounds__(128) rollout_store_kernel(const RolloutArgs A) {
  const int e = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  const catan_rollout_t& R = A.r;
  if (e >= R.N) return;
  const int T = R.T, N = R.N;
  int32_t* cur = R.cursors + static_cast<size_t>(e) * 4;          // n_obs, n_act, n_rew, n_tm
  double* acc = R.acc + static_cast<size_t>(e) * 4;                // doubles: the reference sums Python floats (game_manager.py:94-95)
  const int active = R.active_pid[e];
  int n_obs = cur[0], n_act = cur[1], n_rew = cur[2], n_tm = cur[3];
  int flags = R.flags[e];                                          // bit 0: done_since_prev_turn, bit 1: last terminal mask
  const uint8_t* info = A.env_info + static_cast<size_t>(e) * CATAN_INFO_STRIDE;
  __syncwarp();
  bool store_obs = false;
  if (A.begin) {
    if (A.fresh) {                                                 // game_manager.py:34-56
      n_obs = 0; n_act = 0; n_rew = 0; n_tm = 0;
      if (lane == 0) R.tmasks[e] = 1.0f;
      n_tm = 1; flags |= 2;
      store_obs = info[CATAN_INFO_ACTOR] == active;
    } else {                                                       // _after_rollouts, game_manager.py:142-150
      if (n_obs > 0) copy_row16(R.obs + static_cast<size_t>(e) * CATAN_OBS_STRIDE,
                                R.obs + (static_cast<size_t>(n_obs - 1) * N + e) * CATAN_OBS_STRIDE, CATAN_OBS_STRIDE, lane);
      // the env has been frozen since that observation was stored, so its bound mask row is still the one that goes with it
      if (n_obs > 0) copy_row16(R.masks + static_cast<size_t>(e) * CATAN_MASK_STRIDE, A.env_masks + static_cast<size_t>(e) * CATAN_MASK_STRIDE,
                                CATAN_MASK_STRIDE, lane);
      if (lane == 0 && n_tm > 0) R.tmasks[e] = (flags & 2) ? 1.0f : 0.0f;   // terminal_masks[-1], even when it lay beyond the buffer
      n_obs = n_obs > 0 ? 1 : 0; n_tm = n_tm > 0 ? 1 : 0; n_act = 0; n_rew = 0;
    }
    if (lane < 4) acc[lane] = 0.0;                                 // game_manager.py:76-77: per-call locals
    flags &= 2;                                                    // bit 1 = value of the last terminal mask, kept
  } else if ((A.stepped == nullptr || A.stepped[e]) && n_obs < T + 1) {
    const int pg = info[CATAN_INFO_ACTED], n_pg_pre = info[CATAN_INFO_ACTOR_PRE], n_pg = info[CATAN_INFO_ACTOR];
    const bool done = info[CATAN_INFO_DONE] != 0;
    double a_acc = acc[active - 1] + static_cast<double>(A.env_reward[static_cast<size_t>(e) * 4 + active - 1]);     // :94-95
    __syncwarp();
    if (lane < 4) acc[lane] += static_cast<double>(A.env_reward[static_cast<size_t>(e) * 4 + lane]);
    bool reward_updated = false;
    // The reference's lists may grow past what process_rollouts reads (T rows, T+1 for obs / terminal masks): e.g. the
    // terminal-mask list leads by one when the first observation of a rollout was not the active seat's.  Entries
    // beyond the buffers are counted in the cursors but not stored.
    if (pg == active) {                                            // :102-105
      if (n_act < T && lane < CATAN_ACTION_WORDS) R.actions[(static_cast<size_t>(n_act) * N + e) * CATAN_ACTION_WORDS + lane] =
          A.actions[static_cast<size_t>(e) * CATAN_ACTION_WORDS + lane];
      if (n_act < T && lane == 0) R.logp[static_cast<size_t>(n_act) * N + e] = A.logp[e];
      n_act += 1;
    }
    if (n_pg_pre == active && n_act > 0 && !(flags & 1)) {         // :106-110
      if (n_rew < T && lane == 0) R.rewards[static_cast<size_t>(n_rew) * N + e] = static_cast<float>(a_acc);
      n_rew += 1; a_acc = 0.0; reward_updated = true;
      __syncwarp();
      if (lane == 0) acc[active - 1] = 0.0;
    }
    if (done) {                                                    // :112-124
      if (n_tm <= T && lane == 0) R.tmasks[static_cast<size_t>(n_tm) * N + e] = 0.0f;
      n_tm += 1;
      flags &= ~3;
      if (!reward_updated) {
        if (n_rew < T && lane == 0) R.rewards[static_cast<size_t>(n_rew) * N + e] = static_cast<float>(a_acc);
        n_rew += 1;
      }
      __syncwarp();
      if (lane < 4) acc[lane] = 0.0;
    }
    if (n_pg == active) {                                          // :128-133
      if (!done && !(flags & 1)) {
        if (n_tm <= T && lane == 0) R.tmasks[static_cast<size_t>(n_tm) * N + e] = 1.0f;
        n_tm += 1; flags |= 2;
      }
      flags &= ~1;
      store_obs = true;
    } else if (done) {
      flags |= 1;                                                  // :134-136
    }
  }
  if (store_obs) {
    copy_row16(R.obs + (static_cast<size_t>(n_obs) * N + e) * CATAN_OBS_STRIDE, A.env_obs + static_cast<size_t>(e) * CATAN_OBS_STRIDE,
               CATAN_OBS_STRIDE, lane);
    if (n_obs < T)   // the masks of the (T+1)-th observation are never used (game_manager.py:78: the loop ends first)
      copy_row16(R.masks + (static_cast<size_t>(n_obs) * N + e) * CATAN_MASK_STRIDE, A.env_masks + static_cast<size_t>(e) * CATAN_MASK_STRIDE,
                 CATAN_MASK_STRIDE, lane);
    n_obs += 1;
  }
  __syncwarp();
  if (lane == 0) {
    cur[0] = n_obs; cur[1] = n_act; cur[2] = n_rew; cur[3] = n_tm;
    R.flags[e] = static_cast<uint8_t>(flags);
    if (R.collecting) R.collecting[e] = n_obs < T + 1;             // step mask of the next tick
  }
}

// ---- minibatch gather (RL/ppo/process_batch.py:169-200, generator_standard) ------------------------
// One minibatch = B rows picked by a random permutation of the T*N (time, env) pairs of the rollout buffers.  The
// reference indexes its CPU tensors key by key and ships every piece to the device; here the buffers already live in HBM
// and one launch copies the B rows of all eight arrays into contiguous minibatch tensors: one warp per row, 16-byte
// vectors (120 for the observation, 21 for the masks, 5 for the action), the six scalars by lanes 0..5.
// Algorithmic bytes per row: 2 x (1920 + 336 + 80 + 5 x 4) = 4712.
struct GatherArgs {
  catan_rollout_t r;
  const float* values;        // [T+1][N]
  const float* returns;       // [T][N]
  const float* advantages;    // [T][N]
  const int32_t* indices;     // [B] flat t * N + n, t < T
  catan_minibatch_t out;
  int B;
};

__global__ void __launch_bounds__(256) minibatch_gather_kernel(const __grid_constant__ GatherArgs A) {
  const int lane = threadIdx.x & 31;
  const int row = static_cast<int>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= A.B) return;
  const size_t src = static_cast<size_t>(A.indices[row]);          // (t, n) flattened: the same offset in every [T(+1)][N] array
  {
    const uint4* s = reinterpret_cast<const uint4*>(A.r.obs + src * CATAN_OBS_STRIDE);
    uint4* d = reinterpret_cast<uint4*>(A.out.obs + static_cast<size_t>(row) * CATAN_OBS_STRIDE);
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) if (lane + 32 * k < CATAN_OBS_STRIDE / 16) v[k] = __ldcs(s + lane + 32 * k);   // read once, written once
#pragma unroll
    for (int k = 0; k < 4; ++k) if (lane + 32 * k < CATAN_OBS_STRIDE / 16) __stcs(d + lane + 32 * k, v[k]);
  }
  if (lane < CATAN_MASK_STRIDE / 16)
    __stcs(reinterpret_cast<uint4*>(A.out.masks + static_cast<size_t>(row) * CATAN_MASK_STRIDE) + lane,
           __ldcs(reinterpret_cast<const uint4*>(A.r.masks + src * CATAN_MASK_STRIDE) + lane));
  if (lane < CATAN_ACTION_WORDS / 4)
    __stcs(reinterpret_cast<uint4*>(A.out.actions + static_cast<size_t>(row) * CATAN_ACTION_WORDS) + lane,
           __ldcs(reinterpret_cast<const uint4*>(A.r.actions + src * CATAN_ACTION_WORDS) + lane));
  if (lane < 5) {
    const float* sp = lane == 0 ? A.r.logp : lane == 1 ? A.values : lane == 2 ? A.returns : lane == 3 ? A.r.tmasks : A.advantages;
    float* dp = lane == 0 ? A.out.logp : lane == 1 ? A.out.values : lane == 2 ? A.out.returns : lane == 3 ? A.out.tmasks : A.out.advantages;
    dp[row] = sp[src];                                               // values[:-1], masks[:-1]: t < T, so the flat offset is the same
  }
}

// ---- per-seat policy routing (RL/ppo/game_manager.py:21-31, :82-93) -----------------------------------
// The reference keeps a PlayerId -> policy map per env (four policies: the learner and three earlier snapshots) and, env by
// env, lets the policy of the player whose decision it is act.  Vectorised, a tick needs, per policy, the list of envs it
// acts for: block k compacts the env indices n with policy_map[n][actor(n) - 1] == k, in ascending order (ballot + scan),
// so that every policy runs ONE batched forward over its own envs.
struct RouteArgs {
  const uint8_t* info;        // [N][CATAN_INFO_STRIDE]: CATAN_INFO_ACTOR = PlayerId that takes the next decision
  const uint8_t* policy_map;  // [N][4]: policy index of PlayerId p + 1
  const uint8_t* active;      // [N] or null: envs with a zero byte are left out (frozen by catan_step_masked)
  int32_t* counts;            // [K]
  int32_t* lists;             // [K][N]
  int N;
};

__global__ void __launch_bounds__(1024) route_kernel(const __grid_constant__ RouteArgs A) {
  __shared__ int warp_count[32];
  __shared__ int base;
  const int k = static_cast<int>(blockIdx.x), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base = 0;
  __syncthreads();
  int32_t* out = A.lists + static_cast<size_t>(k) * A.N;
  for (int n0 = 0; n0 < A.N; n0 += 1024) {
    const int n = n0 + tid;
    bool hit = false;
    if (n < A.N && (A.active == nullptr || A.active[n] != 0)) {
      const int actor = A.info[static_cast<size_t>(n) * CATAN_INFO_STRIDE + CATAN_INFO_ACTOR];
      hit = actor >= 1 && actor <= 4 && A.policy_map[static_cast<size_t>(n) * 4 + actor - 1] == k;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) warp_count[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 32; ++w) { const int c = warp_count[w]; before += w < warp ? c : 0; total += c; }
    if (hit) out[base + before + __popc(bal & ((1u << lane) - 1u))] = n;
    __syncthreads();
    if (tid == 0) base += total;
  }
  __syncthreads();
  if (tid == 0) A.counts[k] = base;
}

}  // namespace catanb

extern "C" int catan_rollout_store(const catan_rollout_t* rollout, const uint8_t* env_obs_dev, const uint8_t* env_masks_dev,
                                   const float* env_reward_dev, const uint8_t* env_info_dev, const int32_t* actions_dev,
                                   const float* logp_dev, const uint8_t* stepped_dev, int begin, int fresh, void* stream) {
  if (!rollout || !env_obs_dev || !env_masks_dev || !env_info_dev || rollout->N <= 0 || rollout->T <= 0)
    return ppo_fail(cudaErrorInvalidValue, "catan_rollout_store: bad argument");
  if (!begin && (!env_reward_dev || !actions_dev || !logp_dev)) return ppo_fail(cudaErrorInvalidValue, "catan_rollout_store: bad argument");
  catanb::DeviceScope device_scope_;                                 // the kernel runs where its buffers live; the caller's device comes back
  if (device_scope_.enter_for(env_obs_dev, static_cast<cudaStream_t>(stream))) return ppo_fail(cudaErrorInvalidDevice, "cannot make the device of the buffers current");
  catanb::RolloutArgs A;
  A.r = *rollout;
  A.env_obs = env_obs_dev; A.env_masks = env_masks_dev; A.env_reward = env_reward_dev; A.env_info = env_info_dev;
  A.actions = actions_dev; A.logp = logp_dev; A.stepped = stepped_dev; A.begin = begin; A.fresh = fresh;
  catanb::rollout_store_kernel<<<(rollout->N + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(A);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : ppo_fail(e, "catan_rollout_store launch");
}

extern "C" int catan_minibatch_gather(const catan_rollout_t* rollout, const float* values_dev, const float* returns_dev,
                                      const float* advantages_dev, const int32_t* indices_dev, int B, const catan_minibatch_t* out,
                                      void* stream) {
  if (B == 0) return 0;                                              // (an empty batch has null pointers)
  if (!rollout || !values_dev || !returns_dev || !advantages_dev || !indices_dev || !out || B < 0 || rollout->N <= 0 || rollout->T <= 0)
    return ppo_fail(cudaErrorInvalidValue, "catan_minibatch_gather: bad argument");
  if (!out->obs || !out->masks || !out->actions || !out->logp || !out->values || !out->returns || !out->tmasks || !out->advantages)
    return ppo_fail(cudaErrorInvalidValue, "catan_minibatch_gather: null output buffer");
  if ((reinterpret_cast<uintptr_t>(out->obs) | reinterpret_cast<uintptr_t>(out->masks) | reinterpret_cast<uintptr_t>(out->actions)) & 15)
    return ppo_fail(cudaErrorInvalidValue, "catan_minibatch_gather: row buffers must be 16-byte aligned");
  catanb::DeviceScope device_scope_;                                 // the kernel runs where its buffers live; the caller's device comes back
  if (device_scope_.enter_for(values_dev, static_cast<cudaStream_t>(stream))) return ppo_fail(cudaErrorInvalidDevice, "cannot make the device of the buffers current");
  catanb::GatherArgs A;
  A.r = *rollout; A.values = values_dev; A.returns = returns_dev; A.advantages = advantages_dev; A.indices = indices_dev; A.out = *out; A.B = B;
  catanb::minibatch_gather_kernel<<<(B + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(A);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : ppo_fail(e, "catan_minibatch_gather launch");
}

extern "C" int catan_route_by_policy(const uint8_t* env_info_dev, const uint8_t* policy_map_dev, const uint8_t* active_dev, int N,
                                     int n_policies, int32_t* counts_dev, int32_t* lists_dev, void* stream) {
  if (!env_info_dev || !policy_map_dev || !counts_dev || !lists_dev || N <= 0 || n_policies <= 0 || n_policies > 255)
    return ppo_fail(cudaErrorInvalidValue, "catan_route_by_policy: bad argument");
  catanb::DeviceScope device_scope_;                                 // the kernel runs where its buffers live; the caller's device comes back
  if (device_scope_.enter_for(env_info_dev, static_cast<cudaStream_t>(stream))) return ppo_fail(cudaErrorInvalidDevice, "cannot make the device of the buffers current");
  catanb::RouteArgs A;
  A.info = env_info_dev; A.policy_map = policy_map_dev; A.active = active_dev; A.counts = counts_dev; A.lists = lists_dev; A.N = N;
  catanb::route_kernel<<<n_policies, 1024, 0, static_cast<cudaStream_t>(stream)>>>(A);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : ppo_fail(e, "catan_route_by_policy launch");
}

extern "C" int catan_policy_inputs(const uint8_t* obs_rows_dev, const uint8_t* mask_rows_dev, const int32_t* row_index_dev, int B, int dtype, void* features_dev,
                                   int64_t* lists_dev, void* head_masks_dev, void* stream) {
  if (!obs_rows_dev || !features_dev || !lists_dev || B <= 0 || B > (1 << 22) || (dtype != CATAN_DTYPE_F32 && dtype != CATAN_DTYPE_BF16) ||
      (mask_rows_dev == nullptr) != (head_masks_dev == nullptr))
    return ppo_fail(cudaErrorInvalidValue, "catan_policy_inputs: bad argument");
  catanb::DeviceScope device_scope_;                                 // the kernel runs where its buffers live; the caller's device comes back
  if (device_scope_.enter_for(obs_rows_dev, static_cast<cudaStream_t>(stream))) return ppo_fail(cudaErrorInvalidDevice, "cannot make the device of the buffers current");
  catanb::PolicyInArgs A;
  A.obs = obs_rows_dev; A.masks = mask_rows_dev; A.index = row_index_dev; A.features = features_dev; A.lists = reinterpret_cast<long long*>(lists_dev);
  A.head_masks = head_masks_dev; A.B = static_cast<unsigned>(B);
  A.magic_B = 0xFFFFFFFFFFFFFFFFull / static_cast<unsigned>(B) + 1ull;
  const long long groups = static_cast<long long>(B) * (CATAN_POLICY_FEATURE_STRIDE / 4);
  // grid-stride, 8 blocks of 256 threads per SM; a multiple of 7 blocks (7 * 256 = 4 * 448) keeps every thread on its columns
  const int blocks = static_cast<int>(groups + 255 < 256LL * 1183 ? ((groups + 255) / 256 + 6) / 7 * 7 : 1183);
  if (dtype == CATAN_DTYPE_F32) catanb::policy_inputs_kernel<float><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(A);
  else catanb::policy_inputs_kernel<__nv_bfloat16><<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(A);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : ppo_fail(e, "catan_policy_inputs launch");
}

extern "C" int catan_masked_categorical(const float* logits_dev, const float* mask_dev, const int64_t* given_actions_dev,
                                        const float* uniforms_dev, int B, int D, int64_t* actions_dev, float* logp_dev,
                                        float* entropy_dev, void* stream) {
  if (!logits_dev || !logp_dev || B <= 0 || D <= 0 || (given_actions_dev == nullptr && actions_dev == nullptr))
    return ppo_fail(cudaErrorInvalidValue, "catan_masked_categorical: bad argument");
  catanb::DeviceScope device_scope_;                                 // the kernel runs where its buffers live; the caller's device comes back
  if (device_scope_.enter_for(logits_dev, static_cast<cudaStream_t>(stream))) return ppo_fail(cudaErrorInvalidDevice, "cannot make the device of the buffers current");
  catanb::CategoricalArgs A;
  A.logits = logits_dev; A.mask = mask_dev; A.given = reinterpret_cast<const long long*>(given_actions_dev); A.uniform = uniforms_dev;
  A.action = reinterpret_cast<long long*>(actions_dev); A.logp = logp_dev; A.entropy = entropy_dev; A.B = B; A.D = D;
  catanb::masked_categorical_kernel<<<(B + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(A);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : ppo_fail(e, "catan_masked_categorical launch");
}
