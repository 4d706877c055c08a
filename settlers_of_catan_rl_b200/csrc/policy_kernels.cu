// policy_kernels.cu — sm_100a kernels for the two places where the policy network (RL/models/*, kept in PyTorch) is bound by
// tiny inner dimensions rather than by its GEMMs (profiles/r2_notes.md: at 16 384 envs 35 % of a rollout tick was the library
// attention kernel on 19-token sequences and 31-40 % LayerNorm over 16- / 25- / 64-wide rows):
//
//   tile_attention_fwd / _bwd   self-attention of the tile encoder (RL/models/tile_encoder.py:43-57 with
//                               multi_headed_attention.py:28-39): 19 tokens, 4 heads of 16.  One sample's q, k, v (19 x 192 fp32)
//                               sit in shared memory; one thread per (head, query row) computes its 19 scores, the softmax and
//                               the weighted sum in registers.  The backward recomputes the probabilities.
//   ln_small_fwd / _bwd         LayerNorm over a last dimension of at most 64 (tile_encoder.py:38, :75-76; player_modules.py:29-33):
//                               LPR lanes per row, rows of a warp contiguous in memory, moments by shuffles.
//
// Both are exact restatements (fp32, same formulas); tests/test_gpu_policy_kernels.py checks them and their gradients against
// torch.  The PyTorch side (settlers_of_catan_rl_b200/policy_ops.py) wraps them as autograd functions.
#include <cuda_runtime.h>

#include "device_scope.cuh"

#include <stdint.h>
#include <cstdio>

#include "../../include/catan_b200.h"

extern "C" void catan_set_last_error(const char* msg);

namespace catanb {

constexpr int kS = 19, kH = 4, kD = 16, kE = kH * kD;        // tokens, heads, head dim, model dim
constexpr int kQkv = 3 * kE;                                 // 192 floats per token: q | k | v
constexpr int kRow = kQkv + 1;                               // padded row in shared memory: rows of a head fall into different banks
constexpr int kAttThreads = 256;
constexpr int kAttSamples = 3;                               // 3 x 76 = 228 working threads per block

struct AttSmemF { float qkv[kAttSamples][kS][kRow]; };

__global__ void __launch_bounds__(kAttThreads) tile_attention_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ y, int B) {
  extern __shared__ __align__(16) unsigned char att_raw[];
  AttSmemF& S = *reinterpret_cast<AttSmemF*>(att_raw);
  const int b0 = blockIdx.x * kAttSamples, nb = min(kAttSamples, B - b0);
  const float4* src = reinterpret_cast<const float4*>(qkv + static_cast<size_t>(b0) * kS * kQkv);
  for (int i = threadIdx.x; i < nb * kS * (kQkv / 4); i += kAttThreads) {     // coalesced 16-byte loads
    const float4 v = __ldcs(src + i);
    const int e = i * 4, s = e / (kS * kQkv), r = (e / kQkv) % kS, c = e % kQkv;
    float* d = &S.qkv[s][r][c];
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const int t = threadIdx.x, s = t / (kH * kS), u = t % (kH * kS), h = u / kS, i = u % kS;
  if (s >= nb) return;
  float q[kD], o[kD], p[kS];
#pragma unroll
  for (int d = 0; d < kD; ++d) { q[d] = S.qkv[s][i][h * kD + d] * 0.25f; o[d] = 0.f; }      // 1 / sqrt(16)
  float mx = -3.0e38f;
#pragma unroll
  for (int j = 0; j < kS; ++j) {
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < kD; ++d) a = fmaf(q[d], S.qkv[s][j][kE + h * kD + d], a);
    p[j] = a;
    mx = fmaxf(mx, a);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kS; ++j) { p[j] = __expf(p[j] - mx); sum += p[j]; }
  const float inv = 1.f / sum;
#pragma unroll
  for (int j = 0; j < kS; ++j) {
    const float w = p[j] * inv;
#pragma unroll
    for (int d = 0; d < kD; ++d) o[d] = fmaf(w, S.qkv[s][j][2 * kE + h * kD + d], o[d]);
  }
  float4* dst = reinterpret_cast<float4*>(y + (static_cast<size_t>(b0 + s) * kS + i) * kE + h * kD);
#pragma unroll
  for (int d = 0; d < kD; d += 4) dst[d / 4] = make_float4(o[d], o[d + 1], o[d + 2], o[d + 3]);
}

constexpr int kAttSamplesB = 2;                              // backward: 2 x 76 threads per block
struct AttSmemB {
  float qkv[kAttSamplesB][kS][kRow];
  float dy[kAttSamplesB][kS][kE + 1];
  float p[kAttSamplesB][kH][kS][kS + 1];
  float ds[kAttSamplesB][kH][kS][kS + 1];
};

__global__ void __launch_bounds__(kAttThreads) tile_attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dy,
                                                                         float* __restrict__ dqkv, int B) {
  extern __shared__ __align__(16) unsigned char att_raw[];
  AttSmemB& S = *reinterpret_cast<AttSmemB*>(att_raw);
  const int b0 = blockIdx.x * kAttSamplesB, nb = min(kAttSamplesB, B - b0);
  const float4* src = reinterpret_cast<const float4*>(qkv + static_cast<size_t>(b0) * kS * kQkv);
  for (int i = threadIdx.x; i < nb * kS * (kQkv / 4); i += kAttThreads) {
    const float4 v = __ldcs(src + i);
    const int e = i * 4, s = e / (kS * kQkv), r = (e / kQkv) % kS, c = e % kQkv;
    float* d = &S.qkv[s][r][c];
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  const float4* gsrc = reinterpret_cast<const float4*>(dy + static_cast<size_t>(b0) * kS * kE);
  for (int i = threadIdx.x; i < nb * kS * (kE / 4); i += kAttThreads) {
    const float4 v = __ldcs(gsrc + i);
    const int e = i * 4, s = e / (kS * kE), r = (e / kE) % kS, c = e % kE;
    float* d = &S.dy[s][r][c];
    d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
  }
  __syncthreads();
  const int t = threadIdx.x, s = t / (kH * kS), u = t % (kH * kS), h = u / kS, i = u % kS;
  const bool work = s < nb;
  float* out = dqkv + (static_cast<size_t>(b0 + (work ? s : 0)) * kS + i) * kQkv;
  if (work) {
    // row i of this head: probabilities (recomputed), dP = dO V^T, dS = P (dP - sum_j P dP), dq = dS K / sqrt(d)
    float q[kD], g[kD], p[kS], dp[kS];
#pragma unroll
    for (int d = 0; d < kD; ++d) { q[d] = S.qkv[s][i][h * kD + d] * 0.25f; g[d] = S.dy[s][i][h * kD + d]; }
    float mx = -3.0e38f;
#pragma unroll
    for (int j = 0; j < kS; ++j) {
      float a = 0.f, c = 0.f;
#pragma unroll
      for (int d = 0; d < kD; ++d) {
        a = fmaf(q[d], S.qkv[s][j][kE + h * kD + d], a);
        c = fmaf(g[d], S.qkv[s][j][2 * kE + h * kD + d], c);
      }
      p[j] = a; dp[j] = c;
      mx = fmaxf(mx, a);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kS; ++j) { p[j] = __expf(p[j] - mx); sum += p[j]; }
    const float inv = 1.f / sum;
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < kS; ++j) { p[j] *= inv; dot = fmaf(p[j], dp[j], dot); }
    float dq[kD];
#pragma unroll
    for (int d = 0; d < kD; ++d) dq[d] = 0.f;
#pragma unroll
    for (int j = 0; j < kS; ++j) {
      const float ds = p[j] * (dp[j] - dot);
      S.p[s][h][i][j] = p[j];
      S.ds[s][h][i][j] = ds;
#pragma unroll
      for (int d = 0; d < kD; ++d) dq[d] = fmaf(ds, S.qkv[s][j][kE + h * kD + d], dq[d]);
    }
    float4* dst = reinterpret_cast<float4*>(out + h * kD);
#pragma unroll
    for (int d = 0; d < kD; d += 4) dst[d / 4] = make_float4(dq[d] * 0.25f, dq[d + 1] * 0.25f, dq[d + 2] * 0.25f, dq[d + 3] * 0.25f);
  }
  __syncthreads();
  if (work) {
    // the same thread now owns KEY row j = i of its head: dk_j = sum_i dS_ij q_i / sqrt(d), dv_j = sum_i P_ij dO_i
    const int j = i;
    float dk[kD], dv[kD];
#pragma unroll
    for (int d = 0; d < kD; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
#pragma unroll
    for (int r = 0; r < kS; ++r) {
      const float ds = S.ds[s][h][r][j], pr = S.p[s][h][r][j];
#pragma unroll
      for (int d = 0; d < kD; ++d) {
        dk[d] = fmaf(ds, S.qkv[s][r][h * kD + d], dk[d]);
        dv[d] = fmaf(pr, S.dy[s][r][h * kD + d], dv[d]);
      }
    }
    float4* dstk = reinterpret_cast<float4*>(out + kE + h * kD);
    float4* dstv = reinterpret_cast<float4*>(out + 2 * kE + h * kD);
#pragma unroll
    for (int d = 0; d < kD; d += 4) {
      dstk[d / 4] = make_float4(dk[d] * 0.25f, dk[d + 1] * 0.25f, dk[d + 2] * 0.25f, dk[d + 3] * 0.25f);
      dstv[d / 4] = make_float4(dv[d], dv[d + 1], dv[d + 2], dv[d + 3]);
    }
  }
}

// ---- LayerNorm over a short last dimension --------------------------------------------------------
// LPR lanes share a row (LPR a power of two >= dim / 4, so a lane holds at most 4 elements: e = lane + k * LPR); the 32 / LPR
// rows of a warp are adjacent in memory, so a warp's k-th load covers one contiguous span.
template <int LPR>
__global__ void __launch_bounds__(256) ln_small_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                           float* __restrict__ y, float* __restrict__ stats, long long rows, int dim, float eps) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const long long warp = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) >> 5;
  const long long row = warp * RPW + lane / LPR;
  const bool live = row < rows;
  float v[4];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = sub + k * LPR;
    v[k] = (live && e < dim) ? __ldcs(x + row * dim + e) : 0.f;
    sum += v[k];
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / static_cast<float>(dim);
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = sub + k * LPR;
    const float d = e < dim ? v[k] - mean : 0.f;
    sq = fmaf(d, d, sq);
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / static_cast<float>(dim) + eps);
  if (!live) return;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = sub + k * LPR;
    if (e < dim) y[row * dim + e] = (v[k] - mean) * rstd * w[e] + b[e];
  }
  if (sub == 0 && stats) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) with g = dy * w; dw += dy * xhat, db += dy (block partial sums, then atomics)
template <int LPR>
__global__ void __launch_bounds__(256) ln_small_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ stats,
                                                           const float* __restrict__ dy, float* __restrict__ dx, float* __restrict__ dw,
                                                           float* __restrict__ db, long long rows, int dim, int rows_per_warp_iter) {
  constexpr int RPW = 32 / LPR;
  __shared__ float s_dw[64], s_db[64];
  if (threadIdx.x < 64) { s_dw[threadIdx.x] = 0.f; s_db[threadIdx.x] = 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const long long warp = (static_cast<long long>(blockIdx.x) * 256 + threadIdx.x) >> 5;
  float aw[4] = {0.f, 0.f, 0.f, 0.f}, ab[4] = {0.f, 0.f, 0.f, 0.f};
  for (int it = 0; it < rows_per_warp_iter; ++it) {
    const long long row = (warp * rows_per_warp_iter + it) * RPW + lane / LPR;
    const bool live = row < rows;
    const float mean = live ? stats[2 * row] : 0.f, rstd = live ? stats[2 * row + 1] : 0.f;
    float xh[4], g[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int e = sub + k * LPR;
      const bool ok = live && e < dim;
      const float xv = ok ? __ldcs(x + row * dim + e) : 0.f, d = ok ? __ldcs(dy + row * dim + e) : 0.f;
      xh[k] = ok ? (xv - mean) * rstd : 0.f;
      g[k] = ok ? d * w[e] : 0.f;
      s1 += g[k];
      s2 = fmaf(g[k], xh[k], s2);
      aw[k] = fmaf(d, xh[k], aw[k]);
      ab[k] += d;
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    const float m1 = s1 / static_cast<float>(dim), m2 = s2 / static_cast<float>(dim);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int e = sub + k * LPR;
      if (live && e < dim) dx[row * dim + e] = rstd * (g[k] - m1 - xh[k] * m2);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = sub + k * LPR;
    if (e < dim) { atomicAdd(&s_dw[e], aw[k]); atomicAdd(&s_db[e], ab[k]); }
  }
  __syncthreads();
  if (threadIdx.x < dim) { atomicAdd(dw + threadIdx.x, s_dw[threadIdx.x]); atomicAdd(db + threadIdx.x, s_db[threadIdx.x]); }
}

}  // namespace catanb

static int pol_fail(cudaError_t e, const char* what) {
  char buf[256];
  snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  catan_set_last_error(buf);
  return -1;
}
extern "C" int catan_tile_attention_fwd(const float* qkv_dev, float* y_dev, int B, void* stream) {
  if (!qkv_dev || !y_dev || B <= 0) return pol_fail(cudaErrorInvalidValue, "catan_tile_attention_fwd: bad argument");
  if ((reinterpret_cast<uintptr_t>(qkv_dev) | reinterpret_cast<uintptr_t>(y_dev)) & 15) return pol_fail(cudaErrorInvalidValue, "catan_tile_attention_fwd: buffers must be 16-byte aligned");
  catanb::DeviceScope device_scope_;                                 // the kernel runs where its buffers live; the caller's device comes back
  if (device_scope_.enter_for(qkv_dev, static_cast<cudaStream_t>(stream))) return pol_fail(cudaErrorInvalidDevice, "cannot make the device of the buffers current");
  cudaError_t e0 = cudaFuncSetAttribute(catanb::tile_attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(catanb::AttSmemF)));
  if (e0 != cudaSuccess) return pol_fail(e0, "catan_tile_attention_fwd");
  const int blocks = (B + catanb::kAttSamples - 1) / catanb::kAttSamples;
  catanb::tile_attention_fwd_kernel<<<blocks, catanb::kAttThreads, sizeof(catanb::AttSmemF), static_cast<cudaStream_t>(stream)>>>(qkv_dev, y_dev, B);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : pol_fail(e, "catan_tile_attention_fwd launch");
}

extern "C" int catan_tile_attention_bwd(const float* qkv_dev, const float* dy_dev, float* dqkv_dev, int B, void* stream) {
  if (!qkv_dev || !dy_dev || !dqkv_dev || B <= 0) return pol_fail(cudaErrorInvalidValue, "catan_tile_attention_bwd: bad argument");
  if ((reinterpret_cast<uintptr_t>(qkv_dev) | reinterpret_cast<uintptr_t>(dy_dev) | reinterpret_cast<uintptr_t>(dqkv_dev)) & 15)
    return pol_fail(cudaErrorInvalidValue, "catan_tile_attention_bwd: buffers must be 16-byte aligned");
  catanb::DeviceScope device_scope_;                                 // the kernel runs where its buffers live; the caller's device comes back
  if (device_scope_.enter_for(qkv_dev, static_cast<cudaStream_t>(stream))) return pol_fail(cudaErrorInvalidDevice, "cannot make the device of the buffers current");
  cudaError_t e = cudaFuncSetAttribute(catanb::tile_attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(catanb::AttSmemB)));
  if (e != cudaSuccess) return pol_fail(e, "catan_tile_attention_bwd");
  const int blocks = (B + catanb::kAttSamplesB - 1) / catanb::kAttSamplesB;
  catanb::tile_attention_bwd_kernel<<<blocks, catanb::kAttThreads, sizeof(catanb::AttSmemB), static_cast<cudaStream_t>(stream)>>>(qkv_dev, dy_dev, dqkv_dev, B);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : pol_fail(e, "catan_tile_attention_bwd launch");
}

template <int LPR>
static int launch_ln_fwd(const float* x, const float* w, const float* b, float* y, float* stats, long long rows, int dim, float eps, cudaStream_t s) {
  const long long warps = (rows + (32 / LPR) - 1) / (32 / LPR);
  const long long blocks = (warps + 7) / 8;
  catanb::ln_small_fwd_kernel<LPR><<<static_cast<unsigned>(blocks), 256, 0, s>>>(x, w, b, y, stats, rows, dim, eps);
  return 0;
}
template <int LPR>
static int launch_ln_bwd(const float* x, const float* w, const float* stats, const float* dy, float* dx, float* dw, float* db, long long rows, int dim,
                         cudaStream_t s) {
  const int iters = 8;                                               // rows per warp lane group: fewer atomics on dw / db
  const long long warps = (rows + (32 / LPR) * iters - 1) / ((32 / LPR) * iters);
  const long long blocks = (warps + 7) / 8;
  catanb::ln_small_bwd_kernel<LPR><<<static_cast<unsigned>(blocks), 256, 0, s>>>(x, w, stats, dy, dx, dw, db, rows, dim, iters);
  return 0;
}

extern "C" int catan_ln_small_fwd(const float* x_dev, const float* weight_dev, const float* bias_dev, float* y_dev, float* stats_dev, long long rows,
                                  int dim, float eps, void* stream) {
  if (!x_dev || !weight_dev || !bias_dev || !y_dev || rows <= 0 || dim <= 0 || dim > 64) return pol_fail(cudaErrorInvalidValue, "catan_ln_small_fwd: bad argument (dim <= 64)");
  catanb::DeviceScope device_scope_;                                 // the kernel runs where its buffers live; the caller's device comes back
  if (device_scope_.enter_for(x_dev, static_cast<cudaStream_t>(stream))) return pol_fail(cudaErrorInvalidDevice, "cannot make the device of the buffers current");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dim <= 16) launch_ln_fwd<4>(x_dev, weight_dev, bias_dev, y_dev, stats_dev, rows, dim, eps, s);
  else if (dim <= 32) launch_ln_fwd<8>(x_dev, weight_dev, bias_dev, y_dev, stats_dev, rows, dim, eps, s);
  else launch_ln_fwd<16>(x_dev, weight_dev, bias_dev, y_dev, stats_dev, rows, dim, eps, s);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : pol_fail(e, "catan_ln_small_fwd launch");
}

extern "C" int catan_ln_small_bwd(const float* x_dev, const float* weight_dev, const float* stats_dev, const float* dy_dev, float* dx_dev,
                                  float* dweight_dev, float* dbias_dev, long long rows, int dim, void* stream) {
  if (!x_dev || !weight_dev || !stats_dev || !dy_dev || !dx_dev || !dweight_dev || !dbias_dev || rows <= 0 || dim <= 0 || dim > 64)
    return pol_fail(cudaErrorInvalidValue, "catan_ln_small_bwd: bad argument (dim <= 64)");
  catanb::DeviceScope device_scope_;                                 // the kernel runs where its buffers live; the caller's device comes back
  if (device_scope_.enter_for(x_dev, static_cast<cudaStream_t>(stream))) return pol_fail(cudaErrorInvalidDevice, "cannot make the device of the buffers current");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dim <= 16) launch_ln_bwd<4>(x_dev, weight_dev, stats_dev, dy_dev, dx_dev, dweight_dev, dbias_dev, rows, dim, s);
  else if (dim <= 32) launch_ln_bwd<8>(x_dev, weight_dev, stats_dev, dy_dev, dx_dev, dweight_dev, dbias_dev, rows, dim, s);
  else launch_ln_bwd<16>(x_dev, weight_dev, stats_dev, dy_dev, dx_dev, dweight_dev, dbias_dev, rows, dim, s);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : pol_fail(e, "catan_ln_small_bwd launch");
}
