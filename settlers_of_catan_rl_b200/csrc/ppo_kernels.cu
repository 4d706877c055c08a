// ppo_kernels.cu — rollout-side PPO kernels (RL/ppo/process_batch.py:134-142): GAE reverse scan and
// advantage normalisation.  HBM-bound streaming kernels; arithmetic is plain IEEE fp32 with explicit
// round-to-nearest intrinsics so that no FMA contraction changes the reference's results.
#include <cuda_runtime.h>

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

}  // namespace catanb

static thread_local std::string g_ppo_error;
extern "C" const char* catan_last_error(void);
static int ppo_fail(cudaError_t e, const char* what);

extern "C" {

int catan_gae(const float* rewards_dev, const float* values_dev, const float* masks_dev, int T, int N, double gamma,
              double gae_lambda, float* returns_dev, float* advantages_dev, void* stream) {
  if (!rewards_dev || !values_dev || !masks_dev || !returns_dev || !advantages_dev || T <= 0 || N <= 0) return ppo_fail(cudaErrorInvalidValue, "catan_gae: bad argument");
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
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(stats_dev, 0, 3 * sizeof(double), s);
  if (e != cudaSuccess) return ppo_fail(e, "catan_adv_stats memset");
  catanb::adv_stats_kernel<<<stream_grid(count), 256, 0, s>>>(advantages_dev, count, stats_dev);
  e = cudaGetLastError();
  return e == cudaSuccess ? 0 : ppo_fail(e, "catan_adv_stats launch");
}

int catan_adv_apply(float* advantages_dev, long long count, const double* stats_dev, double eps, void* stream) {
  if (!advantages_dev || !stats_dev || count <= 1) return ppo_fail(cudaErrorInvalidValue, "catan_adv_apply: bad argument");
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
