// Train-mode BatchNorm2d (batch statistics) forward / backward and the average-pool gradient of the
// `basic` architecture's training step.  All HBM-bound: 16-byte vectors (8 bf16 channels per thread),
// a row of C channels is covered by C/8 consecutive threads, so a warp reads whole 128-byte lines;
// per-thread fp32 partial sums over a bounded number of rows, combined per block in shared memory and
// accumulated across blocks with fp64 atomics (the variance is formed from fp64 sums: no cancellation).
#include "ptx.cuh"
#include "vdqn_internal.h"

#include <cuda_bf16.h>

namespace vdqn {

struct alignas(16) BF8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void unpack8(const BF8& b, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(b.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

__device__ __forceinline__ BF8 pack8(const float* f) {
  BF8 b;
#pragma unroll
  for (int i = 0; i < 4; ++i) b.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return b;
}

constexpr int kBnThreads = 256;

// block partials: red[row_lane][C] for two quantities -> thread t < C sums over row lanes -> fp64 atomics
template <bool BWD>
__global__ void __launch_bounds__(kBnThreads)
bn_reduce_kernel(const BF8* __restrict__ a, const BF8* __restrict__ x, const float* __restrict__ mean,
                 const float* __restrict__ rstd, double* __restrict__ sums, long M, int C) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float red[];                     // [2][rows_per_iter][C]
  const int cv = C >> 3;                             // vectors per row
  const int rpi = kBnThreads / cv;                   // rows per block iteration
  const int lane_c = threadIdx.x % cv, lane_r = threadIdx.x / cv;
  float s0[8], s1[8], mu[8], rs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s0[i] = s1[i] = 0.f;
    mu[i] = 0.f; rs[i] = 1.f;
  }
  if (BWD) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mu[i] = mean[lane_c * 8 + i];
      rs[i] = rstd[lane_c * 8 + i];
    }
  }
  if (lane_r < rpi) {
    for (long m = (long)blockIdx.x * rpi + lane_r; m < M; m += (long)gridDim.x * rpi) {
      float fa[8];
      unpack8(a[m * cv + lane_c], fa);
      if (BWD) {
        float fx[8];
        unpack8(x[m * cv + lane_c], fx);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s0[i] += fa[i];
          s1[i] += fa[i] * ((fx[i] - mu[i]) * rs[i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s0[i] += fa[i];
          s1[i] += fa[i] * fa[i];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red[lane_r * C + lane_c * 8 + i] = s0[i];
      red[(rpi + lane_r) * C + lane_c * 8 + i] = s1[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += kBnThreads) {
    const int q = c / C, ch = c - q * C;
    float t = 0.f;
    for (int r = 0; r < rpi; ++r) t += red[(q * rpi + r) * C + ch];
    atomicAdd(sums + c, (double)t);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, long M, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long long* __restrict__ nbt, float momentum,
                                   float eps, float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ scale, float* __restrict__ shift) {
  pdl_launch_dependents();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt != nullptr) *nbt += 1;
  if (c >= C) return;
  const double mu = sums[c] / (double)M;
  double var = sums[C + c] / (double)M - mu * mu;
  if (var < 0.0) var = 0.0;
  const float r = (float)(1.0 / sqrt(var + (double)eps));
  mean[c] = (float)mu;
  rstd[c] = r;
  const float g = gamma[c];
  scale[c] = g * r;
  shift[c] = beta[c] - (float)mu * (g * r);
  if (running_mean != nullptr) {
    const double unbiased = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

__global__ void __launch_bounds__(kBnThreads)
bn_apply_kernel(const BF8* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                const BF8* __restrict__ residual, int relu, BF8* __restrict__ y, long total_vec, int cv) {
  pdl_launch_dependents();
  pdl_wait();
  // blockDim (256) is a multiple of cv, so a thread keeps the same 8 channels over the whole grid-stride loop
  const int c0 = (int)(threadIdx.x % cv) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = scale[c0 + k];
    sh[k] = shift[c0 + k];
  }
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total_vec; i += (long)gridDim.x * blockDim.x) {
    float f[8], r[8];
    unpack8(x[i], f);
    if (residual != nullptr) unpack8(residual[i], r);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = fmaf(f[k], sc[k], sh[k]);
      if (residual != nullptr) v += r[k];
      if (relu) v = fmaxf(v, 0.f);
      f[k] = v;
    }
    y[i] = pack8(f);
  }
}

__global__ void __launch_bounds__(kBnThreads)
bn_bwd_apply_kernel(const BF8* __restrict__ dy, const BF8* __restrict__ x, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ gamma, const double* __restrict__ sums,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, BF8* __restrict__ dx, long M, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int cv = C >> 3;
  const long total_vec = M * cv;
  const double invM = 1.0 / (double)M;
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      dbeta[c] = (float)sums[c];
      dgamma[c] = (float)sums[C + c];
    }
  }
  const int c0 = (int)(threadIdx.x % cv) * 8;          // fixed per thread (256 % cv == 0)
  float mu[8], rs[8], gr[8], k1[8], k2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c0 + k;
    mu[k] = mean[c];
    rs[k] = rstd[c];
    gr[k] = gamma[c] * rs[k];
    k1[k] = (float)(sums[c] * invM);
    k2[k] = (float)(sums[C + c] * invM);
  }
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total_vec; i += (long)gridDim.x * blockDim.x) {
    float g[8], f[8];
    unpack8(dy[i], g);
    unpack8(x[i], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xhat = (f[k] - mu[k]) * rs[k];
      g[k] = gr[k] * (g[k] - k1[k] - xhat * k2[k]);
    }
    dx[i] = pack8(g);
  }
}

__global__ void avgpool_bwd_kernel(const float* __restrict__ dpooled, const __nv_bfloat16* __restrict__ feat,
                                   __nv_bfloat16* __restrict__ dfeat, int N, int P, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = (long)N * P * C;
  const float invP = 1.f / (float)P;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long n = i / ((long)P * C);
    const float v = __bfloat162float(feat[i]) > 0.f ? dpooled[n * C + c] * invP : 0.f;
    dfeat[i] = __float2bfloat16_rn(v);
  }
}

static inline int bn_grid(long work_items, int per_block, int num_sms) {
  long g = (work_items + per_block - 1) / per_block;
  const long cap = (long)num_sms * 8;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

static int bn_shape_check(const char* what, long M, int C) {
  if (C < 8 || C % 8 != 0 || C > 2048 || kBnThreads % (C / 8) != 0)
    return set_error(VDQN_ERR_SHAPE, "%s: C=%d unsupported (multiple of 8 dividing 2048)", what, C);
  if (M < 0) return set_error(VDQN_ERR_SHAPE, "%s: bad shape", what);
  return VDQN_OK;
}

template <bool BWD>
static int launch_reduce(const char* what, const void* a, const void* x, const float* mean, const float* rstd,
                         double* sums, long M, int C, cudaStream_t stream) {
  DeviceInfo* dev = device_info();
  if (dev == nullptr) return VDQN_ERR_CUDA;
  cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)C, stream);
  if (e != cudaSuccess) return set_error(VDQN_ERR_CUDA, "%s memset: %s", what, cudaGetErrorString(e));
  if (M == 0) return VDQN_OK;
  const int cv = C / 8, rpi = kBnThreads / cv;
  const size_t smem = sizeof(float) * 2 * (size_t)rpi * C;          // = 2 * 256 * 8 * 4 = 16 KB
  launch_kernel(bn_reduce_kernel<BWD>, bn_grid(M, rpi * 8, dev->num_sms), kBnThreads, smem, stream,
                static_cast<const BF8*>(a), static_cast<const BF8*>(x), mean, rstd, sums, M, C);
  ::vdqn::count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(VDQN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return VDQN_OK;
}

}  // namespace vdqn

using namespace vdqn;

extern "C" int vdqn_bn_stats(const void* x, double* sums, int64_t M, int32_t C, void* stream_v) {
  if (sums == nullptr || (x == nullptr && M > 0)) return set_error(VDQN_ERR_ARG, "bn_stats: null pointer");
  int rc = bn_shape_check("bn_stats", M, C);
  if (rc != VDQN_OK) return rc;
  return launch_reduce<false>("bn_stats", x, nullptr, nullptr, nullptr, sums, M, C,
                              static_cast<cudaStream_t>(stream_v));
}

extern "C" int vdqn_bn_bwd_reduce(const void* dy, const void* x, const float* mean, const float* rstd, double* sums,
                                  int64_t M, int32_t C, void* stream_v) {
  if (sums == nullptr || mean == nullptr || rstd == nullptr || ((dy == nullptr || x == nullptr) && M > 0))
    return set_error(VDQN_ERR_ARG, "bn_bwd_reduce: null pointer");
  int rc = bn_shape_check("bn_bwd_reduce", M, C);
  if (rc != VDQN_OK) return rc;
  return launch_reduce<true>("bn_bwd_reduce", dy, x, mean, rstd, sums, M, C, static_cast<cudaStream_t>(stream_v));
}

extern "C" int vdqn_bn_finalize(const double* sums, int64_t M, int32_t C, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, int64_t* num_batches_tracked,
                                float momentum, float eps, float* mean, float* rstd, float* scale, float* shift,
                                void* stream_v) {
  if (sums == nullptr || gamma == nullptr || beta == nullptr || mean == nullptr || rstd == nullptr ||
      scale == nullptr || shift == nullptr)
    return set_error(VDQN_ERR_ARG, "bn_finalize: null pointer");
  if ((running_mean == nullptr) != (running_var == nullptr))
    return set_error(VDQN_ERR_ARG, "bn_finalize: running_mean and running_var go together");
  if (M < 1 || C < 1) return set_error(VDQN_ERR_SHAPE, "bn_finalize: bad shape");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  launch_kernel(bn_finalize_kernel, (C + 127) / 128, 128, 0, stream, sums, (long)M, C, gamma, beta, running_mean,
                running_var, reinterpret_cast<long long*>(num_batches_tracked), momentum, eps, mean, rstd, scale,
                shift);
  VDQN_CHECK_LAUNCH("bn_finalize");
  return VDQN_OK;
}

extern "C" int vdqn_bn_apply(const void* x, const float* scale, const float* shift, const void* residual,
                             int32_t relu, void* y, int64_t M, int32_t C, void* stream_v) {
  if (scale == nullptr || shift == nullptr || ((x == nullptr || y == nullptr) && M > 0))
    return set_error(VDQN_ERR_ARG, "bn_apply: null pointer");
  int rc = bn_shape_check("bn_apply", M, C);
  if (rc != VDQN_OK) return rc;
  DeviceInfo* dev = device_info();
  if (dev == nullptr) return VDQN_ERR_CUDA;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (M == 0) return VDQN_OK;
  const long total_vec = (long)M * (C / 8);
  launch_kernel(bn_apply_kernel, bn_grid(total_vec, kBnThreads * 4, dev->num_sms), kBnThreads, 0, stream,
                static_cast<const BF8*>(x), scale, shift, static_cast<const BF8*>(residual), relu,
                static_cast<BF8*>(y), total_vec, C / 8);
  VDQN_CHECK_LAUNCH("bn_apply");
  return VDQN_OK;
}

extern "C" int vdqn_bn_bwd_apply(const void* dy, const void* x, const float* mean, const float* rstd,
                                 const float* gamma, const double* sums, float* dgamma, float* dbeta, void* dx,
                                 int64_t M, int32_t C, void* stream_v) {
  if (mean == nullptr || rstd == nullptr || gamma == nullptr || sums == nullptr || dgamma == nullptr ||
      dbeta == nullptr || ((dy == nullptr || x == nullptr || dx == nullptr) && M > 0))
    return set_error(VDQN_ERR_ARG, "bn_bwd_apply: null pointer");
  int rc = bn_shape_check("bn_bwd_apply", M, C);
  if (rc != VDQN_OK) return rc;
  if (M < 1) return set_error(VDQN_ERR_SHAPE, "bn_bwd_apply: empty batch");
  DeviceInfo* dev = device_info();
  if (dev == nullptr) return VDQN_ERR_CUDA;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const long total_vec = (long)M * (C / 8);
  launch_kernel(bn_bwd_apply_kernel, bn_grid(total_vec, kBnThreads * 4, dev->num_sms), kBnThreads, 0, stream,
                static_cast<const BF8*>(dy), static_cast<const BF8*>(x), mean, rstd, gamma, sums, dgamma, dbeta,
                static_cast<BF8*>(dx), (long)M, C);
  VDQN_CHECK_LAUNCH("bn_bwd_apply");
  return VDQN_OK;
}

extern "C" int vdqn_avgpool_bwd(const float* dpooled, const void* feat, void* dfeat, int32_t N, int32_t P,
                                int32_t C, void* stream_v) {
  if (N < 0 || P < 1 || C < 1) return set_error(VDQN_ERR_SHAPE, "avgpool_bwd: bad shape");
  if (N == 0) return VDQN_OK;
  if (dpooled == nullptr || feat == nullptr || dfeat == nullptr)
    return set_error(VDQN_ERR_ARG, "avgpool_bwd: null pointer");
  DeviceInfo* dev = device_info();
  if (dev == nullptr) return VDQN_ERR_CUDA;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  const long total = (long)N * P * C;
  launch_kernel(avgpool_bwd_kernel, bn_grid(total, 256 * 4, dev->num_sms), 256, 0, stream, dpooled,
                static_cast<const __nv_bfloat16*>(feat), static_cast<__nv_bfloat16*>(dfeat), N, P, C);
  VDQN_CHECK_LAUNCH("avgpool_bwd");
  return VDQN_OK;
}
