// HBM-bound kernels of the Q-learning step: weight preparation (BN fold + bf16 cast + layout),
// weight-gradient finalisation, input packing, max-pool, the fp32 Q-head MLP, the fused TD
// epilogue and the fused Adam + target-sync update.  All are plain coalesced / vectorised
// grid-stride kernels; none is GEMM-shaped except the tiny fp32 MLP (0.95 MMAC per frame).
#include "ptx.cuh"
#include "vdqn_internal.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdint>
#include <cstdlib>

namespace vdqn {

// ------------------------------------------------------------------------------------------
// GEMM-K index -> OIHW source index.  kmap 0: k = (r*S + s)*Cin + ci.
// kmap 1 (stem, space-to-depth): k = (a*4 + b)*16 + c16, c16 = (ph*2 + pw)*3 + c (12..15 pad),
// original tap r = 2a + ph - 1, s = 2b + pw - 1 on the 7x7 filter (r or s == -1 -> zero).
// Returns -1 for a structural zero.
__device__ __forceinline__ int oihw_index(int co, int k, int Cin, int R, int S, int kmap, int* r_out,
                                          int* s_out, int* ci_out) {
  if (kmap == 0) {
    const int tap = k / Cin, ci = k - tap * Cin;
    const int r = tap / S, s = tap - r * S;
    *r_out = r; *s_out = s; *ci_out = ci;
    return ((co * Cin + ci) * R + r) * S + s;
  }
  const int tap = k >> 4, c16 = k & 15;
  if (c16 >= 12) return -1;
  const int a = tap >> 2, b = tap & 3;
  const int par = c16 / 3, c = c16 - par * 3;
  const int r = 2 * a + (par >> 1) - 1, s = 2 * b + (par & 1) - 1;
  if (r < 0 || s < 0) return -1;
  *r_out = r; *s_out = s; *ci_out = c;
  return ((co * 3 + c) * 7 + r) * 7 + s;
}

__device__ __forceinline__ void weight_prep_element(const vdqn_wprep_desc& d, long i) {
  const int co = (int)(i / d.K), k = (int)(i - (long)co * d.K);
  float scale = 1.f;
  if (d.gamma != nullptr) scale = d.gamma[co] * (1.0f / sqrtf(d.var[co] + d.eps));
  int r = 0, s = 0, ci = 0;
  const int src = oihw_index(co, k, d.Cin, d.R, d.S, d.kmap, &r, &s, &ci);
  const float v = src >= 0 ? d.w[src] * scale : 0.f;
  const __nv_bfloat16 b = __float2bfloat16_rn(v);
  static_cast<__nv_bfloat16*>(d.w_fwd)[i] = b;
  if (d.w_dgrad != nullptr && d.kmap == 0) {
    long di;
    if (d.dgrad_parity) {
      // output-parity class (a, b) and tap (u, v) inside it: r = 1 -> (a=0,u=0); r = 2 -> (1,0); r = 0 -> (1,1)
      const int pa = (r == 1) ? 0 : 1, u = (r == 0) ? 1 : 0;
      const int pb = (s == 1) ? 0 : 1, v = (s == 0) ? 1 : 0;
      const int nb = 1 + pb, nt = (1 + pa) * nb;
      const int cls_off = (pa == 0) ? (pb == 0 ? 0 : 1) : (pb == 0 ? 3 : 5);
      di = (long)cls_off * d.Cin * d.Cout + (long)ci * (nt * d.Cout) + (long)(u * nb + v) * d.Cout + co;
    } else {
      di = (long)ci * (d.R * d.S * d.Cout) + (long)((d.R - 1 - r) * d.S + (d.S - 1 - s)) * d.Cout + co;
    }
    static_cast<__nv_bfloat16*>(d.w_dgrad)[di] = b;
  }
  if (k == 0) {
    float sh = 0.f;
    if (d.gamma != nullptr) sh = d.beta[co] - d.mean[co] * scale;
    if (d.bias != nullptr) sh += d.bias[co];
    d.shift[co] = sh;
  }
}

__global__ void weight_prep_kernel(const vdqn_wprep_desc d) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = (long)d.Cout * d.K;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
    weight_prep_element(d, i);
}

// all convolutions of the network in one launch: `offsets[t]` = first flat element of tensor t
__global__ void weight_prep_multi_kernel(const vdqn_wprep_desc* __restrict__ descs,
                                         const long long* __restrict__ offsets, int n, long total) {
  pdl_launch_dependents();
  pdl_wait();
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (offsets[mid] <= i) lo = mid; else hi = mid - 1;
    }
    weight_prep_element(descs[lo], i - offsets[lo]);
  }
}

// Tiled variant for the regular (r,s,ci) layout: one block = 32 output x 32 input channels x all taps
// of one tensor.  OIHW is read as contiguous runs of 32*R*S floats per output channel (four
// independent loads in flight per thread), transposed in shared memory, and both bf16 layouts are
// written as bf16x2 pairs with the fastest index contiguous.  `tile_offsets[t]` = first block of
// tensor t.  RS_T = R*S as a compile-time constant (0: generic).
template <int RS_T>
__device__ __forceinline__ void weight_prep_tile(const vdqn_wprep_desc& d, int co0, int ci0,
                                                 float (*sw)[32 * 9 + 1], const float* sscale) {
  const int RS = RS_T ? RS_T : d.R * d.S, run = 32 * RS;
  const int tid = threadIdx.x;
  const float* wbase = d.w + ((long)co0 * d.Cin + ci0) * RS;
  const long co_pitch = (long)d.Cin * RS;
  // eight independent loads in flight per thread (a block is 9 passes of four otherwise, each a full memory
  // latency: the kernel ran at a third of the HBM rate)
  for (int e0 = 0; e0 < 32 * run; e0 += 2048) {               // 32*run is a multiple of 1024
    float v[8];
    int col[8], j[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + u * 256 + tid;
      col[u] = e / run; j[u] = e - col[u] * run;                // j = cil*RS + tap, contiguous in OIHW
      v[u] = e < 32 * run ? __ldg(wbase + col[u] * co_pitch + j[u]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (e0 + u * 256 + tid < 32 * run) sw[col[u]][j[u]] = v[u] * sscale[col[u]];
  }
  __syncthreads();
  __nv_bfloat162* wf = reinterpret_cast<__nv_bfloat162*>(d.w_fwd);
  for (int e = tid; e < 16 * run; e += 256) {                  // forward: [co][tap][ci], ci fastest
    const int cil = (e & 15) * 2, rest = e >> 4;
    const int c = rest / RS, tap = rest - c * RS;
    const long o = (long)(co0 + c) * (d.ldw_fwd > 0 ? d.ldw_fwd : d.K) + d.fwd_col0 + (long)tap * d.Cin + ci0 + cil;
    wf[o >> 1] = __floats2bfloat162_rn(sw[c][cil * RS + tap], sw[c][(cil + 1) * RS + tap]);
  }
  if (d.w_dgrad != nullptr) {
    __nv_bfloat162* wd = reinterpret_cast<__nv_bfloat162*>(d.w_dgrad);
    for (int e = tid; e < 16 * run; e += 256) {                // data gradient: co fastest
      const int c = (e & 15) * 2, rest = e >> 4;
      const int cil = rest / RS, tap = rest - cil * RS;
      const int r = tap / d.S, sx = tap - r * d.S;
      const int ci = ci0 + cil, co = co0 + c;
      long di;
      if (d.dgrad_parity == 2) {
        // one dense matrix for all four output-parity classes: row (a, b, ci), column (u, v, co)
        const int pa = (r == 1) ? 0 : 1, u = (r == 0) ? 1 : 0;
        const int pb = (sx == 1) ? 0 : 1, v = (sx == 0) ? 1 : 0;
        const long ld = d.ldw_dgrad > 0 ? d.ldw_dgrad : 4 * d.Cout;
        di = ((long)(pa * 2 + pb) * d.Cin + ci) * ld + (long)(u * 2 + v) * d.Cout + co;
      } else if (d.dgrad_parity) {
        const int pa = (r == 1) ? 0 : 1, u = (r == 0) ? 1 : 0;
        const int pb = (sx == 1) ? 0 : 1, v = (sx == 0) ? 1 : 0;
        const int nb = 1 + pb, nt = (1 + pa) * nb;
        const int cls_off = (pa == 0) ? (pb == 0 ? 0 : 1) : (pb == 0 ? 3 : 5);
        di = (long)cls_off * d.Cin * d.Cout + (long)ci * (nt * d.Cout) + (long)(u * nb + v) * d.Cout + co;
      } else {
        di = (long)ci * (d.ldw_dgrad > 0 ? d.ldw_dgrad : RS * d.Cout) + d.dgrad_col0 +
             (long)((d.R - 1 - r) * d.S + (d.S - 1 - sx)) * d.Cout + co;
      }
      wd[di >> 1] = __floats2bfloat162_rn(sw[c][cil * RS + tap], sw[c + 1][cil * RS + tap]);
    }
  }
}

__global__ void __launch_bounds__(256)
weight_prep_tiled_kernel(const vdqn_wprep_desc* __restrict__ descs, const int* __restrict__ tile_offsets, int n) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sw[32][32 * 9 + 1];
  __shared__ float sscale[32];
  // which tensor: up to 32 table entries are searched with one parallel load and a ballot
  __shared__ int s_item;
  int lo = 0;
  if (n <= 32) {
    if (threadIdx.x < 32) {
      const int fb = (int)threadIdx.x < n ? tile_offsets[threadIdx.x] : 0x7fffffff;
      const unsigned m = __ballot_sync(0xffffffffu, fb <= (int)blockIdx.x);
      if (threadIdx.x == 0) s_item = 31 - __clz(m);
    }
    __syncthreads();
    lo = s_item;
  } else {
    int hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (tile_offsets[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
  }
  const vdqn_wprep_desc d = descs[lo];
  const int b = blockIdx.x - tile_offsets[lo];
  const int ci_tiles = d.Cin / 32;
  const int co0 = (b / ci_tiles) * 32, ci0 = (b % ci_tiles) * 32;
  if (threadIdx.x < 32) {
    const int co = co0 + threadIdx.x;
    float scale = 1.f;
    if (d.gamma != nullptr) scale = d.gamma[co] * (1.0f / sqrtf(d.var[co] + d.eps));
    sscale[threadIdx.x] = scale;
    if (ci0 == 0) {
      float sh = 0.f;
      if (d.gamma != nullptr) sh = d.beta[co] - d.mean[co] * scale;
      if (d.bias != nullptr) sh += d.bias[co];
      if (d.gamma_b != nullptr)      // the shift of a second BatchNorm whose conv is accumulated into this one
        sh += d.beta_b[co] - d.mean_b[co] * (d.gamma_b[co] * (1.0f / sqrtf(d.var_b[co] + d.eps)));
      d.shift[co] = sh;
    }
  }
  __syncthreads();
  const int RS = d.R * d.S;
  if (RS == 9) weight_prep_tile<9>(d, co0, ci0, sw, sscale);
  else if (RS == 1) weight_prep_tile<1>(d, co0, ci0, sw, sscale);
  else weight_prep_tile<0>(d, co0, ci0, sw, sscale);
}

// grid = (K blocks of 64, Cout), block = 64 (k) x 4 (split lanes).  A thread sums the partials of the
// splits s = ty, ty+4, ... for its (co, k) (4 independent loads in flight per iteration), the four
// lanes are combined through shared memory in a fixed order (deterministic), lane 0 writes dW and
// feeds d gamma (one atomic per warp into a zeroed slot).
__global__ void __launch_bounds__(256)
wgrad_finalize_kernel(const vdqn_wgrad_fin_desc d) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[4][64];
  const int co = blockIdx.y;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int k = blockIdx.x * 64 + tx;
  float g = 0.f;
  if (k < d.K) {
    const long plane = (long)d.Cout * d.K;
    const float* p = d.part + (long)co * d.K + k;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f, g3 = 0.f;
    int s = ty;
    for (; s + 12 < d.splits; s += 16) {
      g0 += p[(long)s * plane];
      g1 += p[(long)(s + 4) * plane];
      g2 += p[(long)(s + 8) * plane];
      g3 += p[(long)(s + 12) * plane];
    }
    for (; s < d.splits; s += 4) g0 += p[(long)s * plane];
    g = (g0 + g1) + (g2 + g3);
  }
  red[ty][tx] = g;
  __syncthreads();
  if (ty != 0) return;
  g = (red[0][tx] + red[1][tx]) + (red[2][tx] + red[3][tx]);
  float rstd = 1.f, scale = 1.f;
  if (d.gamma != nullptr) {
    rstd = 1.0f / sqrtf(d.var[co] + d.eps);
    scale = d.gamma[co] * rstd;
  }
  float dot = 0.f;
  if (k < d.K) {
    int r, ss, ci;
    const int src = oihw_index(co, k, d.Cin, d.R, d.S, d.kmap, &r, &ss, &ci);
    if (src >= 0) {
      dot = d.w[src] * g;
      d.dw[src] = scale * g;
    }
  }
  if (d.dgamma != nullptr) {
    for (int off = 16; off; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
    if ((tx & 31) == 0) {
      float v = rstd * dot;
      if (blockIdx.x == 0 && tx == 0 && d.dbeta != nullptr) v -= rstd * d.mean[co] * d.dbeta[co];
      atomicAdd(d.dgamma + co, v);
    }
  }
}

// Row form of the finalize step for the (tap, ci)-ordered GEMM-K layouts (kmap 0): one block owns
// output channel co = blockIdx.y and input channels [ci0, ci0 + cn).  Phase 1 sums the split
// partials with 16-byte loads -- item = (tap, 4 consecutive ci), TX items x TY split lanes, four
// independent loads in flight per thread -- into shared memory; phase 2 walks the same elements in
// OIHW order (contiguous for a fixed co: o = ci*RS + tap), so W is read and dW written as full
// lines, and folds the d gamma dot product into one atomic per block.  Fixed summation order.
struct FinRowCfg {
  int cn, items, TX, TY;
};

template <int RS_T>
__device__ __forceinline__ void finalize_rows_body_t(const vdqn_wgrad_fin_desc& d, const FinRowCfg& c, int chunk,
                                                     int co, float* red, float* wsum) {
  const int ci0 = chunk * c.cn;
  const int RS = RS_T ? RS_T : d.R * d.S;            // compile-time for 3x3 / 1x1: the divisions below become multiplies
  const int E = c.cn * RS;
  const int pitch = c.cn + 1;                    // per-tap pitch in shared memory (bank spread)
  const int lane_pitch = RS * pitch;
  const int tid = threadIdx.x;
  const int lane = tid / c.TX, tx = tid - lane * c.TX;
  const long plane = (long)d.Cout * d.K;
  const int cn4 = c.cn >> 2;
  if (lane < c.TY && d.splits <= 2 * c.TY) {
    // few splits (layers 3-4: 2 .. 8): one or two loads per item and lane would leave a thread with two loads in
    // flight; take four items at a time so that up to eight are
    for (int item0 = tx; item0 < c.items; item0 += 4 * c.TX) {
      float4 v[4][2];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int item = item0 + u * c.TX;
        v[u][0] = v[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (item < c.items) {
          const int tap = item / cn4, c4 = item - tap * cn4;
          const float* p = d.part + (long)co * d.K + (long)tap * d.Cin + ci0 + 4 * c4;
          v[u][0] = *reinterpret_cast<const float4*>(p + (long)lane * plane);
          if (lane + c.TY < d.splits) v[u][1] = *reinterpret_cast<const float4*>(p + (long)(lane + c.TY) * plane);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int item = item0 + u * c.TX;
        if (item < c.items) {
          const int tap = item / cn4, c4 = item - tap * cn4;
          float* r = red + lane * lane_pitch + tap * pitch + 4 * c4;
          r[0] = v[u][0].x + v[u][1].x;
          r[1] = v[u][0].y + v[u][1].y;
          r[2] = v[u][0].z + v[u][1].z;
          r[3] = v[u][0].w + v[u][1].w;
        }
      }
    }
  } else if (lane < c.TY) {
    for (int item = tx; item < c.items; item += c.TX) {
      const int tap = item / cn4, c4 = item - tap * cn4;
      const float* p = d.part + (long)co * d.K + (long)tap * d.Cin + ci0 + 4 * c4;
      float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
      int s = lane;
      for (; s + 3 * c.TY < d.splits; s += 4 * c.TY) {
        const float4 v0 = *reinterpret_cast<const float4*>(p + (long)s * plane);
        const float4 v1 = *reinterpret_cast<const float4*>(p + (long)(s + c.TY) * plane);
        const float4 v2 = *reinterpret_cast<const float4*>(p + (long)(s + 2 * c.TY) * plane);
        const float4 v3 = *reinterpret_cast<const float4*>(p + (long)(s + 3 * c.TY) * plane);
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
        a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
        a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
        a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
      }
      for (; s < d.splits; s += c.TY) {
        const float4 v0 = *reinterpret_cast<const float4*>(p + (long)s * plane);
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
      }
      float* r = red + lane * lane_pitch + tap * pitch + 4 * c4;
      r[0] = (a0.x + a1.x) + (a2.x + a3.x);
      r[1] = (a0.y + a1.y) + (a2.y + a3.y);
      r[2] = (a0.z + a1.z) + (a2.z + a3.z);
      r[3] = (a0.w + a1.w) + (a2.w + a3.w);
    }
  }
  __syncthreads();
  float rstd = 1.f, scale = 1.f;
  if (d.gamma != nullptr) {
    rstd = 1.0f / sqrtf(d.var[co] + d.eps);
    scale = d.gamma[co] * rstd;
  }
  const long o0 = ((long)co * d.Cin + ci0) * RS;
  float dot = 0.f;
  // four elements per thread and pass, their W loads issued together (one at a time, each pass waited a full
  // memory latency: a quarter of the kernel's stall samples)
  for (int j0 = tid; j0 < E; j0 += 1024) {
    float wv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) wv[u] = (j0 + u * 256 < E) ? __ldg(d.w + o0 + j0 + u * 256) : 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * 256;
      if (j < E) {
        const int cil = j / RS, tap = j - cil * RS;
        const float* r = red + tap * pitch + cil;
        float g = r[0];
        for (int l = 1; l < c.TY; ++l) g += r[l * lane_pitch];
        dot += wv[u] * g;
        d.dw[o0 + j] = scale * g;
      }
    }
  }
  if (d.dgamma != nullptr) {
    for (int off = 16; off; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
    if ((tid & 31) == 0) wsum[tid >> 5] = dot;
    __syncthreads();
    if (tid == 0) {
      float v = 0.f;
      for (int w = 0; w < 8; ++w) v += wsum[w];
      v *= rstd;
      if (chunk == 0 && d.dbeta != nullptr) v -= rstd * d.mean[co] * d.dbeta[co];
      atomicAdd(d.dgamma + co, v);
    }
  }
}

__device__ __forceinline__ void finalize_rows_body(const vdqn_wgrad_fin_desc& d, const FinRowCfg& c, int chunk,
                                                   int co, float* red, float* wsum) {
  const int RS = d.R * d.S;
  if (RS == 9) finalize_rows_body_t<9>(d, c, chunk, co, red, wsum);
  else if (RS == 1) finalize_rows_body_t<1>(d, c, chunk, co, red, wsum);
  else finalize_rows_body_t<0>(d, c, chunk, co, red, wsum);
}

__global__ void __launch_bounds__(256, 5)
wgrad_finalize_rows_kernel(const vdqn_wgrad_fin_desc d, const FinRowCfg c) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[4608 + 160];
  __shared__ float wsum[8];
  finalize_rows_body(d, c, blockIdx.x, blockIdx.y, red, wsum);
}

// Every split reduction of a backward pass (or of one data-parallel stage) in ONE launch: block ->
// (tensor, output channel, input-channel chunk) through a table of per-tensor descriptors that lives
// on the device.  Removes ~20 launch + tail gaps per step and lets the partials of all layers be
// reduced at full HBM rate.
__global__ void __launch_bounds__(256, 5)
wgrad_finalize_multi_kernel(const vdqn_wgrad_fin_item* __restrict__ items, int n) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[4608 + 160];
  __shared__ float wsum[8];
  // which tensor: the last table entry whose first block is <= this block.  Up to 32 entries: one parallel load
  // and a ballot (a binary search is five dependent L2 round trips in front of every block's work)
  __shared__ int s_item;
  int lo = 0;
  if (n <= 32) {
    if (threadIdx.x < 32) {
      const int fb = (int)threadIdx.x < n ? items[threadIdx.x].first_block : 0x7fffffff;
      const unsigned m = __ballot_sync(0xffffffffu, fb <= (int)blockIdx.x);
      if (threadIdx.x == 0) s_item = 31 - __clz(m);
    }
    __syncthreads();
    lo = s_item;
  } else {
    int hi = n - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (items[mid].first_block <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
  }
  const vdqn_wgrad_fin_item it = items[lo];
  const int b = (int)blockIdx.x - it.first_block;
  FinRowCfg c;
  c.cn = it.cn; c.items = it.items; c.TX = it.TX; c.TY = it.TY;
  finalize_rows_body(it.d, c, b % it.nchunks, b / it.nchunks, red, wsum);
}

// ------------------------------------------------------------------------------------------
// input packing: one thread per packed pixel (n, h2, w2) -> 16 bf16 (32 bytes)
__global__ void stem_pack_f32_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int H,
                                     int W) {
  pdl_launch_dependents();
  pdl_wait();
  const int H2 = H / 2, W2 = W / 2;
  const long total = (long)N * H2 * W2;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int w2 = (int)(i % W2);
    const long t = i / W2;
    const int h2 = (int)(t % H2), n = (int)(t / H2);
    float v[16];
#pragma unroll
    for (int j = 12; j < 16; ++j) v[j] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int ph = 0; ph < 2; ++ph) {
        const float2 f = *reinterpret_cast<const float2*>(x + (((long)n * 3 + c) * H + 2 * h2 + ph) * W + 2 * w2);
        v[(ph * 2 + 0) * 3 + c] = f.x;
        v[(ph * 2 + 1) * 3 + c] = f.y;
      }
    uint4 pk[2];
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
    for (int j = 0; j < 8; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    uint4* o = reinterpret_cast<uint4*>(out + i * 16);
    o[0] = pk[0];
    o[1] = pk[1];
  }
}

// uint8 HWC frames: one thread = TWO horizontally adjacent packed pixels = 12 source bytes per image
// row (three aligned 32-bit loads), normalised as util/torch.py:26-36 (x/255, -mean, /std).
// The two fp32 divisions per element made this kernel instruction-bound; it evaluates
// fma(b, 1/(255 std), -mean/std) instead, which rounds to the SAME bf16 as the reference formula
// for every one of the 3 x 256 possible (channel, byte) inputs (exhaustive check:
// tests/test_oracle_golden.py::test_u8_normalisation_fma_is_exact).
__global__ void stem_pack_u8_kernel(const uint8_t* __restrict__ x, __nv_bfloat16* __restrict__ out, int N, int H,
                                    int W) {
  pdl_launch_dependents();
  pdl_wait();
  const int H2 = H / 2, W4 = W / 4;
  const long total = (long)N * H2 * W4;
  const float ka[3] = {(float)(1.0 / (255.0 * (double)0.229f)), (float)(1.0 / (255.0 * (double)0.224f)),
                       (float)(1.0 / (255.0 * (double)0.225f))};
  const float kb[3] = {(float)(-(double)0.485f / (double)0.229f), (float)(-(double)0.456f / (double)0.224f),
                       (float)(-(double)0.406f / (double)0.225f)};
  const int per_img = H2 * W4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int n = (int)(i / per_img);
    const int t = (int)(i - (long)n * per_img);
    const int h2 = t / W4, w4 = t - h2 * W4;
    uint32_t raw[2][3];
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
      const uint32_t* row = reinterpret_cast<const uint32_t*>(x + (((long)n * H + 2 * h2 + ph) * W + 4 * w4) * 3);
#pragma unroll
      for (int j = 0; j < 3; ++j) raw[ph][j] = __ldg(row + j);
    }
#pragma unroll
    for (int px = 0; px < 2; ++px) {          // packed pixel 2*w4 + px <- source pixels 4*w4 + 2*px + {0,1}
      float v[16];
#pragma unroll
      for (int j = 12; j < 16; ++j) v[j] = 0.f;
#pragma unroll
      for (int ph = 0; ph < 2; ++ph)
#pragma unroll
        for (int pw = 0; pw < 2; ++pw)
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int byte = (2 * px + pw) * 3 + c;                 // 0..11 within the 12-byte row chunk
            const uint32_t b = (raw[ph][byte >> 2] >> (8 * (byte & 3))) & 0xffu;
            v[(ph * 2 + pw) * 3 + c] = fmaf((float)b, ka[c], kb[c]);
          }
      uint4 pk[2];
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(pk);
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      uint4* o = reinterpret_cast<uint4*>(out + ((((long)n * H2 + h2) * (W / 2)) + 2 * w4 + px) * 16);
      o[0] = pk[0];
      o[1] = pk[1];
    }
  }
}

// ------------------------------------------------------------------------------------------
// max_pool2d(3,2,1), NHWC bf16; one thread = 8 channels of one output pixel.  Values stay packed
// as bf16x2: one __hgt2_mask + __hmax2 + select per pair and tap.  Strict '>' against a -inf start
// keeps the first maximum of the window scan (torch's choice of arg-max).  All nine 16-byte loads
// are issued up front from clamped (always valid) addresses and out-of-image taps are replaced by
// -inf afterwards: no branch sits between the loads, so nine are in flight per thread (the branchy
// form had one, and ran at half the HBM rate).
template <bool IDX>
__global__ void __launch_bounds__(256, 6) maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                   uint8_t* __restrict__ idx, int N, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1, CG = C / 8;
  const long total = (long)N * Ho * Wo * CG;
  const int per_img = Ho * Wo * CG;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int n = (int)(i / per_img);
    int t = (int)(i - (long)n * per_img);
    const int cg = t % CG; t /= CG;
    const int q = t % Wo;
    const int p = t / Wo;
    const uint4* base = reinterpret_cast<const uint4*>(x + ((long)n * H * W) * C + cg * 8);
    const int h0 = 2 * p - 1, w0 = 2 * q - 1;
    uint4 raw[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int h = min(max(h0 + r, 0), H - 1);
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int w = min(max(w0 + s, 0), W - 1);
        raw[r * 3 + s] = __ldg(base + (h * W + w) * CG);
      }
    }
    uint32_t best[4], slot[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) { best[e] = 0xFF80FF80u; slot[e] = 0u; }     // -inf, -inf
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const bool valid = (unsigned)(h0 + r) < (unsigned)H && (unsigned)(w0 + s) < (unsigned)W;
        const uint4 rv = raw[r * 3 + s];
        const uint32_t v[4] = {valid ? rv.x : 0xFF80FF80u, valid ? rv.y : 0xFF80FF80u,
                               valid ? rv.z : 0xFF80FF80u, valid ? rv.w : 0xFF80FF80u};
        const uint32_t tap2 = (uint32_t)(r * 3 + s) * 0x00010001u;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 nv = *reinterpret_cast<const __nv_bfloat162*>(&v[e]);
          const __nv_bfloat162 bv = *reinterpret_cast<const __nv_bfloat162*>(&best[e]);
          if (IDX) {
            const uint32_t m = __hgt2_mask(nv, bv);
            slot[e] = (tap2 & m) | (slot[e] & ~m);
          }
          const __nv_bfloat162 mx = __hmax2(nv, bv);
          best[e] = *reinterpret_cast<const uint32_t*>(&mx);
        }
      }
    }
    *reinterpret_cast<uint4*>(y + i * 8) = make_uint4(best[0], best[1], best[2], best[3]);
    if (IDX) {
      uint2 ip;     // 16-bit slots -> bytes, channel order preserved
      ip.x = __byte_perm(slot[0], slot[1], 0x6420);
      ip.y = __byte_perm(slot[2], slot[3], 0x6420);
      *reinterpret_cast<uint2*>(idx + i * 8) = ip;
    }
  }
}

// backward: one thread = 8 channels of a 2x2 block of INPUT pixels (rows 2i, 2i+1; cols 2j, 2j+1).
// The block is covered by the four windows (i..i+1, j..j+1), each loaded once (arg-max slots,
// dy, pooled y); a window's gradient goes to the pixel whose slot matches.  The stem ReLU mask is
// taken from the POOLED output y (the arg-max element is > 0 exactly when the window maximum is), so
// the 4x larger pre-pool activation is not re-read.  Accumulates per-channel sums (d beta of bn1).
// A block walks pooled rows (n, i); its threads are the (j, channel group) items of a row, so no thread
// divides per item (ncu of the flat-index version: 590 instructions per item, issue slots 69 % busy at 46 %
// of DRAM bandwidth -- three 32-bit and one 64-bit runtime division and the emulated byte compare
// __vcmpeq4 were two thirds of them).
__global__ void maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx,
                                   const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ dx,
                                   float* __restrict__ colsum, int N, int H, int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2, CG = C / 8;      // H, W even: 3x3/2/1 pooling halves them
  const int rows = N * Ho, items = Wo * CG;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
  // this thread's items of a row: it0, it0 + blockDim.x, ...  (blockDim.x is a multiple of CG: the channel
  // group is the same for all of them)
  const int j0 = (int)threadIdx.x / CG, cg = (int)threadIdx.x - j0 * CG, dj = (int)blockDim.x / CG;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / Ho, i = row - n * Ho;
    const int i1 = min(i + 1, Ho - 1);
    const long in0 = ((long)n * Ho + i) * Wo * C, in1 = ((long)n * Ho + i1) * Wo * C;   // pooled rows i, i+1
    const long out0 = ((long)n * H + 2 * i) * W * C;                                    // input row 2i
    const bool row1_in = i + 1 < Ho;
    for (int j = j0; j < Wo; j += dj) {
      // masked window gradients gw[a][b] for windows (i+a, j+b) and their arg-max slots.  All twelve
      // loads are issued first, from clamped addresses (no branch in between); windows outside the
      // image are neutralised afterwards.
      uint32_t gw[2][2][4];
      uint2 sl[2][2];
      uint4 gr[2][2], yr[2][2];
      const int jj1 = min(j + 1, Wo - 1);
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const long o = (a ? in1 : in0) + (long)((b ? jj1 : j) * C + cg * 8);
          sl[a][b] = __ldg(reinterpret_cast<const uint2*>(idx + o));
          gr[a][b] = __ldg(reinterpret_cast<const uint4*>(dy + o));
          yr[a][b] = __ldg(reinterpret_cast<const uint4*>(y + o));
        }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const bool in = (a == 0 || row1_in) && (b == 0 || j + 1 < Wo);
          const uint32_t gv[4] = {gr[a][b].x, gr[a][b].y, gr[a][b].z, gr[a][b].w};
          const uint32_t yv[4] = {yr[a][b].x, yr[a][b].y, yr[a][b].z, yr[a][b].w};
          if (!in) sl[a][b] = make_uint2(0xffffffffu, 0xffffffffu);      // slot 255 never matches
#pragma unroll
          for (int e = 0; e < 4; ++e)
            gw[a][b][e] = in ? (gv[e] & __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&yv[e]), zero2)) : 0u;
        }
      // slots as 16-bit lanes next to the bf16 pairs they steer: 0x3C00 | slot, i.e. the fp16 numbers
      // 1 + slot / 1024, so that ONE packed fp16 compare per channel pair yields the 0xFFFF / 0 lane mask
      uint32_t s16[2][2][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          s16[a][b][0] = __byte_perm(sl[a][b].x, 0x3C3C3C3Cu, 0x4140);
          s16[a][b][1] = __byte_perm(sl[a][b].x, 0x3C3C3C3Cu, 0x4342);
          s16[a][b][2] = __byte_perm(sl[a][b].y, 0x3C3C3C3Cu, 0x4140);
          s16[a][b][3] = __byte_perm(sl[a][b].y, 0x3C3C3C3Cu, 0x4342);
        }
      // out[ph][pw] for input pixel (2i+ph, 2j+pw): window (i+a, j+b) reaches it through slot
      // r*3+s with r = ph - 2a + 1, s = pw - 2b + 1 (valid when 0 <= r,s <= 2)
#pragma unroll
      for (int ph = 0; ph < 2; ++ph)
#pragma unroll
        for (int pw = 0; pw < 2; ++pw) {
          uint32_t o2[4] = {0u, 0u, 0u, 0u};
          bool first = true;
#pragma unroll
          for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              const int r = ph - 2 * a + 1, sx = pw - 2 * b + 1;
              if (r < 0 || r > 2 || sx < 0 || sx > 2) continue;
              const uint32_t want2 = 0x3C003C00u | ((uint32_t)(r * 3 + sx) * 0x00010001u);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const uint32_t v = gw[a][b][e] & __heq2_mask(*reinterpret_cast<const __half2*>(&s16[a][b][e]),
                                                             *reinterpret_cast<const __half2*>(&want2));
                if (first) {
                  o2[e] = v;
                } else {
                  const __nv_bfloat162 sum = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&o2[e]),
                                                     *reinterpret_cast<const __nv_bfloat162*>(&v));
                  o2[e] = *reinterpret_cast<const uint32_t*>(&sum);
                }
              }
              first = false;
            }
          // column sums of the values as stored (a pixel that is the arg-max of several windows holds their
          // bf16 sum)
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&o2[e]));
            acc[2 * e] += f.x;
            acc[2 * e + 1] += f.y;
          }
          const long op = out0 + (long)((ph * W + 2 * j + pw) * C + cg * 8);
          *reinterpret_cast<uint4*>(dx + op) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
        }
    }
  }
  if (colsum != nullptr) {
    extern __shared__ float sm[];      // [blockDim.x][8]
#pragma unroll
    for (int e = 0; e < 8; ++e) sm[threadIdx.x * 8 + e] = acc[e];
    __syncthreads();
    if ((int)threadIdx.x < C) {
      const int g0 = threadIdx.x / 8, e = threadIdx.x % 8;
      float s = 0.f;
      for (int t = g0; t < (int)blockDim.x; t += CG) s += sm[t * 8 + e];
      atomicAdd(colsum + threadIdx.x, s);
    }
  }
}

// ------------------------------------------------------------------------------------------
// fp32 strided GEMM for the Q-head MLP:  C[m,n] = act(sum_k A(m,k) * B(k,n) + bias[n])
// A(m,k) = A[m*a_rs + k*a_cs], B(k,n) = B[k*b_rs + n*b_cs].  64x64 tile, 16-deep, 256 threads.
__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ Cm,
                     const float* __restrict__ bias, int M, int N, int K, long a_rs, long a_cs,
                     long b_rs, long b_cs, int ldc, int relu, int k_len) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // split-K: blockIdx.z owns K range [k_begin, k_end) and accumulates with atomics
  const int k_begin = blockIdx.z * k_len;
  const int k_end = min(K, k_begin + k_len);
  const bool split = gridDim.z > 1;
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const int e = threadIdx.x + l * 256;       // 0..1023
      // pick the traversal order that is contiguous in memory for each operand
      int am, ak, bk, bn;
      if (a_cs == 1) { ak = e & 15; am = e >> 4; } else { am = e & 63; ak = e >> 6; }
      if (b_cs == 1) { bn = e & 63; bk = e >> 6; } else { bk = e & 15; bn = e >> 4; }
      const int gm = m0 + am, gk = k0 + ak;
      As[ak][am] = (gm < M && gk < k_end) ? A[gm * a_rs + gk * a_cs] : 0.f;
      const int gn = n0 + bn, gk2 = k0 + bk;
      Bs[bk][bn] = (gn < N && gk2 < k_end) ? Bm[gk2 * b_rs + gn * b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      // one 16-byte shared load per operand (row pitch 68 floats keeps ty*4 / tx*4 16-byte aligned)
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (split) {
        atomicAdd(Cm + (long)gm * ldc + gn, v);
      } else {
        if (bias != nullptr) v += bias[gn];
        if (relu) v = fmaxf(v, 0.f);
        Cm[(long)gm * ldc + gn] = v;
      }
    }
  }
}

__global__ void bias_act_kernel(float* __restrict__ c, const float* __restrict__ bias, long total, int N,
                                int relu) {
  pdl_launch_dependents();
  pdl_wait();
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    float v = c[i];
    if (bias != nullptr) v += bias[i % N];
    if (relu) v = fmaxf(v, 0.f);
    c[i] = v;
  }
}

// dyp = dy * (y > 0 if relu); db[o] = sum_b dyp[b,o]   (one block per 32 columns);
// dyp16: optional bf16 copy of dyp (operand of the tensor-core weight / data gradient of top.0)
__global__ void mask_colsum_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                   float* __restrict__ dyp, float* __restrict__ db, int B, int O,
                                   int relu, __nv_bfloat16* __restrict__ dyp16) {
  pdl_launch_dependents();
  pdl_wait();
  const int o = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r0 = threadIdx.x >> 5, nr = blockDim.x >> 5;
  float s = 0.f;
  if (o < O) {
    for (int b = r0; b < B; b += nr) {
      float v = dy[(long)b * O + o];
      if (relu && !(y[(long)b * O + o] > 0.f)) v = 0.f;
      dyp[(long)b * O + o] = v;
      if (dyp16 != nullptr) dyp16[(long)b * O + o] = __float2bfloat16_rn(v);
      s += v;
    }
  }
  __shared__ float sm[32][33];
  sm[r0][threadIdx.x & 31] = s;
  __syncthreads();
  if (r0 == 0 && o < O) {
    float t = 0.f;
    for (int i = 0; i < nr; ++i) t += sm[i][threadIdx.x & 31];
    db[o] = t;
  }
}

// out[n][c] = mean_p x[n][p][c]  (AdaptiveAvgPool2d(1) of the `basic` architecture's trunk,
// archs/HabitatDQNMultiAction.py:33): one thread per (n, c), fp32 accumulation in pixel order
__global__ void avgpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int N, int P, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = (long)N * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long n = i / C;
    const int c = (int)(i - n * C);
    const __nv_bfloat16* px = x + n * P * C + c;
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += __bfloat162float(px[(long)p * C]);
    out[i] = s / (float)P;
  }
}

__global__ void head_flatten_fwd_kernel(const __nv_bfloat16* __restrict__ h, float* __restrict__ flat,
                                        int B, int P, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = (long)B * P * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long t = i / C;
    const int p = (int)(t % P);
    const long b = t / P;
    flat[(b * C + c) * P + p] = __bfloat162float(h[i]);
  }
}

// one block per image: dh[b,p,c] = h > 0 ? dflat[b, c*P + p] : 0 ; dbias[c] += sum
__global__ void head_flatten_bwd_kernel(const float* __restrict__ dflat, const __nv_bfloat16* __restrict__ h,
                                        __nv_bfloat16* __restrict__ dh, float* __restrict__ dbias,
                                        int B, int P, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int c = threadIdx.x;       // blockDim.x == C
  float s = 0.f;
  for (int p = 0; p < P; ++p) {
    const long i = ((long)b * P + p) * C + c;
    float v = dflat[((long)b * C + c) * P + p];
    if (!(__bfloat162float(h[i]) > 0.f)) v = 0.f;
    const __nv_bfloat16 o = __float2bfloat16_rn(v);
    dh[i] = o;
    s += __bfloat162float(o);
  }
  if (dbias != nullptr) atomicAdd(dbias + c, s);
}

// ------------------------------------------------------------------------------------------
// rew / term / valid_mask as the loader types them: int64 (thresholded detections,
// dataloaders/q_learning_real.py:78-84) or, with CONFIDENCE_REWARD, the detector scores themselves, which
// the reference casts with `.float()` (train_q_network.py:158-160): labels_f32 = the already cast fp32
__device__ __forceinline__ float td_label(const void* p, long i, int f32) {
  return f32 ? static_cast<const float*>(p)[i] : (float)static_cast<const int64_t*>(p)[i];
}

// fused TD epilogue: one thread per (sample, class); elements [i_begin, B*C)
__global__ void td_epilogue_kernel(const vdqn_td_desc d, const long i_begin) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = (long)d.B * d.C;
  float local = 0.f;
  for (long i = i_begin + blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long b = i / d.C;
    const float* qs = d.q_s + i * d.A;
    if (d.ground_truth) {
      // regression onto the ground-truth value (train_q_network.py:170-178)
      const double gd = d.gt[i];
      const int act = (int)d.act[b];
      float mask = 1.f, g0 = (float)gd;
      if (d.value_learning && isnan(gd)) { mask = 0.f; g0 = 0.f; }
      const float diff = d.value_learning ? qs[act] * mask - g0 : qs[act] - g0;
      local += 0.5f * diff * diff;
      if (d.dq != nullptr) {
        float* dq = d.dq + i * d.A;
        for (int a = 0; a < d.A; ++a) dq[a] = (a == act) ? diff * mask * d.inv_count : 0.f;
      }
      if (d.y_out != nullptr) d.y_out[i] = g0;
      if (d.best_out != nullptr) d.best_out[i] = 0;
      continue;
    }
    const float* qt = d.q_next_target + i * d.A;
    const float* qsel = d.double_dqn ? d.q_next_online + i * d.A : qt;
    int best = 0;
    float bv = qsel[0];
    for (int a = 1; a < d.A; ++a) {
      const float v = qsel[a];
      if (v > bv) { bv = v; best = a; }            // strict > : first maximum wins (torch.argmax)
    }
    const float term = td_label(d.term, i, d.labels_f32);
    const float q_a = qt[best] * (1.f - term);
    const float rew = td_label(d.rew, i, d.labels_f32);
    float y = d.linear ? rew + (q_a - 0.1f) : rew + d.gamma * q_a;
    if (d.clip_rect) y = fminf(fmaxf(y, 0.f), 1.f);
    const int act = (int)d.act[b];
    const float diff = qs[act] - y;
    float l = 0.5f * diff * diff;
    float mask = 1.f;
    if (d.use_valid) { mask = td_label(d.valid, i, d.labels_f32); l *= mask; }
    local += l;
    if (d.dq != nullptr) {
      float* dq = d.dq + i * d.A;
      for (int a = 0; a < d.A; ++a) dq[a] = (a == act) ? diff * mask * d.inv_count : 0.f;
    }
    if (d.best_out != nullptr) d.best_out[i] = best;
    if (d.y_out != nullptr) d.y_out[i] = y;
  }
  __shared__ float red[32];
  for (int off = 16; off; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (threadIdx.x == 0 && d.loss_out != nullptr) atomicAdd(d.loss_out, v * d.inv_count);
  }
}

// Same computation for 16-byte aligned Q tensors (Bellman branch only), with the three Q arrays and dQ moved
// through shared memory: thread t of a block owns element base + t and its A actions, i.e. a warp touches
// 32*A consecutive floats per array with a stride-A pattern -- as direct loads that is A wavefronts per
// request over the same sectors (the kernel ran at 61 % of HBM copy bandwidth at B = 2^20: LSU-bound, not
// DRAM-bound).  Here each 256-element chunk of every array is loaded and stored as full 16-byte vectors
// (chunk offsets are multiples of 1024*A bytes) and the stride-A accesses hit shared memory, where an odd A
// is conflict-free.  Element -> thread mapping and per-element arithmetic are those of td_epilogue_kernel, so
// results (including the summation order of the loss) are bit-identical.
constexpr int kTdChunk = 256;
constexpr long kTdBulkMinElems = 64 * 1024;     // below this the launch, not HBM, is what the kernel costs

__device__ __forceinline__ void td_stage_in(float* __restrict__ dst, const float* __restrict__ src, int nf) {
  const int nv = nf >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int v = threadIdx.x; v < nv; v += kTdChunk) d4[v] = s4[v];
  for (int v = (nv << 2) + threadIdx.x; v < nf; v += kTdChunk) dst[v] = src[v];
}

__global__ void __launch_bounds__(kTdChunk) td_epilogue_staged_kernel(const vdqn_td_desc d) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float td_sm[];
  const int A = d.A;
  float* sm_qs = td_sm;                       // later reused for dQ
  float* sm_qt = td_sm + kTdChunk * A;
  float* sm_qo = td_sm + 2 * kTdChunk * A;
  const long total = (long)d.B * d.C;
  float local = 0.f;
  for (long base = (long)blockIdx.x * kTdChunk; base < total; base += (long)gridDim.x * kTdChunk) {
    const int n = (int)min((long)kTdChunk, total - base);
    const int nf = n * A;
    // issue the scalar per-element loads first so they are in flight while the tiles are staged
    const long i = base + threadIdx.x;
    const bool live = threadIdx.x < n;
    float term = 0.f, rew = 0.f, mask = 1.f;
    int act = 0;
    if (live) {
      term = td_label(d.term, i, d.labels_f32);
      rew = td_label(d.rew, i, d.labels_f32);
      act = (int)d.act[i / d.C];
      if (d.use_valid) mask = td_label(d.valid, i, d.labels_f32);
    }
    td_stage_in(sm_qs, d.q_s + base * A, nf);
    td_stage_in(sm_qt, d.q_next_target + base * A, nf);
    if (d.double_dqn) td_stage_in(sm_qo, d.q_next_online + base * A, nf);
    __syncthreads();
    if (live) {
      const float* qs = sm_qs + threadIdx.x * A;
      const float* qt = sm_qt + threadIdx.x * A;
      const float* qsel = d.double_dqn ? sm_qo + threadIdx.x * A : qt;
      int best = 0;
      float bv = qsel[0];
      for (int a = 1; a < A; ++a) {
        const float v = qsel[a];
        if (v > bv) { bv = v; best = a; }            // strict > : first maximum wins (torch.argmax)
      }
      const float q_a = qt[best] * (1.f - term);
      float y = d.linear ? rew + (q_a - 0.1f) : rew + d.gamma * q_a;
      if (d.clip_rect) y = fminf(fmaxf(y, 0.f), 1.f);
      const float diff = qs[act] - y;
      float l = 0.5f * diff * diff;
      if (d.use_valid) l *= mask;
      local += l;
      if (d.dq != nullptr) {
        float* dq = sm_qs + threadIdx.x * A;         // own slots only: read above, overwritten here
        for (int a = 0; a < A; ++a) dq[a] = (a == act) ? diff * mask * d.inv_count : 0.f;
      }
      if (d.best_out != nullptr) d.best_out[i] = best;
      if (d.y_out != nullptr) d.y_out[i] = y;
    }
    __syncthreads();
    if (d.dq != nullptr) {
      float* g = d.dq + base * A;
      const int nv = nf >> 2;
      float4* g4 = reinterpret_cast<float4*>(g);
      const float4* s4 = reinterpret_cast<const float4*>(sm_qs);
      for (int v = threadIdx.x; v < nv; v += kTdChunk) g4[v] = s4[v];
      for (int v = (nv << 2) + threadIdx.x; v < nf; v += kTdChunk) g[v] = sm_qs[v];
    }
    __syncthreads();                                  // the tiles are overwritten by the next chunk
  }
  __shared__ float red[32];
  for (int off = 16; off; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (threadIdx.x == 0 && d.loss_out != nullptr) atomicAdd(d.loss_out, v * d.inv_count);
  }
}

// value[b,c] = max_a q[b,c,a] (+ first arg-max): the `model(images).max(2)` of the value-map / policy
// callers (visualize_value.py:96-97, evaluation/evaluate.py:110-114)
__global__ void q_max_kernel(const float* __restrict__ q, float* __restrict__ value, int64_t* __restrict__ arg,
                             long total, int A) {
  pdl_launch_dependents();
  pdl_wait();
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const float* p = q + i * A;
    float bv = p[0];
    int best = 0;
    for (int a = 1; a < A; ++a)
      if (p[a] > bv) { bv = p[a]; best = a; }
    value[i] = bv;
    if (arg != nullptr) arg[i] = best;
  }
}

// ------------------------------------------------------------------------------------------
// inverse-dynamics training step (train_inverse_model.py:85-110): softmax cross-entropy (mean) with
// its gradient and the accuracy count, one thread per sample; element dropout with a counter-based
// generator.
__global__ void cross_entropy_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels,
                                     float* __restrict__ dlogits, float* __restrict__ loss_out,
                                     int* __restrict__ correct_out, int B, int C, float inv_count) {
  pdl_launch_dependents();
  pdl_wait();
  float local = 0.f;
  int hits = 0;
  for (long b = blockIdx.x * (long)blockDim.x + threadIdx.x; b < B; b += (long)gridDim.x * blockDim.x) {
    const float* y = logits + b * C;
    float v[32];
    float mx = y[0];
    int best = 0;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      if (c < C) {
        v[c] = y[c];
        if (v[c] > mx) { mx = v[c]; best = c; }      // strict > : first maximum wins (torch.argmax)
      }
    }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (c < C) { v[c] = expf(v[c] - mx); se += v[c]; }
    const int lab = (int)labels[b];
    const float inv = 1.f / se;
    local += logf(se) + mx - y[lab];
    hits += (best == lab);
    if (dlogits != nullptr) {
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c < C) dlogits[b * C + c] = (v[c] * inv - (c == lab ? 1.f : 0.f)) * inv_count;
    }
  }
  __shared__ float red[32];
  __shared__ int redh[32];
  for (int off = 16; off; off >>= 1) {
    local += __shfl_xor_sync(0xffffffffu, local, off);
    hits += __shfl_xor_sync(0xffffffffu, hits, off);
  }
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = local; redh[threadIdx.x >> 5] = hits; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const bool in = threadIdx.x < (blockDim.x >> 5);
    float t = in ? red[threadIdx.x] : 0.f;
    int h = in ? redh[threadIdx.x] : 0;
    for (int off = 16; off; off >>= 1) {
      t += __shfl_xor_sync(0xffffffffu, t, off);
      h += __shfl_xor_sync(0xffffffffu, h, off);
    }
    if (threadIdx.x == 0) {
      if (loss_out != nullptr) atomicAdd(loss_out, t * inv_count);
      if (correct_out != nullptr && h != 0) atomicAdd(correct_out, h);
    }
  }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__global__ void dropout_mask_kernel(uint8_t* __restrict__ keep, long n, float p, uint64_t seed, uint64_t counter) {
  pdl_launch_dependents();
  pdl_wait();
  const uint64_t key = splitmix64(seed ^ splitmix64(counter));
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const uint64_t r = splitmix64(key + (uint64_t)i);
    const float u = (float)(r >> 40) * (1.0f / 16777216.0f);      // 24 random bits -> [0, 1)
    keep[i] = u >= p ? 1 : 0;
  }
}

__global__ void dropout_apply_kernel(const float* __restrict__ x, const uint8_t* __restrict__ keep, float scale,
                                     float* __restrict__ y, long n) {
  pdl_launch_dependents();
  pdl_wait();
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    y[i] = keep[i] ? x[i] * scale : 0.f;
}

// ------------------------------------------------------------------------------------------
// fused Adam (+ target sync), flat fp32 arenas, 16-byte vectors
__global__ void __launch_bounds__(256)
adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
            float4* __restrict__ v, float4* __restrict__ target, long n4, float step_size,
            float beta1, float beta2, float eps, float sqrt_bc2, float grad_scale,
            const float* __restrict__ dev_scalars) {
  pdl_launch_dependents();
  pdl_wait();
  if (dev_scalars != nullptr) {          // graph-replay mode: step-dependent scalars live in HBM
    step_size = dev_scalars[0];
    sqrt_bc2 = dev_scalars[1];
  }
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4;
       i += (long)gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float* P = &pp.x; float* G = &gg.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gr = G[e] * grad_scale;
      M[e] = beta1 * M[e] + (1.f - beta1) * gr;
      V[e] = beta2 * V[e] + (1.f - beta2) * gr * gr;
      const float denom = sqrtf(V[e]) / sqrt_bc2 + eps;
      P[e] = P[e] - step_size * (M[e] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (target != nullptr) target[i] = pp;
  }
}

// step counter and bias corrections kept on the device so a captured CUDA graph can be replayed
__global__ void adam_scalars_kernel(int* step, float* out, double lr, double b1, double b2) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = *step + 1;
  *step = t;
  const double bc1 = 1.0 - pow(b1, (double)t);
  const double bc2 = 1.0 - pow(b2, (double)t);
  out[0] = (float)(lr / bc1);
  out[1] = (float)sqrt(bc2);
}

static inline int grid_for(long total, int block, int num_sms, int per_sm = 8) {
  long g = (total + block - 1) / block;
  const long cap = (long)num_sms * per_sm;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

}  // namespace vdqn

using namespace vdqn;

#define GET_DEV()                         \
  DeviceInfo* dev = device_info();        \
  if (dev == nullptr) return VDQN_ERR_CUDA; \
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v)

extern "C" int vdqn_weight_prep(const vdqn_wprep_desc* d, void* stream_v) {
  if (d == nullptr || d->w == nullptr || d->w_fwd == nullptr || d->shift == nullptr)
    return set_error(VDQN_ERR_ARG, "weight_prep: null pointer");
  GET_DEV();
  const long total = (long)d->Cout * d->K;
  launch_kernel(weight_prep_kernel, grid_for(total, 256, dev->num_sms), 256, 0, stream, *d);
  VDQN_CHECK_LAUNCH("weight_prep");
  return VDQN_OK;
}

extern "C" int vdqn_weight_prep_multi(const vdqn_wprep_desc* descs_dev, const int64_t* offsets_dev, int32_t n,
                                      int64_t total, void* stream_v) {
  if (descs_dev == nullptr || offsets_dev == nullptr || n < 1)
    return set_error(VDQN_ERR_ARG, "weight_prep_multi: bad arguments");
  GET_DEV();
  if (total == 0) return VDQN_OK;
  launch_kernel(weight_prep_multi_kernel, grid_for(total, 256, dev->num_sms, 16), 256, 0, stream, descs_dev, reinterpret_cast<const long long*>(offsets_dev), n, total);
  VDQN_CHECK_LAUNCH("weight_prep_multi");
  return VDQN_OK;
}

extern "C" int vdqn_weight_prep_tiled(const vdqn_wprep_desc* descs_dev, const int32_t* tile_offsets_dev, int32_t n,
                                      int32_t total_tiles, void* stream_v) {
  if (descs_dev == nullptr || tile_offsets_dev == nullptr || n < 1)
    return set_error(VDQN_ERR_ARG, "weight_prep_tiled: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (total_tiles == 0) return VDQN_OK;
  launch_kernel(weight_prep_tiled_kernel, total_tiles, 256, 0, stream, descs_dev, tile_offsets_dev, n);
  VDQN_CHECK_LAUNCH("weight_prep_tiled");
  return VDQN_OK;
}

// Row-form applicability + launch shape of one reduction (host helper of the two entry points
// below; also exported so a caller can build the table of vdqn_wgrad_finalize_multi).
extern "C" int vdqn_wgrad_finalize_plan(const vdqn_wgrad_fin_desc* d, vdqn_wgrad_fin_item* out) {
  if (d == nullptr || out == nullptr) return 0;
  DeviceInfo* dev = device_info();
  const int num_sms = dev != nullptr ? dev->num_sms : 148;
  const int RS = d->R * d->S;
  if (!(d->kmap == 0 && d->K == RS * d->Cin && d->Cin % 16 == 0 && d->K <= 4608 &&
        (reinterpret_cast<uintptr_t>(d->part) & 15) == 0))
    return 0;
  // input-channel chunks: enough blocks to cover the SMs while a chunk stays >= 16 channels
  int nchunks = 1;
  while (d->Cout * nchunks < 2 * num_sms && d->Cin / (2 * nchunks) >= 16 && d->Cin % (2 * nchunks) == 0)
    nchunks *= 2;
  out->d = *d;
  out->cn = d->Cin / nchunks;
  out->items = RS * (out->cn / 4);
  out->TX = out->items < 256 ? out->items : 256;
  out->TY = 256 / out->TX;
  if (out->TY > d->splits) out->TY = d->splits;
  out->nchunks = nchunks;
  out->first_block = 0;
  // shared memory: TY lanes x RS x (cn + 1) floats
  if ((long)out->TY * RS * (out->cn + 1) > 4608 + 160) return 0;
  return nchunks * d->Cout;
}

extern "C" int vdqn_wgrad_finalize_multi(const vdqn_wgrad_fin_item* items_dev, int32_t n, int32_t total_blocks,
                                         void* stream_v) {
  if (items_dev == nullptr || n < 1 || total_blocks < 1)
    return set_error(VDQN_ERR_ARG, "wgrad_finalize_multi: empty table");
  GET_DEV();
  (void)dev;
  launch_kernel(wgrad_finalize_multi_kernel, total_blocks, 256, 0, stream, items_dev, n);
  VDQN_CHECK_LAUNCH("wgrad_finalize_multi");
  return VDQN_OK;
}

extern "C" int vdqn_wgrad_finalize(const vdqn_wgrad_fin_desc* d, void* stream_v) {
  if (d == nullptr || d->part == nullptr || d->w == nullptr || d->dw == nullptr)
    return set_error(VDQN_ERR_ARG, "wgrad_finalize: null pointer");
  GET_DEV();
  (void)dev;
  vdqn_wgrad_fin_item item;
  if (vdqn_wgrad_finalize_plan(d, &item) > 0) {
    FinRowCfg c;
    c.cn = item.cn; c.items = item.items; c.TX = item.TX; c.TY = item.TY;
    launch_kernel(wgrad_finalize_rows_kernel, dim3(item.nchunks, d->Cout), 256, 0, stream, *d, c);
    VDQN_CHECK_LAUNCH("wgrad_finalize_rows");
    return VDQN_OK;
  }
  launch_kernel(wgrad_finalize_kernel, dim3((d->K + 63) / 64, d->Cout), dim3(64, 4), 0, stream, *d);
  VDQN_CHECK_LAUNCH("wgrad_finalize");
  return VDQN_OK;
}

extern "C" int vdqn_stem_pack_f32(const float* x, void* out, int32_t N, int32_t H, int32_t W, void* stream_v) {
  if (x == nullptr || out == nullptr) return set_error(VDQN_ERR_ARG, "stem_pack: null pointer");
  if ((H | W) & 1) return set_error(VDQN_ERR_SHAPE, "stem_pack: H and W must be even");
  GET_DEV();
  const long total = (long)N * (H / 2) * (W / 2);
  if (total == 0) return VDQN_OK;
  launch_kernel(stem_pack_f32_kernel, grid_for(total, 256, dev->num_sms, 16), 256, 0, stream, x, static_cast<__nv_bfloat16*>(out), N, H, W);
  VDQN_CHECK_LAUNCH("stem_pack_f32");
  return VDQN_OK;
}

extern "C" int vdqn_stem_pack_u8(const uint8_t* x, void* out, int32_t N, int32_t H, int32_t W, void* stream_v) {
  if (x == nullptr || out == nullptr) return set_error(VDQN_ERR_ARG, "stem_pack: null pointer");
  if ((H & 1) || (W & 3)) return set_error(VDQN_ERR_SHAPE, "stem_pack_u8: H must be even and W a multiple of 4");
  if (reinterpret_cast<uintptr_t>(x) & 3) return set_error(VDQN_ERR_ARG, "stem_pack_u8: frames must be 4-byte aligned");
  GET_DEV();
  const long total = (long)N * (H / 2) * (W / 4);
  if (total == 0) return VDQN_OK;
  launch_kernel(stem_pack_u8_kernel, grid_for(total, 256, dev->num_sms, 16), 256, 0, stream, x, static_cast<__nv_bfloat16*>(out), N, H, W);
  VDQN_CHECK_LAUNCH("stem_pack_u8");
  return VDQN_OK;
}

extern "C" int vdqn_maxpool_fwd(const void* x, void* y, uint8_t* idx, int32_t N, int32_t H, int32_t W,
                                int32_t C, void* stream_v) {
  if (x == nullptr || y == nullptr) return set_error(VDQN_ERR_ARG, "maxpool_fwd: null pointer");
  if (C % 8) return set_error(VDQN_ERR_SHAPE, "maxpool: C must be a multiple of 8");
  GET_DEV();
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long total = (long)N * Ho * Wo * (C / 8);
  if (total == 0) return VDQN_OK;
  if (idx != nullptr)
    launch_kernel(maxpool_fwd_kernel<true>, grid_for(total, 256, dev->num_sms, 16), 256, 0, stream, static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), idx, N, H, W, C);
  else
    launch_kernel(maxpool_fwd_kernel<false>, grid_for(total, 256, dev->num_sms, 16), 256, 0, stream, static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), idx, N, H, W, C);
  VDQN_CHECK_LAUNCH("maxpool_fwd");
  return VDQN_OK;
}

extern "C" int vdqn_maxpool_bwd(const void* dy, const uint8_t* idx, const void* y, void* dx, float* colsum,
                                int32_t N, int32_t H, int32_t W, int32_t C, void* stream_v) {
  if (dy == nullptr || idx == nullptr || y == nullptr || dx == nullptr)
    return set_error(VDQN_ERR_ARG, "maxpool_bwd: null pointer");
  if (C % 8 || 256 % (C / 8) || C > 256) return set_error(VDQN_ERR_SHAPE, "maxpool_bwd: unsupported C=%d", C);
  if ((H | W) & 1) return set_error(VDQN_ERR_SHAPE, "maxpool_bwd: H and W must be even");
  GET_DEV();
  const long rows = (long)N * (H / 2);
  if (rows == 0 || W == 0) return VDQN_OK;
  if (rows > 0x7fffffffL) return set_error(VDQN_ERR_SHAPE, "maxpool_bwd: too many rows");
  // a block per pooled row at a time; its threads are the row's (column, channel group) items, split evenly
  // over the passes a block of <= 256 threads needs (W/2 = 56, C = 64: 448 items = 2 passes of 224 threads)
  const int items = (W / 2) * (C / 8);
  const int passes = (items + 255) / 256;
  int threads = ((items + passes - 1) / passes + 31) / 32 * 32;
  if (threads < C) threads = (C + 31) / 32 * 32;         // the block reduction of the column sums needs C threads
  const long want = (long)dev->num_sms * 8;
  const int grid = (int)(rows < want ? rows : want);
  launch_kernel(maxpool_bwd_kernel, grid, threads, threads * 8 * sizeof(float), stream, static_cast<const __nv_bfloat16*>(dy), idx, static_cast<const __nv_bfloat16*>(y),
      static_cast<__nv_bfloat16*>(dx), colsum, N, H, W, C);
  VDQN_CHECK_LAUNCH("maxpool_bwd");
  return VDQN_OK;
}

static int launch_sgemm(const float* A, const float* B, float* C, const float* bias, int M, int N, int K,
                        long a_rs, long a_cs, long b_rs, long b_cs, int ldc, int relu, cudaStream_t stream) {
  if (M == 0 || N == 0) return VDQN_OK;
  DeviceInfo* dev = device_info();
  if (dev == nullptr) return VDQN_ERR_CUDA;
  dim3 grid((N + 63) / 64, (M + 63) / 64, 1);
  // under-filled grids (the MLP has M = batch): split K across blockIdx.z.  The kernel is single-
  // buffered, so a block's k-loop runs at global-memory latency; about four resident blocks per SM
  // hide it (one block per SM left the 1600-deep top.0 products at 40-65 us).
  const int tiles = grid.x * grid.y;
  int splitk = 1;
  if (tiles < 4 * dev->num_sms && ldc == N) {
    splitk = (4 * dev->num_sms + tiles - 1) / tiles;
    if (splitk > 16) splitk = 16;
    while (splitk > 1 && K / splitk < 64) --splitk;
  }
  int k_len = (K + splitk - 1) / splitk;
  k_len = (k_len + 15) / 16 * 16;
  splitk = (K + k_len - 1) / k_len;
  grid.z = splitk;
  if (splitk > 1) {
    cudaError_t e = cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * N, stream);
    if (e != cudaSuccess) return set_error(VDQN_ERR_CUDA, "sgemm memset: %s", cudaGetErrorString(e));
  }
  launch_kernel(sgemm_strided_kernel, grid, 256, 0, stream, A, B, C, bias, M, N, K, a_rs, a_cs, b_rs, b_cs, ldc, relu,
                                                 k_len);
  VDQN_CHECK_LAUNCH("sgemm");
  if (splitk > 1 && (bias != nullptr || relu)) {
    const long total = (long)M * N;
    launch_kernel(bias_act_kernel, grid_for(total, 256, dev->num_sms), 256, 0, stream, C, bias, total, N, relu);
    VDQN_CHECK_LAUNCH("bias_act");
  }
  return VDQN_OK;
}

extern "C" int vdqn_linear_fwd(const float* x, const float* w, const float* bias, float* y, int32_t B,
                               int32_t K, int32_t O, int32_t relu, void* stream_v) {
  if (x == nullptr || w == nullptr || y == nullptr) return set_error(VDQN_ERR_ARG, "linear_fwd: null pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  // y[b,o] = sum_k x[b,k] * w[o,k]
  return launch_sgemm(x, w, y, bias, B, O, K, K, 1, 1, K, O, relu, stream);
}

extern "C" int vdqn_linear_bwd(const float* x, const float* w, const float* y, float* dy, float* dx,
                               float* dw, float* db, int32_t B, int32_t K, int32_t O, int32_t relu,
                               void* stream_v) {
  if (x == nullptr || w == nullptr || dy == nullptr || dw == nullptr || db == nullptr)
    return set_error(VDQN_ERR_ARG, "linear_bwd: null pointer");
  if (relu && y == nullptr) return set_error(VDQN_ERR_ARG, "linear_bwd: relu needs y");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (B == 0) return VDQN_OK;
  // dy is overwritten in place with the masked gradient (the caller owns it as scratch)
  float* dyp = dy;
  launch_kernel(mask_colsum_kernel, (O + 31) / 32, 1024, 0, stream, dy, y, dyp, db, B, O, relu, nullptr);
  VDQN_CHECK_LAUNCH("mask_colsum");
  // dw[o,k] = sum_b dyp[b,o] * x[b,k]
  int rc = launch_sgemm(dyp, x, dw, nullptr, O, K, B, 1, O, K, 1, K, 0, stream);
  if (rc != VDQN_OK) return rc;
  // dx[b,k] = sum_o dyp[b,o] * w[o,k]
  if (dx != nullptr) rc = launch_sgemm(dyp, w, dx, nullptr, B, K, O, O, 1, K, 1, K, 0, stream);
  return rc;
}

extern "C" int vdqn_relu_mask_colsum(float* dy, const float* y, void* dy_bf16, float* db, int32_t B, int32_t O,
                                     int32_t relu, void* stream_v) {
  if (dy == nullptr || db == nullptr) return set_error(VDQN_ERR_ARG, "relu_mask_colsum: null pointer");
  if (relu && y == nullptr) return set_error(VDQN_ERR_ARG, "relu_mask_colsum: relu needs y");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (B == 0) return VDQN_OK;
  launch_kernel(mask_colsum_kernel, (O + 31) / 32, 1024, 0, stream, dy, y, dy, db, B, O, relu,
                static_cast<__nv_bfloat16*>(dy_bf16));
  VDQN_CHECK_LAUNCH("relu_mask_colsum");
  return VDQN_OK;
}

extern "C" int vdqn_head_flatten_fwd(const void* h, float* flat, int32_t B, int32_t P, int32_t C, void* stream_v) {
  if (h == nullptr || flat == nullptr) return set_error(VDQN_ERR_ARG, "head_flatten_fwd: null pointer");
  GET_DEV();
  const long total = (long)B * P * C;
  if (total == 0) return VDQN_OK;
  launch_kernel(head_flatten_fwd_kernel, grid_for(total, 256, dev->num_sms), 256, 0, stream, static_cast<const __nv_bfloat16*>(h), flat, B, P, C);
  VDQN_CHECK_LAUNCH("head_flatten_fwd");
  return VDQN_OK;
}

extern "C" int vdqn_avgpool_fwd(const void* x, float* out, int32_t N, int32_t P, int32_t C, void* stream_v) {
  if (x == nullptr || out == nullptr) return set_error(VDQN_ERR_ARG, "avgpool_fwd: null pointer");
  GET_DEV();
  const long total = (long)N * C;
  if (total == 0) return VDQN_OK;
  launch_kernel(avgpool_fwd_kernel, grid_for(total, 256, dev->num_sms), 256, 0, stream,
                static_cast<const __nv_bfloat16*>(x), out, N, P, C);
  VDQN_CHECK_LAUNCH("avgpool_fwd");
  return VDQN_OK;
}

extern "C" int vdqn_head_flatten_bwd(const float* dflat, const void* h, void* dh, float* dbias, int32_t B,
                                     int32_t P, int32_t C, void* stream_v) {
  if (dflat == nullptr || h == nullptr || dh == nullptr) return set_error(VDQN_ERR_ARG, "head_flatten_bwd: null pointer");
  if (C > 1024) return set_error(VDQN_ERR_SHAPE, "head_flatten_bwd: C too large");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (B == 0) return VDQN_OK;
  launch_kernel(head_flatten_bwd_kernel, B, C, 0, stream, dflat, static_cast<const __nv_bfloat16*>(h),
                                               static_cast<__nv_bfloat16*>(dh), dbias, B, P, C);
  VDQN_CHECK_LAUNCH("head_flatten_bwd");
  return VDQN_OK;
}

extern "C" int vdqn_td_epilogue(const vdqn_td_desc* d, void* stream_v) {
  if (d == nullptr || d->q_s == nullptr || d->act == nullptr)
    return set_error(VDQN_ERR_ARG, "td_epilogue: null pointer");
  if (d->ground_truth) {
    if (d->gt == nullptr) return set_error(VDQN_ERR_ARG, "td_epilogue: ground-truth mode needs gt");
  } else {
    if (d->q_next_target == nullptr || d->rew == nullptr || d->term == nullptr)
      return set_error(VDQN_ERR_ARG, "td_epilogue: null pointer");
    if (d->double_dqn && d->q_next_online == nullptr)
      return set_error(VDQN_ERR_ARG, "td_epilogue: double DQN needs q_next_online");
  }
  if (d->use_valid && d->valid == nullptr) return set_error(VDQN_ERR_ARG, "td_epilogue: valid mask missing");
  if (d->A < 1 || d->C < 1 || d->B < 0) return set_error(VDQN_ERR_SHAPE, "td_epilogue: bad shape");
  GET_DEV();
  const long total = (long)d->B * d->C;
  if (total == 0) return VDQN_OK;
  // staged variant: Bellman branch, a few actions, every Q tensor and dQ 16-byte aligned (chunk offsets are
  // multiples of 1024*A bytes, so base alignment is all that is needed); VDQN_TD_STAGED=0 forces the direct one
  auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  static const bool staged_ok = []() {
    const char* e = getenv("VDQN_TD_STAGED");
    return !(e != nullptr && e[0] == '0');
  }();
  // large batches: the streaming form (td_bulk.cu) on the full 1024-element chunks, the per-thread kernel
  // on the tail; VDQN_TD_BULK=0 switches it off
  static const bool bulk_ok = []() {
    const char* e = getenv("VDQN_TD_BULK");
    return !(e != nullptr && e[0] == '0');
  }();
  if (bulk_ok && total >= kTdBulkMinElems && td_bulk_supported(d)) {
    const long n_chunks = total / 1024;
    const int rc = td_bulk_launch(d, n_chunks, stream);
    if (rc != VDQN_OK) return rc;
    const long done = n_chunks * 1024;
    if (done < total) {
      launch_kernel(td_epilogue_kernel, grid_for(total - done, 256, dev->num_sms, 8), 256, 0, stream, *d, done);
      VDQN_CHECK_LAUNCH("td_epilogue (tail)");
    }
    return VDQN_OK;
  }
  if (staged_ok && !d->ground_truth && d->A <= 8 && al16(d->q_s) && al16(d->q_next_target) &&
      al16(d->q_next_online) && al16(d->dq)) {
    const size_t smem = sizeof(float) * 3 * kTdChunk * (size_t)d->A;
    launch_kernel(td_epilogue_staged_kernel, grid_for(total, kTdChunk, dev->num_sms, 8), kTdChunk, smem, stream, *d);
    VDQN_CHECK_LAUNCH("td_epilogue_staged");
    return VDQN_OK;
  }
  launch_kernel(td_epilogue_kernel, grid_for(total, 256, dev->num_sms, 8), 256, 0, stream, *d, 0L);
  VDQN_CHECK_LAUNCH("td_epilogue");
  return VDQN_OK;
}

static int adam_check(const float* p, const float* g, const float* m, const float* v, const float* target,
                      int64_t n) {
  if (p == nullptr || g == nullptr || m == nullptr || v == nullptr)
    return set_error(VDQN_ERR_ARG, "adam: null pointer");
  if (n % 4 != 0) return set_error(VDQN_ERR_SHAPE, "adam: arena length must be a multiple of 4");
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
       reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(target)) & 15)
    return set_error(VDQN_ERR_ARG, "adam: arenas must be 16-byte aligned");
  return VDQN_OK;
}

extern "C" int vdqn_q_max(const float* q, float* value, int64_t* arg, int64_t rows, int32_t A, void* stream_v) {
  if (q == nullptr || value == nullptr) return set_error(VDQN_ERR_ARG, "q_max: null pointer");
  if (A < 1 || rows < 0) return set_error(VDQN_ERR_SHAPE, "q_max: bad shape");
  GET_DEV();
  if (rows == 0) return VDQN_OK;
  launch_kernel(q_max_kernel, grid_for(rows, 256, dev->num_sms), 256, 0, stream, q, value, arg, rows, A);
  VDQN_CHECK_LAUNCH("q_max");
  return VDQN_OK;
}

extern "C" int vdqn_cross_entropy(const float* logits, const int64_t* labels, float* dlogits, float* loss_out,
                                  int32_t* correct_out, int32_t B, int32_t C, float inv_count, void* stream_v) {
  if (C < 1 || C > 32 || B < 0) return set_error(VDQN_ERR_SHAPE, "cross_entropy: bad shape (1 <= C <= 32)");
  if (B == 0) return VDQN_OK;                     // empty batch: nothing to add to loss / correct
  if (logits == nullptr || labels == nullptr) return set_error(VDQN_ERR_ARG, "cross_entropy: null pointer");
  GET_DEV();
  launch_kernel(cross_entropy_kernel, grid_for(B, 128, dev->num_sms), 128, 0, stream, logits, labels, dlogits,
                loss_out, correct_out, B, C, inv_count);
  VDQN_CHECK_LAUNCH("cross_entropy");
  return VDQN_OK;
}

extern "C" int vdqn_dropout_mask(uint8_t* keep, int64_t n, float p, uint64_t seed, uint64_t counter,
                                 void* stream_v) {
  if (keep == nullptr) return set_error(VDQN_ERR_ARG, "dropout_mask: null pointer");
  if (!(p >= 0.f && p < 1.f) || n < 0) return set_error(VDQN_ERR_ARG, "dropout_mask: need 0 <= p < 1");
  GET_DEV();
  if (n == 0) return VDQN_OK;
  launch_kernel(dropout_mask_kernel, grid_for(n, 256, dev->num_sms), 256, 0, stream, keep, (long)n, p,
                (uint64_t)seed, (uint64_t)counter);
  VDQN_CHECK_LAUNCH("dropout_mask");
  return VDQN_OK;
}

extern "C" int vdqn_dropout_apply(const float* x, const uint8_t* keep, float scale, float* y, int64_t n,
                                  void* stream_v) {
  if (x == nullptr || keep == nullptr || y == nullptr) return set_error(VDQN_ERR_ARG, "dropout_apply: null pointer");
  if (n < 0) return set_error(VDQN_ERR_SHAPE, "dropout_apply: bad shape");
  GET_DEV();
  if (n == 0) return VDQN_OK;
  launch_kernel(dropout_apply_kernel, grid_for(n, 256, dev->num_sms), 256, 0, stream, x, keep, scale, y, (long)n);
  VDQN_CHECK_LAUNCH("dropout_apply");
  return VDQN_OK;
}

extern "C" int vdqn_adam_fused(float* p, const float* g, float* m, float* v, float* target, int64_t n,
                               double lr, double beta1, double beta2, double eps, int32_t step,
                               float grad_scale, void* stream_v) {
  int rc = adam_check(p, g, m, v, target, n);
  if (rc != VDQN_OK) return rc;
  if (step < 1) return set_error(VDQN_ERR_ARG, "adam: step must be >= 1");
  GET_DEV();
  if (n == 0) return VDQN_OK;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  launch_kernel(adam_kernel, grid_for(n / 4, 256, dev->num_sms, 8), 256, 0, stream, reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
      reinterpret_cast<float4*>(v), reinterpret_cast<float4*>(target), n / 4, (float)(lr / bc1),
      (float)beta1, (float)beta2, (float)eps, (float)sqrt(bc2), grad_scale, nullptr);
  VDQN_CHECK_LAUNCH("adam");
  return VDQN_OK;
}

extern "C" int vdqn_adam_fused_graph(float* p, const float* g, float* m, float* v, float* target,
                                     int64_t n, double lr, double beta1, double beta2, double eps,
                                     float grad_scale, int32_t* step_dev, float* scalars_dev,
                                     void* stream_v) {
  int rc = adam_check(p, g, m, v, target, n);
  if (rc != VDQN_OK) return rc;
  if (step_dev == nullptr || scalars_dev == nullptr)
    return set_error(VDQN_ERR_ARG, "adam_graph: device step counter / scalar buffer missing");
  GET_DEV();
  launch_kernel(adam_scalars_kernel, 1, 1, 0, stream, step_dev, scalars_dev, lr, beta1, beta2);
  VDQN_CHECK_LAUNCH("adam_scalars");
  if (n == 0) return VDQN_OK;
  launch_kernel(adam_kernel, grid_for(n / 4, 256, dev->num_sms, 8), 256, 0, stream, reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
      reinterpret_cast<float4*>(v), reinterpret_cast<float4*>(target), n / 4, 0.f, (float)beta1,
      (float)beta2, (float)eps, 1.f, grad_scale, scalars_dev);
  VDQN_CHECK_LAUNCH("adam");
  return VDQN_OK;
}
