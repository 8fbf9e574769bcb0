// Q-head MLP (top.0 / top.2 / top.4: Linear 1600F-512-256-5A, archs/HabitatDQNMultiAction.py:31,53) on the
// tensor cores at fp32-grade accuracy.
//
// The MLP is 0.05 % of the step's FLOPs but was 6 % of its time as ~35 fp32 SIMT launches.  Plain bf16
// operands are not an option: rounding top.0's weights alone moves the worst Q value by 8e-3, past the
// 1e-2 parity bar.  So every fp32 operand x is SPLIT into two bf16 numbers, x = hi + lo with
// hi = bf16(x), lo = bf16(x - hi) (|x - hi - lo| <= 2^-17 |x|), and a product a*b is accumulated in fp32 as
// a_hi*b_hi + a_lo*b_hi + a_hi*b_lo (the dropped lo*lo term is 2^-18 relative): one GEMM becomes up to
// three "segments" that accumulate into the same TMEM tile.  The head-conv output is bf16 already (one
// segment fewer).  Results agree with an fp32 reference to ~1e-5 (tests/test_gpu_teacher_forced.py).
//
//   D[m, n] = sum_seg sum_k A_seg[m, k] * B_seg[n, k]          (fp32 accumulation in TMEM)
//
// Each operand is a row-major bf16 matrix read through a 2-D tensor map either K-major (rows = m or n,
// k contiguous) or MN-major (rows = k, m or n contiguous), so the same buffers serve the forward GEMM
// (x W^T), the data gradient (dy W: W read MN-major) and the weight gradient (dy^T x: both read MN-major)
// without a transposed copy.  One CTA per 128 x BN output tile: warp 0 TMA producer, warp 1 MMA issuer,
// warps 2-5 epilogue (+bias, ReLU, ReLU mask of a stored activation, hi/lo split of the result for the
// next GEMM, bf16 output, per-column sums = bias gradients, un-permuting store for d top.0.weight).
#include <cuda_bf16.h>

#include "epilogue.cuh"
#include "ptx.cuh"
#include "vdqn_internal.h"

namespace vdqn {

constexpr int kMlpBM = 128, kMlpKB = 64, kMlpMaxStages = 8, kMlpMaxSeg = 3;
constexpr int kMlpSlab = 64 * 64 * 2;          // MN-major slab: 64 k-rows x 64 elements

struct MlpArgs {
  int M, N, nseg, BN;
  int K[kMlpMaxSeg];
  int a_mn, b_mn;
  int split_mtile;                 // m-tiles >= this use the second set of B maps / bias2 (0: off)
  const float* bias; const float* bias2;
  int relu;
  const float* mask_f32; const __nv_bfloat16* mask_bf16; int ldmask;
  float* out_f32; int ld_f32;
  int perm_c, perm_p;              // != 0: column n = f*(c*p) + pp*perm_c + cc is stored at f*(c*p) + cc*perm_p + pp
  __nv_bfloat16* out_hi; __nv_bfloat16* out_lo; int ld_hl;
  __nv_bfloat16* out_bf16; int ld_bf16;
  float* colsum; int colsum_mod;
};

constexpr int kMlpMaxProb = 4;      // problems per launch (grouped launch: e.g. the three weight gradients + the head dy)

struct MlpMaps {
  CUtensorMap a[kMlpMaxProb][kMlpMaxSeg], b[kMlpMaxProb][kMlpMaxSeg];
  CUtensorMap b2[kMlpMaxSeg];       // second operand set of problem 0 (dual-network forward)
};

struct MlpLaunch {
  int nprob, stages;
  int first_cta[kMlpMaxProb + 1];
  MlpArgs p[kMlpMaxProb];
};

// Grouped launch: CTA ranges [first_cta[i], first_cta[i+1]) work on problem i.  These GEMMs are latency-bound
// (a few hundred KB per CTA through a ~1.5 us TMA round trip), so independent ones share one launch and the
// pipeline is as deep as shared memory allows (`stages`: 8 for a launch with few CTAs, 4 when two CTAs
// should fit on an SM).
__global__ void __launch_bounds__(192, 1) mlp_gemm_kernel(const __grid_constant__ MlpMaps maps,
                                                          const __grid_constant__ MlpLaunch L) {
  int pi = 0;
  while (pi + 1 < L.nprob && (int)blockIdx.x >= L.first_cta[pi + 1]) ++pi;
  const MlpArgs& a = L.p[pi];
  const int cta = (int)blockIdx.x - L.first_cta[pi];
  const int nstages = L.stages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int BN = a.BN;
  const uint32_t a_bytes = kMlpBM * kMlpKB * 2, b_bytes = (uint32_t)BN * kMlpKB * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const uint32_t bar_base = smem_base + nstages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMlpMaxStages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * kMlpMaxStages);
  const uint32_t tmem_slot = tfull_bar + 8u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t tmem_cols = BN <= 32 ? 32u : (BN <= 64 ? 64u : (BN <= 128 ? 128u : 256u));
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    for (int sg = 0; sg < a.nseg; ++sg) {
      tma_prefetch_desc(&maps.a[pi][sg]);
      tma_prefetch_desc(&maps.b[pi][sg]);
    }
    for (int s = 0; s < nstages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  const int n_tiles = (a.N + BN - 1) / BN;
  const int m_t = cta / n_tiles, n_t = cta % n_tiles;
  const int m0 = m_t * kMlpBM, n0 = n_t * BN;
  const bool second = a.split_mtile > 0 && m_t >= a.split_mtile;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int sg = 0; sg < a.nseg; ++sg) {
      const CUtensorMap* tmA = &maps.a[pi][sg];
      const CUtensorMap* tmB = second ? &maps.b2[sg] : &maps.b[pi][sg];
      const int nkb = (a.K[sg] + kMlpKB - 1) / kMlpKB;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (elect_one()) {
          const uint32_t sA = smem_base + stage * stage_bytes, sB = sA + a_bytes;
          const int k0 = kb * kMlpKB;
          mbar_expect_tx(full_bar(stage), stage_bytes);
          if (a.a_mn) {
            tma_load_2d(sA, tmA, full_bar(stage), m0, k0);
            tma_load_2d(sA + kMlpSlab, tmA, full_bar(stage), m0 + 64, k0);
          } else {
            tma_load_2d(sA, tmA, full_bar(stage), k0, m0);
          }
          if (a.b_mn) {
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + j * kMlpSlab, tmB, full_bar(stage), n0 + 64 * j, k0);
          } else {
            tma_load_2d(sB, tmB, full_bar(stage), k0, n0);
          }
        }
        __syncwarp();
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(128, BN, a.a_mn, a.b_mn);
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    for (int sg = 0; sg < a.nseg; ++sg) {
      const int nkb = (a.K[sg] + kMlpKB - 1) / kMlpKB;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sA = smem_base + stage * stage_bytes, sB = sA + a_bytes;
#pragma unroll
          for (int k = 0; k < kMlpKB / 16; ++k) {
            const uint64_t ad = a.a_mn ? make_smem_desc(sA + k * 16 * 128, kMlpSlab, 8 * 128, kSwz128)
                                       : make_smem_desc(sA + k * 32, 16, 8 * 128, kSwz128);
            const uint64_t bd = a.b_mn ? make_smem_desc(sB + k * 16 * 128, kMlpSlab, 8 * 128, kSwz128)
                                       : make_smem_desc(sB + k * 32, 16, 8 * 128, kSwz128);
            umma_f16(tmem_base, ad, bd, idesc, (first && k == 0) ? 0u : 1u);
          }
          first = false;
          umma_commit(empty_bar(stage));
        }
        first = false;
        __syncwarp();
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
    }
    if (elect_one()) umma_commit(tfull_bar);
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int m = m0 + quad * 32 + lane;
    const bool mvalid = m < a.M;
    const float* bias = second ? a.bias2 : a.bias;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t raw[32];
      tmem_ld_32x32(tmem_base + c0 + ((uint32_t)(quad * 32) << 16), raw);
      tmem_ld_wait();
      const int nb = n0 + c0;
      if (nb >= a.N) break;
      const bool full = nb + 32 <= a.N;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
      if (bias != nullptr) {
        if (full) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + nb) + j);
            v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += (nb + j < a.N) ? __ldg(bias + nb + j) : 0.f;
        }
      }
      if (a.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (a.mask_f32 != nullptr && mvalid) {
        const float* mp = a.mask_f32 + (long)m * a.ldmask + nb;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (nb + j < a.N && !(__ldg(mp + j) > 0.f)) v[j] = 0.f;
      }
      if (a.mask_bf16 != nullptr && mvalid) {
        const __nv_bfloat16* mp = a.mask_bf16 + (long)m * a.ldmask + nb;
        if (full && (a.ldmask & 7) == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(mp) + j);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r4);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(h[e]);
              if (!(f.x > 0.f)) v[8 * j + 2 * e] = 0.f;
              if (!(f.y > 0.f)) v[8 * j + 2 * e + 1] = 0.f;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nb + j < a.N && !(__bfloat162float(mp[j]) > 0.f)) v[j] = 0.f;
        }
      }
      if (!mvalid) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (nb + j >= a.N) v[j] = 0.f;
      if (mvalid) {
        if (a.out_f32 != nullptr) {
          if (a.perm_c != 0) {
            const int cp = a.perm_c * a.perm_p;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = nb + j;
              if (n < a.N) {
                const int f = n / cp, r = n - f * cp;
                const int pp = r / a.perm_c, cc = r - pp * a.perm_c;
                a.out_f32[(long)m * a.ld_f32 + f * cp + cc * a.perm_p + pp] = v[j];
              }
            }
          } else if (full && (a.ld_f32 & 3) == 0) {
            float4* op = reinterpret_cast<float4*>(a.out_f32 + (long)m * a.ld_f32 + nb);
#pragma unroll
            for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < a.N) a.out_f32[(long)m * a.ld_f32 + nb + j] = v[j];
          }
        }
        if (a.out_hi != nullptr) {
          // columns up to the padded leading dimension are written (zeros beyond N): the next GEMM reads them
          __nv_bfloat16* hp = a.out_hi + (long)m * a.ld_hl + nb;
          __nv_bfloat16* lp = a.out_lo + (long)m * a.ld_hl + nb;
          if (nb + 32 <= a.ld_hl && (a.ld_hl & 7) == 0) {
            // 16-byte stores: a thread owns a row, 2-byte stores would touch 32 sectors per instruction
            uint4 ph[4], pl[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&ph[j]);
              __nv_bfloat162* l2 = reinterpret_cast<__nv_bfloat162*>(&pl[j]);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float x0 = v[8 * j + 2 * e], x1 = v[8 * j + 2 * e + 1];
                const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
                const float2 hf = __bfloat1622float2(hh);
                h2[e] = hh;
                l2[e] = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              reinterpret_cast<uint4*>(hp)[j] = ph[j];
              reinterpret_cast<uint4*>(lp)[j] = pl[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (nb + j < a.ld_hl) {
                const __nv_bfloat16 h = __float2bfloat16_rn(v[j]);
                hp[j] = h;
                lp[j] = __float2bfloat16_rn(v[j] - __bfloat162float(h));
              }
            }
          }
        }
        if (a.out_bf16 != nullptr) {
          __nv_bfloat16* op = a.out_bf16 + (long)m * a.ld_bf16 + nb;
          if (full && (a.ld_bf16 & 7) == 0) {
            uint4 pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk[j]);
#pragma unroll
              for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) reinterpret_cast<uint4*>(op)[j] = pk[j];
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (nb + j < a.N) op[j] = __float2bfloat16_rn(v[j]);
          }
        }
      }
      if (a.colsum != nullptr) {
        if (a.out_bf16 != nullptr) {       // sums of the values as stored (what the consumer of dh sees)
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
        }
        const float cs = warp_transpose_reduce(v, lane);
        const int n = nb + lane;
        if (n < a.N) atomicAdd(a.colsum + (a.colsum_mod > 0 ? n % a.colsum_mod : n), cs);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// x fp32 [rows][cols] -> hi / lo bf16 [rows][ld_out] (columns >= cols zeroed) and optional column sums;
// with perm_c != 0 the columns are re-ordered on the way (top.0.weight: reference order c*P + p per frame
// -> the NHWC order p*C + c in which the head-conv output lies in memory)
__global__ void split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, int rows, int cols, int ld_out, int perm_c,
                                  int perm_p, float* __restrict__ colsum) {
  pdl_launch_dependents();
  pdl_wait();
  const long total = (long)rows * ld_out;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld_out), c = (int)(i - (long)r * ld_out);
    float v = 0.f;
    if (c < cols) {
      int src = c;
      if (perm_c != 0) {
        const int cp = perm_c * perm_p, f = c / cp, rr = c - f * cp;
        const int pp = rr / perm_c, cc = rr - pp * perm_c;
        src = f * cp + cc * perm_p + pp;
      }
      v = x[(long)r * cols + src];
      if (colsum != nullptr) atomicAdd(colsum + c, v);
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

}  // namespace vdqn

using namespace vdqn;

extern "C" int vdqn_split_bf16(const float* x, void* hi, void* lo, int32_t rows, int32_t cols, int32_t ld_out,
                               int32_t perm_c, int32_t perm_p, float* colsum, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (x == nullptr || hi == nullptr || lo == nullptr) return set_error(VDQN_ERR_ARG, "split_bf16: null pointer");
  if (rows < 0 || cols < 1 || ld_out < cols) return set_error(VDQN_ERR_SHAPE, "split_bf16: bad shape");
  if (perm_c != 0 && (perm_p < 1 || cols % (perm_c * perm_p) != 0))
    return set_error(VDQN_ERR_SHAPE, "split_bf16: columns are not whole (c, p) blocks");
  DeviceInfo* dev = device_info();
  if (dev == nullptr) return VDQN_ERR_CUDA;
  const long total = (long)rows * ld_out;
  if (total == 0) return VDQN_OK;
  long blocks = (total + 255) / 256;
  if (blocks > (long)dev->num_sms * 8) blocks = (long)dev->num_sms * 8;
  launch_kernel(split_bf16_kernel, (int)blocks, 256, 0, stream, x, static_cast<__nv_bfloat16*>(hi),
                static_cast<__nv_bfloat16*>(lo), rows, cols, ld_out, perm_c, perm_p, colsum);
  VDQN_CHECK_LAUNCH("split_bf16");
  return VDQN_OK;
}

static int build_problem(const vdqn_mlp_gemm_desc* d, int pi, MlpMaps& maps, MlpArgs& a) {
  if (d->nseg < 1 || d->nseg > kMlpMaxSeg) return set_error(VDQN_ERR_ARG, "mlp_gemm: 1..3 segments");
  if (d->M < 1 || d->N < 1) return set_error(VDQN_ERR_SHAPE, "mlp_gemm: empty output");
  if (!(d->BN == 64 || d->BN == 128 || d->BN == 256)) return set_error(VDQN_ERR_SHAPE, "mlp_gemm: BN must be 64/128/256");
  if (d->split_m > 0 && pi != 0) return set_error(VDQN_ERR_ARG, "mlp_gemm: only the first problem of a launch may be dual");
  a = MlpArgs{};
  a.M = d->M; a.N = d->N; a.nseg = d->nseg; a.BN = d->BN;
  a.a_mn = d->a_mn ? 1 : 0; a.b_mn = d->b_mn ? 1 : 0;
  for (int s = 0; s < d->nseg; ++s) {
    const vdqn_mlp_operand& A = d->a[s];
    const vdqn_mlp_operand& B = d->b[s];
    if (A.ptr == nullptr || B.ptr == nullptr || d->K[s] < 1) return set_error(VDQN_ERR_ARG, "mlp_gemm: bad segment %d", s);
    if ((A.ld % 8) != 0 || (B.ld % 8) != 0 || ((reinterpret_cast<uintptr_t>(A.ptr) | reinterpret_cast<uintptr_t>(B.ptr)) & 15))
      return set_error(VDQN_ERR_SHAPE, "mlp_gemm: operands need 16-byte aligned rows (ld %% 8 == 0)");
    a.K[s] = d->K[s];
    int rc;
    // K-major: matrix [rows = M or N][cols = K];  MN-major: matrix [rows = K][cols = M or N]
    if (a.a_mn) rc = make_tiled_map_2d(&maps.a[pi][s], A.ptr, (uint64_t)A.cols, (uint64_t)d->K[s], 64, 64, 128, A.ld);
    else rc = make_tiled_map_2d(&maps.a[pi][s], A.ptr, (uint64_t)d->K[s], (uint64_t)A.rows, 64, 128, 128, A.ld);
    if (rc != VDQN_OK) return rc;
    for (int net = 0; net < 2; ++net) {
      if (net == 1 && d->split_m <= 0) continue;
      const vdqn_mlp_operand& Bn = net == 0 ? B : d->b2[s];
      CUtensorMap* mp = net == 0 ? &maps.b[pi][s] : &maps.b2[s];
      if (Bn.ptr == nullptr || (Bn.ld % 8) != 0) return set_error(VDQN_ERR_ARG, "mlp_gemm: bad second operand set");
      if (a.b_mn) rc = make_tiled_map_2d(mp, Bn.ptr, (uint64_t)Bn.cols, (uint64_t)d->K[s], 64, 64, 128, Bn.ld);
      else rc = make_tiled_map_2d(mp, Bn.ptr, (uint64_t)d->K[s], (uint64_t)Bn.rows, 64, (uint32_t)d->BN, 128, Bn.ld);
      if (rc != VDQN_OK) return rc;
    }
  }
  for (int s = d->nseg; s < kMlpMaxSeg; ++s) { maps.a[pi][s] = maps.a[pi][0]; maps.b[pi][s] = maps.b[pi][0]; }
  a.split_mtile = 0;
  if (d->split_m > 0) {
    if (d->split_m % kMlpBM != 0 || d->split_m >= d->M)
      return set_error(VDQN_ERR_SHAPE, "mlp_gemm: the second network's rows must start at a multiple of 128");
    a.split_mtile = d->split_m / kMlpBM;
  }
  a.bias = d->bias; a.bias2 = d->bias2; a.relu = d->relu;
  a.mask_f32 = d->mask_f32; a.mask_bf16 = static_cast<const __nv_bfloat16*>(d->mask_bf16); a.ldmask = d->ldmask;
  a.out_f32 = d->out_f32; a.ld_f32 = d->ld_f32; a.perm_c = d->perm_c; a.perm_p = d->perm_p;
  a.out_hi = static_cast<__nv_bfloat16*>(d->out_hi); a.out_lo = static_cast<__nv_bfloat16*>(d->out_lo); a.ld_hl = d->ld_hl;
  a.out_bf16 = static_cast<__nv_bfloat16*>(d->out_bf16); a.ld_bf16 = d->ld_bf16;
  a.colsum = d->colsum; a.colsum_mod = d->colsum_mod;
  if ((a.out_hi != nullptr) != (a.out_lo != nullptr)) return set_error(VDQN_ERR_ARG, "mlp_gemm: out_hi and out_lo go together");
  if (a.perm_c != 0 && (a.perm_p < 1 || d->N % (a.perm_c * a.perm_p) != 0))
    return set_error(VDQN_ERR_SHAPE, "mlp_gemm: permuted store needs whole (c, p) blocks");
  return VDQN_OK;
}

extern "C" int vdqn_mlp_gemm_grouped(const vdqn_mlp_gemm_desc* descs, int32_t n, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (descs == nullptr || n < 1 || n > kMlpMaxProb) return set_error(VDQN_ERR_ARG, "mlp_gemm: 1..4 problems per launch");
  DeviceInfo* dev = device_info();
  if (dev == nullptr) return VDQN_ERR_CUDA;
  MlpMaps maps;                 // host staging only: copied into the launch's parameter buffer
  MlpLaunch L{};
  L.nprob = n;
  int grid = 0, max_stage = 0;
  for (int i = 0; i < n; ++i) {
    const int rc = build_problem(&descs[i], i, maps, L.p[i]);
    if (rc != VDQN_OK) return rc;
    L.first_cta[i] = grid;
    grid += ((descs[i].M + kMlpBM - 1) / kMlpBM) * ((descs[i].N + descs[i].BN - 1) / descs[i].BN);
    const int sb = kMlpBM * kMlpKB * 2 + descs[i].BN * kMlpKB * 2;
    if (sb > max_stage) max_stage = sb;
  }
  for (int i = n; i <= kMlpMaxProb; ++i) L.first_cta[i] = grid;
  if (descs[0].split_m <= 0) for (int s = 0; s < kMlpMaxSeg; ++s) maps.b2[s] = maps.b[0][s];
  // few CTAs: the deepest pipeline that fits; more CTAs than SMs: leave room for two CTAs per SM
  const int budget = grid > dev->num_sms ? 100 * 1024 : 200 * 1024;
  int stages = budget / max_stage;
  if (stages > kMlpMaxStages) stages = kMlpMaxStages;
  if (stages < 2) stages = 2;
  L.stages = stages;
  const size_t smem = (size_t)stages * max_stage + 1024 + 256;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(mlp_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return set_error(VDQN_ERR_CUDA, "cudaFuncSetAttribute(mlp_gemm): %s", cudaGetErrorString(e));
    }
    attr_smem = smem;
  }
  launch_kernel(mlp_gemm_kernel, grid, 192, smem, stream, maps, L);
  VDQN_CHECK_LAUNCH("mlp_gemm");
  return VDQN_OK;
}

extern "C" int vdqn_mlp_gemm(const vdqn_mlp_gemm_desc* d, void* stream_v) {
  if (d == nullptr) return set_error(VDQN_ERR_ARG, "mlp_gemm: null descriptor");
  return vdqn_mlp_gemm_grouped(d, 1, stream_v);
}
