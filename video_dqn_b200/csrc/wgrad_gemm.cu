// Convolution weight gradient on tcgen05 (sm_100a).
//   part[split][co][(r,s,ci)] = sum_{m in split} dy[m, co] * X[n, p*stride + r*dil - pad, q*stride + s*dil - pad, ci]
// i.e. a GEMM whose reduction dimension is the output pixel index m.  Both operands are therefore
// "MN-major" for the tensor core: the dy tile [64 pixels][64 co] and the im2col tile
// [64 pixels][CK ci] land in shared memory pixel-row by pixel-row (TMA tiled / TMA im2col, 128B or
// 32B swizzle) and are described to tcgen05.mma with MN-major descriptors -- no transposes.
// (replaces cudnn_convolution_backward_weight under loss.backward(), train_q_network.py:226.)
//
// Work unit = (pixel split, 128-row co tile, group of up to 256 K-columns); persistent CTAs walk
// the units; fp32 partial tiles go to a workspace that vdqn_wgrad_finalize reduces
// deterministically (no atomics).
#include "ptx.cuh"
#include "vdqn_internal.h"

#include <cuda_bf16.h>

namespace vdqn {

struct WgradArgs {
  int M_total, Ho, Wo, Cout, Ktot;
  int R, S, Cin, stride, dil, lower_h, lower_w;
  int splits, pix_per_split;     // pix_per_split is a multiple of 64
  int co_tiles, groups;
  float* part;
};

template <int CK>
struct WgradCfg {
  static constexpr int PIX = 64;                       // pixels (GEMM-K) per stage
  static constexpr int SLABS = (CK == 64) ? 4 : 16;    // B slabs per group -> N = 256
  static constexpr int BN = SLABS * CK;                // 256
  static constexpr int A_SLAB_BYTES = PIX * 64 * 2;    // dy slab: 64 pixels x 64 co
  static constexpr int A_BYTES = 2 * A_SLAB_BYTES;
  static constexpr int B_SLAB_BYTES = PIX * CK * 2;
  static constexpr int B_BYTES = SLABS * B_SLAB_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 48 KB
  static constexpr int STAGES = 4;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr uint64_t SWZ_B = (CK == 64) ? kSwz128 : kSwz32;
  static constexpr int B_ROW_BYTES = CK * 2;
};

template <int CK>
__global__ void __launch_bounds__(192, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmDy, const __grid_constant__ CUtensorMap tmX,
             const WgradArgs a) {
  using Cfg = WgradCfg<CK>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * Cfg::STAGES + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + i); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDy);
    tma_prefetch_desc(&tmX);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  const int units = a.splits * a.co_tiles * a.groups;
  const int cblks = a.Cin / CK;
  const int total_slabs = a.Ktot / CK;
  const int HoWo = a.Ho * a.Wo;
  const int co_slabs = a.Cout >= 128 ? 2 : 1;

  // unit -> (split, co_tile, group); group fastest so neighbours share the dy tile in L2
  auto decode = [&](int u, int& split, int& co_t, int& grp) {
    grp = u % a.groups;
    const int r = u / a.groups;
    co_t = r % a.co_tiles;
    split = r / a.co_tiles;
  };
  auto ksteps_of = [&](int split) {
    const int begin = split * a.pix_per_split;
    int end = begin + a.pix_per_split;
    if (end > a.M_total) end = a.M_total;
    return end > begin ? (end - begin + Cfg::PIX - 1) / Cfg::PIX : 0;
  };

  if (warp == 0) {
    // The producer is ONE thread: every integer division here is ~100 dependent cycles on the
    // critical path of a stage whose MMAs take 512.  Slab coordinates are therefore decoded once per
    // unit and the pixel coordinate advances incrementally (no division inside the k loop).
    int stage = 0;
    uint32_t phase = 0;
    const int dq = Cfg::PIX % a.Wo, dp = Cfg::PIX / a.Wo;
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int split, co_t, grp;
      decode(u, split, co_t, grp);
      const int nslab = min(Cfg::SLABS, total_slabs - grp * Cfg::SLABS);
      const int ksteps = ksteps_of(split);
      const uint32_t tx = co_slabs * Cfg::A_SLAB_BYTES + nslab * Cfg::B_SLAB_BYTES;
      int sl_c0[Cfg::SLABS];
      uint32_t sl_off[Cfg::SLABS];          // (offset_w, offset_h) packed 16:16
#pragma unroll
      for (int sl = 0; sl < Cfg::SLABS; ++sl) {
        const int j = grp * Cfg::SLABS + (sl < nslab ? sl : 0);
        const int tap = j / cblks;
        sl_c0[sl] = (j - tap * cblks) * CK;
        const int r = tap / a.S, sx = tap - r * a.S;
        sl_off[sl] = (uint32_t)(sx * a.dil) | ((uint32_t)(r * a.dil) << 16);
      }
      int m0 = split * a.pix_per_split;
      int img = m0 / HoWo;
      const int rem = m0 - img * HoWo;
      int p0 = rem / a.Wo, q0 = rem - p0 * a.Wo;
      for (int i = 0; i < ksteps; ++i) {
        const int cw = q0 * a.stride + a.lower_w, ch = p0 * a.stride + a.lower_h;
        mbar_wait(empty_bar(stage), phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(full_bar(stage), tx);
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
          tma_load_2d(sA, &tmDy, full_bar(stage), co_t * 128, m0);
          if (co_slabs == 2) tma_load_2d(sA + Cfg::A_SLAB_BYTES, &tmDy, full_bar(stage), co_t * 128 + 64, m0);
#pragma unroll
          for (int sl = 0; sl < Cfg::SLABS; ++sl) {
            if (sl < nslab)
              tma_load_im2col_4d(sB + sl * Cfg::B_SLAB_BYTES, &tmX, full_bar(stage), sl_c0[sl], cw, ch, img,
                                 (uint16_t)(sl_off[sl] & 0xffffu), (uint16_t)(sl_off[sl] >> 16));
          }
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        m0 += Cfg::PIX;
        q0 += dq; p0 += dp;
        if (q0 >= a.Wo) { q0 -= a.Wo; ++p0; }
        while (p0 >= a.Ho) { p0 -= a.Ho; ++img; }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
      int split, co_t, grp;
      decode(u, split, co_t, grp);
      const int nslab = min(Cfg::SLABS, total_slabs - grp * Cfg::SLABS);
      const int ksteps = ksteps_of(split);
      const uint32_t idesc = make_idesc_bf16(128, nslab * CK, 1, 1);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * Cfg::BN;
      for (int i = 0; i < ksteps; ++i) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < Cfg::PIX / 16; ++k) {
            // 16 pixels (GEMM-K) per instruction = two 8-row swizzle groups
            const uint64_t ad = make_smem_desc(sA + k * 16 * 128, Cfg::A_SLAB_BYTES, 8 * 128, kSwz128);
            const uint64_t bd = make_smem_desc(sB + k * 16 * Cfg::B_ROW_BYTES, Cfg::B_SLAB_BYTES,
                                               8 * Cfg::B_ROW_BYTES, Cfg::SWZ_B);
            umma_f16(d_tmem, ad, bd, idesc, (i | k) != 0);
          }
          umma_commit(empty_bar(stage));
          if (i == ksteps - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (ksteps == 0) {
        if (elect_one()) umma_commit(tfull_bar(acc));
        __syncwarp();
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    int it = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++it) {
      int split, co_t, grp;
      decode(u, split, co_t, grp);
      const int nslab = min(Cfg::SLABS, total_slabs - grp * Cfg::SLABS);
      const int ncols = nslab * CK;
      const int ksteps = ksteps_of(split);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int co = co_t * 128 + row;
      const bool valid = co < a.Cout;
      float* dst = a.part + ((long)split * a.Cout + co) * a.Ktot + (long)grp * Cfg::BN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + acc * Cfg::BN + c0 + ((uint32_t)(quad * 32) << 16), raw);
        tmem_ld_wait();
        if (valid) {
          float4* op = reinterpret_cast<float4*>(dst + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o;
            if (ksteps > 0) {
              o = make_float4(__uint_as_float(raw[4 * j]), __uint_as_float(raw[4 * j + 1]),
                              __uint_as_float(raw[4 * j + 2]), __uint_as_float(raw[4 * j + 3]));
            } else {
              o = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            op[j] = o;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int CK>
static int launch_wgrad(const CUtensorMap& tmDy, const CUtensorMap& tmX, const WgradArgs& a,
                        int num_sms, cudaStream_t stream) {
  using Cfg = WgradCfg<CK>;
  static bool attr_set = false;
  auto kfn = wgrad_kernel<CK>;
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess)
      return set_error(VDQN_ERR_CUDA, "cudaFuncSetAttribute(wgrad): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int units = a.splits * a.co_tiles * a.groups;
  const int grid = units < num_sms ? units : num_sms;
  launch_kernel(kfn, grid, 192, Cfg::SMEM_BYTES, stream, tmDy, tmX, a);
  VDQN_CHECK_LAUNCH("wgrad launch");
  return VDQN_OK;
}

}  // namespace vdqn

using namespace vdqn;

extern "C" int vdqn_conv_wgrad(const vdqn_wgrad_desc* d, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (d == nullptr) return set_error(VDQN_ERR_ARG, "conv_wgrad: null descriptor");
  const int CK = (d->Cin % 64 == 0) ? 64 : 16;
  if (d->Cin % CK != 0) return set_error(VDQN_ERR_SHAPE, "conv_wgrad: Cin=%d not a multiple of 16", d->Cin);
  if (d->Cout % 64 != 0 || (d->Cout > 64 && d->Cout % 128 != 0))
    return set_error(VDQN_ERR_SHAPE, "conv_wgrad: Cout=%d unsupported", d->Cout);
  if (d->splits < 1) return set_error(VDQN_ERR_ARG, "conv_wgrad: splits must be >= 1");
  if (d->algo == 2) {
    if (halo_wgrad_supported(d)) return halo_wgrad_launch(d, stream);
    if (halo_wgrad_stem_supported(d)) return halo_wgrad_stem_launch(d, stream);
    return set_error(VDQN_ERR_SHAPE, "conv_wgrad: halo algorithm requested for an unsupported shape");
  }
  if (d->ldy % 8 != 0) return set_error(VDQN_ERR_SHAPE, "conv_wgrad: ldy must be a multiple of 8");
  const int Ho = (d->H + d->pad_lo + d->pad_hi - (d->R - 1) * d->dil - 1) / d->stride + 1;
  const int Wo = (d->W + d->pad_lo + d->pad_hi - (d->S - 1) * d->dil - 1) / d->stride + 1;
  if (Ho <= 0 || Wo <= 0) return set_error(VDQN_ERR_SHAPE, "conv_wgrad: empty output");
  DeviceInfo* dev = device_info();
  if (dev == nullptr) return VDQN_ERR_CUDA;

  WgradArgs a{};
  a.M_total = d->N * Ho * Wo;
  a.Ho = Ho; a.Wo = Wo; a.Cout = d->Cout;
  a.Ktot = d->R * d->S * d->Cin;
  a.R = d->R; a.S = d->S; a.Cin = d->Cin; a.stride = d->stride; a.dil = d->dil;
  a.lower_h = -d->pad_lo; a.lower_w = -d->pad_lo;
  a.splits = d->splits;
  a.pix_per_split = ((a.M_total + d->splits - 1) / d->splits + 63) / 64 * 64;
  a.co_tiles = (d->Cout + 127) / 128;
  const int slabs = a.Ktot / CK;
  const int per_group = (CK == 64) ? 4 : 16;
  a.groups = (slabs + per_group - 1) / per_group;
  a.part = d->part;

  CUtensorMap tmDy, tmX;
  int rc = make_tiled_map_2d(&tmDy, d->dy, d->Cout, a.M_total, 64, 64, 128, d->ldy);
  if (rc != VDQN_OK) return rc;
  rc = make_im2col_map(&tmX, d->x, d->N, d->H, d->W, d->Cin, CK, 64, d->stride, -d->pad_lo,
                       -d->pad_lo, d->pad_hi - (d->R - 1) * d->dil, d->pad_hi - (d->S - 1) * d->dil,
                       CK == 64 ? 128 : 32);
  if (rc != VDQN_OK) return rc;
  const int sms = d->max_ctas > 0 && d->max_ctas < dev->num_sms ? d->max_ctas : dev->num_sms;
  return CK == 64 ? launch_wgrad<64>(tmDy, tmX, a, sms, stream)
                  : launch_wgrad<16>(tmDy, tmX, a, sms, stream);
}
