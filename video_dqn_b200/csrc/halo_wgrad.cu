// Weight gradient of the 64->64 3x3 convolutions (layer1) from one shared-memory copy of the
// input window -- the weight-gradient counterpart of halo_conv.cu.
//
//   g[co][(r,s)][ci] = sum_pix dy[pix][co] * x[pix + (r-1, s-1)][ci]
//
// computed transposed, D[(tap, ci)][co] = sum_pix X[pix + tap][ci] * dY[pix][co], so the 128 MMA rows
// are two filter taps x 64 input channels and N = 64 output channels carries no padding:
//   * A operand (MN-major, 128B swizzle): the x halo tile [18 x 10 pixels][64 ci] written once per
//     8x16-pixel tile by a tiled TMA box.  A tap is a start-address shift of (r*10+s) pixel rows;
//     the second tap of a pair is reached through the descriptor's leading-dimension offset
//     (LBO = shift difference), the 8-pixel K groups through SBO = one halo row (1280 B).
//   * B operand (MN-major): the dy tile [16 x 8 pixels][64 co], tiled TMA box.
//   * 5 tap pairs -> 5 accumulators of 64 TMEM columns, accumulated over ALL tiles a persistent CTA
//     owns; one epilogue at the end writes the CTA's partial [Cout=64][K=576], reduced over CTAs by
//     vdqn_wgrad_finalize (splits = number of CTAs).
// The im2col weight-gradient kernel moves 1.23 GB through L2 for these layers; this one ~0.3 GB.
#include "ptx.cuh"
#include "vdqn_internal.h"

namespace vdqn {

struct HwgradArgs {
  int N, H, W;
  int tiles_w, tiles_h, num_tiles;
  float* part;      // [gridDim.x][Cout = 64][K = 576]
};

struct HwCfg {
  static constexpr int TH = 16, TW = 8;
  static constexpr int HALO_W = 10, HALO_H = 18;
  static constexpr int X_BYTES = HALO_H * HALO_W * 128;          // 23040
  static constexpr int X_STAGE = 23 * 1024;
  static constexpr int DY_BYTES = TH * TW * 128;                 // 16384
  static constexpr int STAGE_BYTES = X_STAGE + DY_BYTES;         // 39 KB
  static constexpr int STAGES = 5;
  static constexpr int PAIRS = 5;                                // 9 taps -> 5 pairs (last half unused)
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

__global__ void __launch_bounds__(192, 1)
halo_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy,
                  const HwgradArgs a) {
  using Cfg = HwCfg;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * Cfg::STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 1);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDy);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
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

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
      const int tw = t % a.tiles_w;
      const int rr = t / a.tiles_w;
      const int th = rr % a.tiles_h;
      const int n = rr / a.tiles_h;
      const int h0 = th * Cfg::TH, w0 = tw * Cfg::TW;
      mbar_wait(empty_bar(stage), phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(full_bar(stage), Cfg::X_BYTES + Cfg::DY_BYTES);
        const uint32_t sX = smem_base + stage * Cfg::STAGE_BYTES;
        tma_load_4d(sX, &tmX, full_bar(stage), 0, w0 - 1, h0 - 1, n);
        tma_load_4d(sX + Cfg::X_STAGE, &tmDy, full_bar(stage), 0, w0, h0, n);
      }
      __syncwarp();
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);      // both operands MN-major
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x, ++it) {
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sX = smem_base + stage * Cfg::STAGE_BYTES;
        const uint32_t sDy = sX + Cfg::X_STAGE;
#pragma unroll
        for (int p = 0; p < Cfg::PAIRS; ++p) {
          // taps 2p and 2p+1 (the 10th "tap" of the last pair repeats tap 8; its rows are ignored)
          const int t1 = 2 * p, t2 = (2 * p + 1 < 9) ? 2 * p + 1 : 8;
          const int sh1 = (t1 / 3) * Cfg::HALO_W + (t1 % 3), sh2 = (t2 / 3) * Cfg::HALO_W + (t2 % 3);
          const uint32_t lbo = (sh2 > sh1) ? (uint32_t)(sh2 - sh1) * 128u : 128u;
#pragma unroll
          for (int k = 0; k < 8; ++k) {          // 16 pixels per step = 2 image rows x 8
            const uint64_t ad = make_smem_desc(sX + (sh1 + 2 * k * Cfg::HALO_W) * 128, lbo,
                                               Cfg::HALO_W * 128, kSwz128);
            const uint64_t bd = make_smem_desc(sDy + k * 2048, 8192, 1024, kSwz128);
            umma_f16(tmem_base + p * 64, ad, bd, idesc, (it | k) != 0);
          }
        }
        umma_commit(empty_bar(stage));
      }
      __syncwarp();
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;            // (tap parity, ci)
    const int half = row >> 6, ci = row & 63;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    float* dst = a.part + (long)blockIdx.x * 64 * 576;
#pragma unroll 1
    for (int p = 0; p < Cfg::PAIRS; ++p) {
      const int tap = 2 * p + half;
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + p * 64 + c0 + ((uint32_t)(quad * 32) << 16), raw);
        tmem_ld_wait();
        if (tap < 9) {
          // part[cta][co][tap*64 + ci]: for a fixed co the warp's 32 lanes (consecutive ci) write one
          // contiguous 128-byte line
          float* op = dst + (long)c0 * 576 + tap * 64 + ci;
#pragma unroll
          for (int j = 0; j < 32; ++j) op[(long)j * 576] = __uint_as_float(raw[j]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

bool halo_wgrad_supported(const vdqn_wgrad_desc* d) {
  return d->Cin == 64 && d->Cout == 64 && d->R == 3 && d->S == 3 && d->stride == 1 && d->dil == 1 &&
         d->pad_lo == 1 && d->pad_hi == 1 && d->ldy == 64;
}

int halo_wgrad_launch(const vdqn_wgrad_desc* d, cudaStream_t stream) {
  using Cfg = HwCfg;
  DeviceInfo* dev = device_info();
  if (dev == nullptr) return VDQN_ERR_CUDA;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(halo_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess)
      return set_error(VDQN_ERR_CUDA, "cudaFuncSetAttribute(halo_wgrad): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  HwgradArgs a{};
  a.N = d->N; a.H = d->H; a.W = d->W;
  a.tiles_w = (d->W + Cfg::TW - 1) / Cfg::TW;
  a.tiles_h = (d->H + Cfg::TH - 1) / Cfg::TH;
  a.num_tiles = d->N * a.tiles_w * a.tiles_h;
  a.part = d->part;
  if (d->splits < 1 || d->splits > a.num_tiles)
    return set_error(VDQN_ERR_ARG, "halo_wgrad: splits (= CTAs) must be in [1, %d]", a.num_tiles);
  CUtensorMap tmX, tmDy;
  int rc = make_tiled_map_nhwc(&tmX, d->x, d->N, d->H, d->W, 64, 64, Cfg::HALO_W, Cfg::HALO_H, 128);
  if (rc != VDQN_OK) return rc;
  rc = make_tiled_map_nhwc(&tmDy, d->dy, d->N, d->H, d->W, 64, 64, Cfg::TW, Cfg::TH, 128);
  if (rc != VDQN_OK) return rc;
  launch_kernel(halo_wgrad_kernel, d->splits, 192, Cfg::SMEM_BYTES, stream, tmX, tmDy, a);
  VDQN_CHECK_LAUNCH("halo_wgrad launch");
  return VDQN_OK;
}

}  // namespace vdqn
