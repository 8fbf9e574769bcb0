// Weight gradient of the packed stem (4x4 taps over the 16-channel space-to-depth input, Cout = 64)
// from one shared-memory copy of the input window -- the 32-byte-row sibling of halo_wgrad.cu.
//
//   D[(shift j, ci)][co] = sum_pix X[pix + (r, j)][ci] * dY[pix][co]        for each filter row r
//
// The 128 MMA rows are 8 horizontal pixel shifts x 16 packed channels: with the window stored as
// 32-byte pixel rows (32B swizzle) the MN-major A descriptor's leading-dimension offset is simply one
// pixel (32 B), so ONE descriptor covers eight consecutive shifts; shifts 0..3 are the filter's four
// column taps, 4..7 are by-products that are not stored (the tensor pipe is idle anyway: this kernel
// is bound by the 411 MB dy read).  Four accumulators (one per filter row, 64 TMEM columns each) stay
// in TMEM across every 8x16-pixel tile a persistent CTA owns; one epilogue at the end writes the
// CTA's partial [Cout=64][K=256], reduced over CTAs by vdqn_wgrad_finalize (kmap = 1).
// The window (19 x 11 pixels x 32 B) is gathered with 16-byte cp.async (zero-fill = padding);
// the dy tile [16 x 8 pixels][64 co] comes through a tiled TMA box.
#include "ptx.cuh"
#include "vdqn_internal.h"

#include <cuda_bf16.h>

namespace vdqn {

struct HwsArgs {
  int N, H, W;
  int tiles_w, tiles_h, num_tiles;
  const __nv_bfloat16* x;
  float* part;      // [gridDim.x][64][256]
};

struct HwsCfg {
  static constexpr int TH = 16, TW = 8;
  static constexpr int HALO_W = 11, HALO_H = 19;
  static constexpr int X_BYTES = HALO_H * HALO_W * 32;           // 6688
  static constexpr int X_STAGE = 7 * 1024;
  static constexpr int DY_BYTES = TH * TW * 128;                 // 16384
  static constexpr int STAGE_BYTES = X_STAGE + DY_BYTES;         // 23 KB
  static constexpr int STAGES = 8;
  static constexpr int DEPTH = 4;                                // cp.async tiles in flight
  static constexpr int TMEM_COLS = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

__global__ void __launch_bounds__(192, 1)
halo_wgrad_stem_kernel(const __grid_constant__ CUtensorMap tmDy, const HwsArgs a) {
  using Cfg = HwsCfg;
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
    tma_prefetch_desc(&tmDy);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 32);       // every producer lane arrives once its cp.async data has landed
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
    constexpr int CHUNKS = Cfg::HALO_H * Cfg::HALO_W * 2;
    int stage = 0, done_stage = 0, issued = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < a.num_tiles; t += gridDim.x) {
      const int tw = t % a.tiles_w;
      const int rr = t / a.tiles_w;
      const int th = rr % a.tiles_h;
      const int n = rr / a.tiles_h;
      const int h0 = th * Cfg::TH, w0 = tw * Cfg::TW;
      mbar_wait(empty_bar(stage), phase ^ 1);
      const uint32_t sX = smem_base + stage * Cfg::STAGE_BYTES;
      if (elect_one()) {
        // the dy tile's bytes are accounted on the same barrier (expect only; lanes arrive later)
        asm volatile("mbarrier.expect_tx.shared::cta.b64 [%0], %1;" ::"r"(full_bar(stage)), "r"(Cfg::DY_BYTES)
                     : "memory");
        tma_load_4d(sX + Cfg::X_STAGE, &tmDy, full_bar(stage), 0, w0, h0, n);
      }
      __syncwarp();
      const __nv_bfloat16* img = a.x + (long)n * a.H * a.W * 16;
      for (int c = lane; c < CHUNKS; c += 32) {
        const int p = c >> 1, hf = c & 1;
        const int hy = p / Cfg::HALO_W, wx = p - hy * Cfg::HALO_W;
        const int gh = h0 - 2 + hy, gw = w0 - 2 + wx;
        const bool inb = gh >= 0 && gh < a.H && gw >= 0 && gw < a.W;
        const __nv_bfloat16* src = inb ? img + ((long)gh * a.W + gw) * 16 + hf * 8 : a.x;
        const uint32_t dst = sX + p * 32 + ((uint32_t)(hf ^ ((p >> 2) & 1)) << 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(inb ? 16 : 0)
                     : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      ++issued;
      if (issued >= Cfg::DEPTH) {
        asm volatile("cp.async.wait_group %0;" ::"n"(Cfg::DEPTH - 1) : "memory");
        fence_proxy_async();
        mbar_arrive(full_bar(done_stage));
        if (++done_stage == Cfg::STAGES) done_stage = 0;
      }
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    fence_proxy_async();
    const int pending = issued < Cfg::DEPTH - 1 ? issued : Cfg::DEPTH - 1;
    for (int i = 0; i < pending; ++i) {
      mbar_arrive(full_bar(done_stage));
      if (++done_stage == Cfg::STAGES) done_stage = 0;
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
        for (int r = 0; r < 4; ++r) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {          // 16 pixels per step = 2 image rows x 8
            // A: 8 MN chunks (pixel shifts 0..7) of 16 channels, LBO = one pixel (32 B); the 8-pixel
            // K groups are one halo row (352 B) apart
            const uint64_t ad = make_smem_desc(sX + ((r + 2 * k) * Cfg::HALO_W) * 32, 32, Cfg::HALO_W * 32, kSwz32);
            const uint64_t bd = make_smem_desc(sDy + k * 2048, 8192, 1024, kSwz128);
            umma_f16(tmem_base + r * 64, ad, bd, idesc, (it | k) != 0);
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
    const int row = quad * 32 + lane;            // (pixel shift j, packed channel)
    const int j = row >> 4, c16 = row & 15;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    float* dst = a.part + (long)blockIdx.x * 64 * 256;
#pragma unroll 1
    for (int r = 0; r < 4; ++r) {
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + r * 64 + c0 + ((uint32_t)(quad * 32) << 16), raw);
        tmem_ld_wait();
        if (j < 4) {
          // part[cta][co][k], k = (r*4 + j)*16 + c16
          float* op = dst + (long)c0 * 256 + (r * 4 + j) * 16 + c16;
#pragma unroll
          for (int q = 0; q < 32; ++q) op[(long)q * 256] = __uint_as_float(raw[q]);
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

bool halo_wgrad_stem_supported(const vdqn_wgrad_desc* d) {
  return d->Cin == 16 && d->Cout == 64 && d->R == 4 && d->S == 4 && d->stride == 1 && d->dil == 1 &&
         d->pad_lo == 2 && d->pad_hi == 1 && d->ldy == 64;
}

int halo_wgrad_stem_launch(const vdqn_wgrad_desc* d, cudaStream_t stream) {
  using Cfg = HwsCfg;
  DeviceInfo* dev = device_info();
  if (dev == nullptr) return VDQN_ERR_CUDA;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(halo_wgrad_stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess)
      return set_error(VDQN_ERR_CUDA, "cudaFuncSetAttribute(halo_wgrad_stem): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  HwsArgs a{};
  a.N = d->N; a.H = d->H; a.W = d->W;
  a.tiles_w = (d->W + Cfg::TW - 1) / Cfg::TW;
  a.tiles_h = (d->H + Cfg::TH - 1) / Cfg::TH;
  a.num_tiles = d->N * a.tiles_w * a.tiles_h;
  a.x = static_cast<const __nv_bfloat16*>(d->x);
  a.part = d->part;
  if (d->splits < 1 || d->splits > a.num_tiles)
    return set_error(VDQN_ERR_ARG, "halo_wgrad_stem: splits (= CTAs) must be in [1, %d]", a.num_tiles);
  CUtensorMap tmDy;
  int rc = make_tiled_map_nhwc(&tmDy, d->dy, d->N, d->H, d->W, 64, 64, Cfg::TW, Cfg::TH, 128);
  if (rc != VDQN_OK) return rc;
  launch_kernel(halo_wgrad_stem_kernel, d->splits, 192, Cfg::SMEM_BYTES, stream, tmDy, a);
  VDQN_CHECK_LAUNCH("halo_wgrad_stem launch");
  return VDQN_OK;
}

}  // namespace vdqn
