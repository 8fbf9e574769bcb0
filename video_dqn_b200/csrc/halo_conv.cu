// "Halo" convolution kernel for the wide, shallow layers (Cout = 64: the packed stem and the four
// 3x3 convolutions of layer1, forward and data-gradient).
//
// The im2col kernel (conv_gemm.cu) re-reads every input pixel R*S times from L2; for these layers
// (K per pixel small, N = 64) that makes them L2->SM bandwidth bound (ncu: 1.39 GB through the
// crossbar for a 103 MB input, ~9 TB/s, tensor pipe 19 %).  Here an output tile is a spatial
// rectangle of 8 (W) x 16 (H) pixels of one image and its input window -- (16+R-1) x (8+S-1)
// pixels x Cin -- is brought into shared memory ONCE by a tiled 4-D TMA box (out-of-image pixels
// zero-filled = the padding).  Every filter tap is then an A-operand descriptor that starts
// (r*(8+S-1)+s) pixel rows further into the same buffer, with the 8-pixel group pitch (SBO) equal
// to one halo row: tcgen05.mma applies the 128B/32B swizzle on absolute shared-memory addresses,
// so shifted starts and a non-1024B pitch read exactly what TMA wrote (verified on hardware by
// tools/umma_probe.cu).  The whole filter (<= 72 KB) stays resident in shared memory for the
// lifetime of the persistent CTA.  L2->SM traffic per layer1 conv drops 1.39 GB -> ~0.15 GB.
#include <cstdlib>

#include "epilogue.cuh"
#include "ptx.cuh"
#include "vdqn_internal.h"
#include "role_profile.cuh"

namespace vdqn {

struct HaloArgs {
  int N, H, W, Cout;
  int R, S, pad_lo;
  int tiles_w, tiles_h, num_tiles;
  int fast;            // staged TMA epilogue (bf16 compact output)
  int stages;          // 64-channel variants: halo-window stages of this launch (HaloCfg::stages_for)
  const __nv_bfloat16* x;   // input tensor (the 16-channel variant gathers it with cp.async)
  // dual-network launch: tiles [split_tile, num_tiles) (images >= split_n) use the second filter;
  // CTAs [0, split_cta) work on the first range, the rest on the second (split_cta == 0: off)
  int split_tile, split_cta;
  const float* shift2;
  int alias_from, alias_shift;     // images >= alias_from are read from image n - alias_shift (stem only)
  EpiArgs epi;
  // POOL variant (packed stem fused with max_pool2d(3, 2, 1), torchvision resnet.maxpool): pooled output
  // [N][H/2][W/2][64] bf16, arg-max slots (uint8, window scan order r*3+s) for images < idx_images
  __nv_bfloat16* pool_out;
  uint8_t* pool_idx;
  int idx_images;
};

// PAIR (CK = 64 only): the CTA is one half of a cta_group::2 pair.  The pair works on two tiles at a
// time with ONE M = 256 MMA per (tap, k-step); each CTA stages its own halo window and only HALF of
// the filter rows -- these kernels are bound by the shared-memory reads of the MMA operands
// (role profile: ~1900 cycles per tile against 1152 of tensor-pipe time), and the filter is a third
// of them.
// POOL (CK = 16 only): the stem's ReLU output never goes to HBM.  Tiles overlap by two rows -- a tile covers
// stem rows 14*th - 1 .. 14*th + 14, i.e. the rows of 7 pooled rows (row 15 of the tile is computed and not
// used: 14 % more MMA work) -- and a CTA walks the 14 tiles of a strip (image, th) from left to right, keeping
// the last column of a tile for the next one, so every 3x3/2 window is complete inside the CTA.  The eight
// epilogue warps put the tile (bf16, after shift + ReLU) into one of three shared buffers [16][9][64] (column 0
// = the column carried over), meet at a named barrier, and 224 threads form the 7 x 4 pooled pixels (and the
// arg-max slots the backward pass needs) from it.  HBM: 1176 MB written + 1234 MB re-read by the pooling
// kernels per 768 frames become 308 MB written.
template <int CK, bool PAIR = false, bool POOL = false>
struct HaloCfg {
  static constexpr int TH = 16, TW = 8, BN = 64;
  static constexpr int POOL_ROWS = 14;                  // stem rows a pooled tile advances by
  static constexpr int POOL_BUF_BYTES = 16 * 9 * 128;   // [16 rows][1 carried + 8 columns][64 ch] bf16
  static constexpr int ROW_BYTES = CK * 2;
  static constexpr uint64_t SWZ = (CK == 64) ? kSwz128 : kSwz32;
  static constexpr int R = (CK == 64) ? 3 : 4, S = R;   // filter size is fixed per variant
  static constexpr int MAX_TAPS = R * S;
  static constexpr int HALO_H = (CK == 64) ? 18 : 19, HALO_W = (CK == 64) ? 10 : 11;
  static constexpr int HALO_BYTES = HALO_H * HALO_W * ROW_BYTES;
  static constexpr int STAGE_BYTES = (HALO_BYTES + 1023) / 1024 * 1024;
  static constexpr int W_ROWS = PAIR ? BN / 2 : BN;               // filter rows staged by this CTA
  static constexpr int W_TILE_BYTES = W_ROWS * ROW_BYTES;          // one tap: W_ROWS x CK
  static constexpr int W_BYTES = MAX_TAPS * W_TILE_BYTES;          // 72 KB / 32 KB, resident
#ifndef VDQN_POOL_STAGES
#define VDQN_POOL_STAGES 15      // (sweep on B200: 12/4 480 us, 15/5 453 us, 18/6 453 us for 768 frames)
#endif
#ifndef VDQN_POOL_DEPTH
#define VDQN_POOL_DEPTH 5
#endif
  static constexpr int STAGES = (CK == 64) ? 3 : (POOL ? VDQN_POOL_STAGES : 12);
  // EIGHT epilogue warps, two per TMEM lane quadrant, each owning 32 of the 64 output columns with
  // private staging tiles [32 pixels][32 ch] (64-byte rows, 64B swizzle): OUT_BUFS x out,
  // 2 x (residual, mask).  The out tile is double-buffered wherever shared memory allows (not next to
  // the full 72 KB filter of the single-CTA 64-channel variant): with one buffer a warp waits ~700
  // cycles per tile for the previous TMA store to have read it.
  // (role profile: with four warps the epilogue was the longest role of every variant of this kernel
  //  -- 1900 cycles per tile against 950 of MMA issue for the stem -- a single warp per scheduler
  //  cannot hide its own TMEM / shared-memory / mbarrier latencies.)
  static constexpr int EPI_WARPS = 8;
  static constexpr int OUT_BUFS = (CK == 64 && !PAIR) ? 1 : 2;
  static constexpr int EPI_WARP_BYTES = (OUT_BUFS + 4) * 2048;
  static constexpr int EPI_BYTES = POOL ? 3 * POOL_BUF_BYTES : EPI_WARPS * EPI_WARP_BYTES;
  // 64-channel variants: shared memory is split AT LAUNCH between halo-window stages and epilogue staging (as
  // in the im2col kernel): a launch reserves the out tiles and two prefetch sets of only the inputs it has, the
  // rest becomes stages.  Role profile with the fixed 3 stages: the MMA warp waited 350 of 1750 cycles per tile
  // for windows while the producer waited 1200 for a free stage.  Pair: 6 / 5 / 4 stages for 0 / 1 / 2 inputs.
  static constexpr int MAX_STAGES = (CK == 64) ? 6 : STAGES;
  static constexpr int SMEM_LAYOUT = (CK == 64) ? 225 * 1024 : W_BYTES + STAGES * STAGE_BYTES + EPI_BYTES;
  __host__ __device__ static constexpr int epi_warp_bytes(int n_in) { return (OUT_BUFS + 2 * n_in) * 2048; }
  __host__ __device__ static constexpr int stages_for(int n_in) {
    return CK != 64 ? STAGES
           : (SMEM_LAYOUT - W_BYTES - EPI_WARPS * epi_warp_bytes(n_in)) / STAGE_BYTES > MAX_STAGES
               ? MAX_STAGES
               : (SMEM_LAYOUT - W_BYTES - EPI_WARPS * epi_warp_bytes(n_in)) / STAGE_BYTES;
  }
  // 12 warps = 384 threads: the register file then allows 168 registers per thread (416 threads were
  // compiled against the 512-thread limit of 128 and spilled)
  static constexpr int PROD_WARPS = 3;
  static constexpr int MMA_WARP = PROD_WARPS;
  // POOL: four more warps do the pooling, so a tile's pooling overlaps the next tile's TMEM drain
  static constexpr int POOL_WARPS = POOL ? 4 : 0;
  static constexpr int THREADS = (PROD_WARPS + 1 + EPI_WARPS + POOL_WARPS) * 32;
  static constexpr int SMEM_BYTES = SMEM_LAYOUT + 1024 + 512;
  // accumulator ring: with 64-column accumulators TMEM holds 4 of them, so the MMA warp can run up to
  // 4 tiles ahead of the epilogue and the mbarrier hand-off latencies (MMA -> epilogue -> MMA) overlap
  static constexpr int NACC = 4;
  static constexpr int TMEM_COLS = NACC * BN;
};

__device__ __forceinline__ void tma_load_tiled_4d(uint32_t dst, const void* tmap, uint32_t bar, int c,
                                                  int w, int h, int n) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n)
      : "memory");
}

template <int CK, bool PAIR, bool POOL = false>
__global__ void __launch_bounds__(HaloCfg<CK, PAIR, POOL>::THREADS, 1)
halo_conv_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                 const __grid_constant__ CUtensorMap tmMask, const HaloArgs a) {
  using Cfg = HaloCfg<CK, PAIR, POOL>;
  static_assert(!PAIR || CK == 64, "CTA pairs: 64-channel variant only");
  static_assert(!POOL || (CK == 16 && !PAIR), "fused pooling: packed stem only");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW = smem_base;
  const uint32_t sA0 = smem_base + Cfg::W_BYTES;
  const int nstages = CK == 64 ? a.stages : Cfg::STAGES;      // halo-window stages of this launch
  const uint32_t epi_base = sA0 + nstages * Cfg::STAGE_BYTES;
  const uint32_t bar_base = smem_base + Cfg::SMEM_LAYOUT;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::MAX_STAGES + s); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * Cfg::MAX_STAGES + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * Cfg::MAX_STAGES + Cfg::NACC + i); };
  const uint32_t w_bar = bar_base + 8u * (2 * Cfg::MAX_STAGES + 2 * Cfg::NACC);
  const uint32_t ld_bar0 = w_bar + 8u;                                  // one barrier per epilogue warp
  const uint32_t tmem_slot = w_bar + 8u * (1 + 2 * Cfg::EPI_WARPS);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();
  const int taps = a.R * a.S;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < Cfg::MAX_STAGES; ++s) {
      mbar_init(full_bar(s), CK == 16 ? 32 : 1);     // cp.async variant: every producer lane arrives
      mbar_init(empty_bar(s), 1);
    }
    for (int i = 0; i < Cfg::NACC; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), (PAIR ? 2 : 1) * Cfg::EPI_WARPS);
    }
    mbar_init(w_bar, 1);
    if constexpr (POOL) {
      // ld_bar0 + 8*i: i = 0..2 "buffer i holds a drained tile" (one arrive per drain warp),
      //                i = 3..5 "buffer i-3 has been pooled" (one arrive per pooling warp)
      for (int i = 0; i < 3; ++i) {
        mbar_init(ld_bar0 + 8u * i, Cfg::EPI_WARPS);
        mbar_init(ld_bar0 + 8u * (3 + i), Cfg::POOL_WARPS);
      }
    } else {
      for (int i = 0; i < 2 * Cfg::EPI_WARPS; ++i) mbar_init(ld_bar0 + 8u * i, 1);
    }
    fence_mbar_init();
  }
  // warps 0-2: producers (the cp.async variant uses all three, the TMA variant only warp 0),
  // warp 3: MMA issuer + TMEM owner, warps 4-11: epilogue (TMEM lane quadrant = warp & 3, column
  // half = (warp - 4) / 4)
  if (warp == Cfg::MMA_WARP) {
    if (PAIR) {
      tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();      // the peer's barriers must be initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  // PAIR: the schedule runs over CTA pairs and pairs of consecutive tiles; rank r takes tile 2*i + r
  // (the host guarantees an even number of tiles in every range)
  constexpr int MT = PAIR ? 2 : 1;
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int cta = (int)blockIdx.x / MT, ncta = (int)gridDim.x / MT, split_cta = a.split_cta / MT;
  const bool second = split_cta > 0 && cta >= split_cta;
  const int tile0 = (second ? a.split_tile : 0) + (cta - (second ? split_cta : 0)) * MT + rank;
  const int tstep = (split_cta > 0 ? (second ? ncta - split_cta : split_cta) : ncta) * MT;
  const int tile_end = (split_cta > 0 && !second) ? a.split_tile : a.num_tiles;
  const CUtensorMap* tmWp = second ? &tmW2 : &tmW;

  // Tile coordinates advance incrementally (a role is one warp walking its tiles in order; integer
  // divisions per tile would sit on its critical path): `step` tiles = (dn images, dth rows, dtw cols).
  struct TileIt { int tw, th, n, dtw, dth, dn; };
  auto tile_it = [&](int t, int step) {
    TileIt c;
    c.tw = t % a.tiles_w;
    const int r = t / a.tiles_w;
    c.th = r % a.tiles_h;
    c.n = r / a.tiles_h;
    c.dtw = step % a.tiles_w;
    const int rs = step / a.tiles_w;
    c.dth = rs % a.tiles_h;
    c.dn = rs / a.tiles_h;
    return c;
  };
  auto advance = [&](TileIt& c) {
    c.tw += c.dtw; c.th += c.dth; c.n += c.dn;
    if (c.tw >= a.tiles_w) { c.tw -= a.tiles_w; ++c.th; }
    if (c.th >= a.tiles_h) { c.th -= a.tiles_h; ++c.n; }
  };
  // POOL: the CTA's k-th tile.  Tiles are numbered strip-major (t = strip * tiles_w + tw, strip = n * tiles_h
  // + th) exactly as above, but a CTA takes WHOLE strips: its strips are strip0 + i * sstep, and its k-th
  // tile is column k % tiles_w of its (k / tiles_w)-th strip.
  const int strip0 = (second ? a.split_tile / a.tiles_w : 0) + (cta - (second ? split_cta : 0));
  const int sstep = tstep;                       // CTAs of this range
  const int strip_end = tile_end / a.tiles_w;
  const int my_strips = strip0 < strip_end ? (strip_end - strip0 + sstep - 1) / sstep : 0;
  const int my_tiles = POOL ? my_strips * a.tiles_w : 0;
  auto pool_tile = [&](int k, int& n, int& th, int& tw) {
    const int si = k / a.tiles_w;
    tw = k - si * a.tiles_w;
    const int strip = strip0 + si * sstep;
    n = strip / a.tiles_h;
    th = strip - n * a.tiles_h;
  };
  // the same walk without divisions (two per tile sat on the critical path of every role: ~350 cycles of a
  // ~1100-cycle tile): `step` tiles further along this CTA's tile sequence
  struct PoolIt { int n, th, tw, dn, dth; };
  auto pool_it = [&](int k) {
    PoolIt c;
    pool_tile(k, c.n, c.th, c.tw);
    c.dn = sstep / a.tiles_h;
    c.dth = sstep - c.dn * a.tiles_h;
    return c;
  };
  auto pool_next_strip = [&](PoolIt& c) {
    c.n += c.dn; c.th += c.dth;
    if (c.th >= a.tiles_h) { c.th -= a.tiles_h; ++c.n; }
  };
  auto pool_advance = [&](PoolIt& c, int step) {       // step < tiles_w
    c.tw += step;
    if (c.tw >= a.tiles_w) { c.tw -= a.tiles_w; pool_next_strip(c); }
  };

  if (warp < Cfg::PROD_WARPS) {
    // the filter: one [64 x CK] tile per tap, resident for the whole kernel
    if (warp == 0) {
      if (elect_one()) {
        if (PAIR) {
          // both CTAs' filter halves signal the leader's barrier (its MMAs read both)
          if (rank == 0) mbar_expect_tx(w_bar, 2 * taps * Cfg::W_TILE_BYTES);
          const uint32_t wb = mapa_u32(w_bar, 0);
          for (int j = 0; j < taps; ++j)
            tma_load_2d_pair(sW + j * Cfg::W_TILE_BYTES, tmWp, wb, j * CK, rank * Cfg::W_ROWS);
        } else {
          mbar_expect_tx(w_bar, taps * Cfg::W_TILE_BYTES);
          for (int j = 0; j < taps; ++j) tma_load_2d(sW + j * Cfg::W_TILE_BYTES, tmWp, w_bar, j * CK, 0);
        }
      }
      __syncwarp();
    }
    if constexpr (CK == 16) {
      // 32-byte pixel rows are slow through the TMA unit (one request per row) and a single warp
      // gathering them with cp.async is issue-bound (measured: 133 us of the 190 us stem kernel with
      // everything else switched off), so THREE producer warps share the work: warp p gathers the
      // 19 x 11 x 32 B windows of the CTA's tiles p, p+3, ... with 16-byte cp.async (zero-fill =
      // padding) into the 32B-swizzled layout the MMA descriptors expect, DEPTH tiles in flight each.
      // each warp waits for the tile it issued DEPTH-1 iterations ago: with too few tiles in flight the
      // loop period is the memory latency (measured: 84 us floor at DEPTH 2); 3 warps x 4 tiles in flight of the 16 stages
      constexpr int DEPTH = POOL ? VDQN_POOL_DEPTH : 4;
      constexpr int CHUNKS = Cfg::HALO_H * Cfg::HALO_W * 2;
      constexpr int PER_LANE = (CHUNKS + 31) / 32;
      // the chunk -> (window pixel, smem offset) map is the same for every tile: keep it in registers
      uint32_t dst_off[PER_LANE], hywx[PER_LANE];
#pragma unroll
      for (int i = 0; i < PER_LANE; ++i) {
        const int c = lane + 32 * i;
        const int p = c >> 1, hf = c & 1;
        const int hy = p / Cfg::HALO_W, wx = p - hy * Cfg::HALO_W;
        dst_off[i] = (uint32_t)(p * 32) + ((uint32_t)(hf ^ ((p >> 2) & 1)) << 4);
        hywx[i] = (c < CHUNKS) ? ((uint32_t)hy | ((uint32_t)wx << 8) | ((uint32_t)hf << 16)) : 0xffffffffu;
      }
      int issued = 0, done_k = warp;
      int k = warp;
      PROF_BEGIN
      constexpr int NP = Cfg::PROD_WARPS;
      TileIt ti = tile_it(tile0 + warp * tstep, NP * tstep);
      PoolIt pit = pool_it(POOL ? warp : 0);
      for (int t = tile0 + warp * tstep; POOL ? k < my_tiles : t < tile_end; t += NP * tstep, k += NP, advance(ti)) {
        int n = ti.n, h0 = ti.th * Cfg::TH, w0 = ti.tw * Cfg::TW;
        if constexpr (POOL) {
          n = pit.n;
          h0 = pit.th * Cfg::POOL_ROWS - 1;
          w0 = pit.tw * Cfg::TW;
          pool_advance(pit, NP);
        }
        PROF_TILE
        const int stage = k % Cfg::STAGES;
        PROF_WAIT_A(mbar_wait(empty_bar(stage), (((uint32_t)(k / Cfg::STAGES)) & 1u) ^ 1u))
        const uint32_t base = sA0 + stage * Cfg::STAGE_BYTES;
        const int n_src = (a.alias_from > 0 && n >= a.alias_from) ? n - a.alias_shift : n;
        const __nv_bfloat16* img = a.x + (long)n_src * a.H * a.W * 16;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) {
          if (hywx[i] == 0xffffffffu) continue;
          const int gh = h0 - a.pad_lo + (int)(hywx[i] & 0xff), gw = w0 - a.pad_lo + (int)((hywx[i] >> 8) & 0xff);
          const bool inb = (unsigned)gh < (unsigned)a.H && (unsigned)gw < (unsigned)a.W;
          const __nv_bfloat16* src = inb ? img + ((long)gh * a.W + gw) * 16 + ((hywx[i] >> 16) & 1) * 8 : a.x;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(base + dst_off[i]), "l"(src),
                       "r"(inb ? 16 : 0)
                       : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        ++issued;
        if (issued >= DEPTH) {          // this warp's oldest in-flight tile has landed: publish it
          PROF_WAIT_B(asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory"))
          fence_proxy_async();
          mbar_arrive(full_bar(done_k % Cfg::STAGES));
          done_k += Cfg::PROD_WARPS;
        }
      }
      PROF_END(warp)
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async();
      const int pending = issued < DEPTH - 1 ? issued : DEPTH - 1;
      for (int i = 0; i < pending; ++i) {
        mbar_arrive(full_bar(done_k % Cfg::STAGES));
        done_k += Cfg::PROD_WARPS;
      }
    } else if (warp == 0) {
      int stage = 0;
      uint32_t phase = 0;
      PROF_BEGIN
      TileIt ti = tile_it(tile0, tstep);
      for (int t = tile0; t < tile_end; t += tstep, advance(ti)) {
        const int n = ti.n, h0 = ti.th * Cfg::TH, w0 = ti.tw * Cfg::TW;
        PROF_TILE
        PROF_WAIT_A(mbar_wait(empty_bar(stage), phase ^ 1))
        if (elect_one()) {
          if (PAIR) {
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * Cfg::HALO_BYTES);
            tma_load_4d_pair(sA0 + stage * Cfg::STAGE_BYTES, &tmX, mapa_u32(full_bar(stage), 0), 0, w0 - a.pad_lo,
                             h0 - a.pad_lo, n);
          } else {
            mbar_expect_tx(full_bar(stage), Cfg::HALO_BYTES);
            tma_load_tiled_4d(sA0 + stage * Cfg::STAGE_BYTES, &tmX, full_bar(stage), 0, w0 - a.pad_lo,
                              h0 - a.pad_lo, n);
          }
        }
        __syncwarp();
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
      PROF_END(0)
    }
  } else if (warp == Cfg::MMA_WARP) {
    constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, Cfg::BN, 0, 0);
    if (PAIR && rank != 0) goto teardown;       // only the leader CTA issues MMAs
    mbar_wait(w_bar, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    PROF_BEGIN
    for (int t = tile0; POOL ? it < my_tiles : t < tile_end; t += tstep, ++it) {
      const int acc = it % Cfg::NACC;
      const uint32_t acc_phase = (it / Cfg::NACC) & 1;
      PROF_TILE
      PROF_WAIT_A(mbar_wait(tempty_bar(acc), acc_phase ^ 1))
      PROF_WAIT_B(mbar_wait(full_bar(stage), phase))
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + acc * Cfg::BN;
        const uint32_t sA = sA0 + stage * Cfg::STAGE_BYTES;
        // descriptors advance by compile-time constants: (tap row, tap col, k16 step)
        const uint64_t a0 = make_smem_desc(sA, 16, Cfg::HALO_W * Cfg::ROW_BYTES, Cfg::SWZ);
        const uint64_t b0 = make_smem_desc(sW, 16, 8 * Cfg::ROW_BYTES, Cfg::SWZ);
        if (!(a.epi.flags & 16))        // flag 16: debug, skip the MMAs (bottleneck probing)
#pragma unroll
        for (int r = 0; r < Cfg::R; ++r) {
#pragma unroll
          for (int s = 0; s < Cfg::S; ++s) {
#pragma unroll
            for (int k = 0; k < CK / 16; ++k) {
              const uint64_t ad = a0 + (uint64_t)(((r * Cfg::HALO_W + s) * Cfg::ROW_BYTES + k * 32) >> 4);
              const uint64_t bd = b0 + (uint64_t)(((r * Cfg::S + s) * Cfg::W_TILE_BYTES + k * 32) >> 4);
              if (PAIR) umma_f16_pair(d_tmem, ad, bd, idesc, (r | s | k) != 0);
              else umma_f16(d_tmem, ad, bd, idesc, (r | s | k) != 0);
            }
          }
        }
        if (PAIR) {
          umma_commit_pair(empty_bar(stage));
          umma_commit_pair(tfull_bar(acc));
        } else {
          umma_commit(empty_bar(stage));
          umma_commit(tfull_bar(acc));
        }
      }
      __syncwarp();
      if (++stage == nstages) { stage = 0; phase ^= 1; }
    }
    PROF_END(4)
  } else if (POOL && warp >= Cfg::MMA_WARP + 1 + Cfg::EPI_WARPS) {
    // ---------------------------------------------------------------- pooling warps (POOL variant)
    if constexpr (POOL) {
      const int ptid = (warp - (Cfg::MMA_WARP + 1 + Cfg::EPI_WARPS)) * 32 + lane;      // 0..127
      const int Hp = a.H / 2, Wp = a.W / 2;
      PROF_BEGIN
      PoolIt pit = pool_it(0);
      int buf = 0;
      uint32_t ready_phase = 0;
      for (int k = 0; k < my_tiles; ++k, buf = (buf == 2 ? 0 : buf + 1)) {
        PROF_TILE
        const int n = pit.n, th = pit.th, tw = pit.tw;
        pool_advance(pit, 1);
        const uint32_t sbuf = epi_base + (uint32_t)buf * Cfg::POOL_BUF_BYTES;
        const uint8_t* sbuf_p = smem_raw + (sbuf - smem_u32(smem_raw));
        const bool want_idx = a.pool_idx != nullptr && n < a.idx_images;
        PROF_WAIT_A(mbar_wait(ld_bar0 + 8u * buf, ready_phase))
        if (buf == 2) ready_phase ^= 1u;
        if (a.epi.flags & 8) {             // debug: no pooling work (bottleneck probing)
          __syncwarp();
          if (lane == 0) mbar_arrive(ld_bar0 + 8u * (3 + buf));
          continue;
        }
        // Separable 3x3 maximum: thread (row group rg, pooled column pc, channel group pk) reduces the three
        // columns of each of its (up to) five tile rows once, then combines rows {0,1,2} and {2,3,4} into its two
        // pooled rows -- 15 loads and ~half the compares of nine taps per output, and all 28 x 8 outputs of the
        // tile in ONE pass of the four warps.  Slots keep torch's choice: first maximum in (row, column) order.
        const int rg = ptid >> 5, pc = (ptid >> 3) & 3, pk = ptid & 7;
        const int nrow = rg < 3 ? 5 : 3;
        // all fifteen loads first (plain loads: the compiler keeps them in flight together)
        uint4 ld[5][3];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
#pragma unroll
          for (int sx = 0; sx < 3; ++sx) {
            const int gg = 4 * rg + (r < nrow ? r : 0), cc = 2 * pc + sx;
            ld[r][sx] = *reinterpret_cast<const uint4*>(sbuf_p + (gg * 9 + cc) * 128 + ((pk ^ ((cc + 7) & 7)) << 4));
          }
        }
        uint32_t best[2][4], slot[2][4];
        if (!want_idx) {
          // values only (two thirds of the frames of a step: s' is never back-propagated through).  Every value
          // is >= +0 or the -inf of the padding, so bf16 order is the order of the bit patterns as signed
          // 16-bit integers and a three-way integer maximum does two columns / rows per instruction.
          uint32_t hv[5][4];
#pragma unroll
          for (int r = 0; r < 5; ++r) {
            const uint32_t v0[4] = {ld[r][0].x, ld[r][0].y, ld[r][0].z, ld[r][0].w};
            const uint32_t v1[4] = {ld[r][1].x, ld[r][1].y, ld[r][1].z, ld[r][1].w};
            const uint32_t v2[4] = {ld[r][2].x, ld[r][2].y, ld[r][2].z, ld[r][2].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) hv[r][e] = __vimax3_s16x2(v0[e], v1[e], v2[e]);
          }
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            best[0][e] = __vimax3_s16x2(hv[0][e], hv[1][e], hv[2][e]);
            best[1][e] = __vimax3_s16x2(hv[2][e], hv[3][e], hv[4][e]);
            slot[0][e] = slot[1][e] = 0u;
          }
        } else {
        uint32_t hv[5][4], hs[5][4];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          if (r < nrow) {
            const uint4 c3[3] = {ld[r][0], ld[r][1], ld[r][2]};
            const uint32_t v0[4] = {c3[0].x, c3[0].y, c3[0].z, c3[0].w};
            const uint32_t v1[4] = {c3[1].x, c3[1].y, c3[1].z, c3[1].w};
            const uint32_t v2[4] = {c3[2].x, c3[2].y, c3[2].z, c3[2].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __nv_bfloat162 a0 = *reinterpret_cast<const __nv_bfloat162*>(&v0[e]);
              const __nv_bfloat162 a1 = *reinterpret_cast<const __nv_bfloat162*>(&v1[e]);
              const __nv_bfloat162 a2 = *reinterpret_cast<const __nv_bfloat162*>(&v2[e]);
              const __nv_bfloat162 m01 = __hmax2(a0, a1);
              const __nv_bfloat162 m012 = __hmax2(m01, a2);
              hv[r][e] = *reinterpret_cast<const uint32_t*>(&m012);
              const uint32_t g1 = __hgt2_mask(a1, a0);                 // column 1 beats column 0
              const uint32_t g2 = __hgt2_mask(a2, m01);                // column 2 beats both
              hs[r][e] = (g2 & 0x00020002u) | (~g2 & g1 & 0x00010001u);
            }
          }
        }
#pragma unroll
        for (int o = 0; o < 2; ++o) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 b0 = *reinterpret_cast<const __nv_bfloat162*>(&hv[2 * o][e]);
            if (2 * o + 2 < nrow) {
              const __nv_bfloat162 b1 = *reinterpret_cast<const __nv_bfloat162*>(&hv[2 * o + 1][e]);
              const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&hv[2 * o + 2][e]);
              const __nv_bfloat162 m01 = __hmax2(b0, b1);
              const __nv_bfloat162 m012 = __hmax2(m01, b2);
              best[o][e] = *reinterpret_cast<const uint32_t*>(&m012);
              const uint32_t g1 = __hgt2_mask(b1, b0);
              const uint32_t g2 = __hgt2_mask(b2, m01);
              const uint32_t s01 = (g1 & (hs[2 * o + 1][e] + 0x00030003u)) | (~g1 & hs[2 * o][e]);
              slot[o][e] = (g2 & (hs[2 * o + 2][e] + 0x00060006u)) | (~g2 & s01);
            } else {
              best[o][e] = 0u; slot[o][e] = 0u;
            }
          }
        }
        }
        __syncwarp();                                   // every lane's reads of the buffer are complete
        if (lane == 0) mbar_arrive(ld_bar0 + 8u * (3 + buf));
        PROF_WAIT_B(;)
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const int pr = 2 * rg + o;
          if (pr < 7 && !(a.epi.flags & 64)) {       // flag 64: debug, no global stores
            const long off = ((((long)n * Hp + th * (Cfg::POOL_ROWS / 2) + pr) * Wp + tw * (Cfg::TW / 2) + pc) * 64 + pk * 8);
            *reinterpret_cast<uint4*>(a.pool_out + off) = make_uint4(best[o][0], best[o][1], best[o][2], best[o][3]);
            if (want_idx) {
              uint2 ip;
              ip.x = __byte_perm(slot[o][0], slot[o][1], 0x6420);
              ip.y = __byte_perm(slot[o][2], slot[o][3], 0x6420);
              *reinterpret_cast<uint2*>(a.pool_idx + off) = ip;
            }
          }
        }
      }
      PROF_END(13 + (warp & 1))
    }
  } else {
    const int ew = warp - (Cfg::MMA_WARP + 1);   // 0..7
    const int quad = warp & 3;                // TMEM lane quadrant this warp may touch
    const int half = ew >> 2;                 // its 32 output columns: [half*32, half*32 + 32)
    const int row = quad * 32 + lane;
    const int g = row >> 3, j = row & 7;
    // per-warp staging of this launch: OUT_BUFS out tiles, then two prefetch sets of [residual][mask] with absent
    // inputs taking no room (64-channel variants; the stem keeps the fixed layout)
    const int n_in = CK == 64 ? (a.fast ? (a.epi.residual != nullptr ? 1 : 0) + (a.epi.mask_src != nullptr ? 1 : 0) : 0) : 2;
    const uint32_t set_bytes = (uint32_t)n_in * 2048u;
    const uint32_t mask_off = (CK != 64 || a.epi.residual != nullptr) ? 2048u : 0u;
    const uint32_t stg_out0 = epi_base + ew * (uint32_t)Cfg::epi_warp_bytes(n_in);
    const uint32_t stg_in0 = stg_out0 + Cfg::OUT_BUFS * 2048;   // input set s at + s * set_bytes: residual, then mask
    const uint32_t ld_bar = ld_bar0 + 16u * ew;        // two barriers, one per input set
    const int c0 = half * 32;
    EpiArgs epi = a.epi;
    if (second) epi.shift = a.shift2;
    // accumulator drained: tell the MMA issuer (PAIR: the leader CTA's barrier, counted over both CTAs)
    auto release_acc = [&](int acc_i) {
      if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(tempty_bar(acc_i), 0));
      else mbar_arrive(tempty_bar(acc_i));
    };
    // The tile loop is instantiated once per combination of optional epilogue steps and the launch
    // picks its specialisation (epilogue.cuh, epi_dispatch).
    auto epi_loop = [&](auto mode_tag) {
      constexpr int EPI = decltype(mode_tag)::value;
      constexpr bool COLSUM = (EPI & EPI_HAS_COLSUM) != 0;
      const bool has_res = (EPI & EPI_HAS_RES) && epi.residual != nullptr;
      const bool has_mask = (EPI & EPI_HAS_MASK) && epi.mask_src != nullptr;
      const bool has_in = has_res || has_mask;
      float csum = 0.f;                    // per-lane column sum (channel c0 + lane) over all tiles
      float row_acc[COLSUM ? 32 : 1];      // staged path: this pixel row's running sums, reduced once at the end
#pragma unroll
      for (int i = 0; i < (COLSUM ? 32 : 1); ++i) row_acc[i] = 0.f;
      // this warp's 32 shift values stay in registers for the whole kernel
      float shift_r[(EPI & EPI_HAS_SHIFT_RELU) ? 32 : 1];
      if constexpr ((EPI & EPI_HAS_SHIFT_RELU) != 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) shift_r[i] = epi.shift != nullptr ? __ldg(epi.shift + c0 + i) : 0.f;
      }
      int it = 0;
      PROF_BEGIN
      TileIt ti = tile_it(tile0, tstep);
      // this warp's 32 pixels = image rows h0+4*quad .. +3, 8 pixels each: a [1][4][8][32] TMA box.
      // Residual / mask tiles are prefetched TWO TILES AHEAD into two staging sets: a set is refilled
      // as soon as its tile's math is done, for the tile after the next one (one tile of lead was not
      // enough -- the role profile showed 600-800 cycles per tile waiting for these loads).
      auto issue_inputs = [&](const TileIt& c, int set) {
        if (elect_one()) {
          const uint32_t bar = ld_bar + 8u * set, dst = stg_in0 + (uint32_t)set * set_bytes;
          mbar_expect_tx(bar, (has_res ? 2048u : 0u) + (has_mask ? 2048u : 0u));
          if (has_res) tma_load_4d(dst, &tmRes, bar, c0, c.tw * Cfg::TW, c.th * Cfg::TH + 4 * quad, c.n);
          if (has_mask) tma_load_4d(dst + mask_off, &tmMask, bar, c0, c.tw * Cfg::TW, c.th * Cfg::TH + 4 * quad, c.n);
        }
        __syncwarp();
      };
      TileIt tp = ti;                      // prefetch cursor: two tiles ahead of the loop
      if (a.fast && has_in) {
        if (tile0 < tile_end) issue_inputs(tp, 0);
        advance(tp);
        if (tile0 + tstep < tile_end) issue_inputs(tp, 1);
        advance(tp);
      }
      for (int t = tile0; t < tile_end; t += tstep, ++it) {
        const int n = ti.n, h0 = ti.th * Cfg::TH, w0 = ti.tw * Cfg::TW;
        advance(ti);                         // ti = the NEXT tile from here on
        PROF_TILE
        const int acc = it % Cfg::NACC;
        const uint32_t acc_phase = (it / Cfg::NACC) & 1;
        const int h = h0 + g, w = w0 + j;
        const bool valid = h < a.H && w < a.W;
        PROF_WAIT_A(mbar_wait(tfull_bar(acc), acc_phase))
        tc_fence_after();
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + acc * Cfg::BN + c0 + ((uint32_t)(quad * 32) << 16), raw);
        if (a.fast) {
          const uint32_t stg_out = stg_out0 + (uint32_t)(it % Cfg::OUT_BUFS) * 2048u;
          const int set = it & 1;
          const uint32_t stg_res = stg_in0 + (uint32_t)set * set_bytes, stg_mask = stg_res + mask_off;
          PROF_WAIT_B(if (elect_one()) tma_store_wait_read<Cfg::OUT_BUFS - 1>(); __syncwarp())   // this out tile's last store has been read
          tmem_ld_wait();
          if (has_in) PROF_WAIT_B(mbar_wait(ld_bar + 8u * set, (uint32_t)(it >> 1) & 1u))
          if (!(epi.flags & 32))          // flag 32: debug, skip the epilogue math (bottleneck probing)
            csum += epilogue_half_staged<64, EPI>(epi, raw, valid, c0, 0, lane, stg_out, stg_res, stg_mask,
                                                  COLSUM ? row_acc : nullptr,
                                                  (EPI & EPI_HAS_SHIFT_RELU) ? shift_r : nullptr);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(acc);
          if (has_in) {
            if (t + 2 * tstep < tile_end) issue_inputs(tp, set);   // this set is free again
            advance(tp);
          }
          fence_proxy_async();
          __syncwarp();
          if (elect_one() && !(epi.flags & 8)) {      // flag 8: debug, skip the store (bottleneck probing)
            tma_store_4d(&tmOut, stg_out, c0, w0, h0 + 4 * quad, n);
            tma_store_commit();
          }
          __syncwarp();
          continue;
        }
        tmem_ld_wait();
        const long opix = ((long)n * a.H + h) * a.W + w;
        csum += epilogue_chunk(epi, raw, valid, opix, opix, 0, c0, lane);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(acc);
      }
      PROF_END(5 + ew)
      if (COLSUM && a.epi.colsum != nullptr) {
        if constexpr (COLSUM) {
          if (a.fast) csum += warp_transpose_reduce(row_acc, lane);
        }
        atomicAdd(a.epi.colsum + c0 + lane, csum);
      }
    };
    if constexpr (POOL) {
      // ---------------------------------------------------------------- stem drain (the pooling warps take it from here)
      float shift_r[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) shift_r[i] = epi.shift != nullptr ? __ldg(epi.shift + c0 + i) : 0.f;
      PROF_BEGIN
      PoolIt pit = pool_it(0);
      // buf = k % 3: this tile's buffer; nbuf = (k + 1) % 3 = (k - 2) % 3: the next tile's, which is also the one
      // tile k-2 was pooled from
      int buf = 0, nbuf = 1;
      uint32_t done_phase = 0;
      if (j == 7) {            // column -1 of the CTA's first tile (a strip start): padding
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          sts128(epi_base + (uint32_t)(g * 9 * 128) + ((uint32_t)((half * 4 + q4) ^ 7) << 4),
                 make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u));
      }
      for (int k = 0; k < my_tiles; ++k, buf = nbuf, nbuf = (nbuf == 2 ? 0 : nbuf + 1)) {
        PROF_TILE
        const int th = pit.th;
        const bool last_of_strip = pit.tw == a.tiles_w - 1;
        pool_advance(pit, 1);
        const int acc = k % Cfg::NACC;
        const uint32_t acc_phase = (k / Cfg::NACC) & 1;
        PROF_WAIT_A(mbar_wait(tfull_bar(acc), acc_phase))
        tc_fence_after();
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + acc * Cfg::BN + c0 + ((uint32_t)(quad * 32) << 16), raw);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(acc);
        // this thread's pixel: tile row g = stem row 14*th - 1 + g (rows outside the image are padding: -inf)
        const int hs = th * Cfg::POOL_ROWS - 1 + g;
        const bool rvalid = (unsigned)hs < (unsigned)a.H;
        uint4 pkd[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pkd[q4]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float x0 = fmaxf(__uint_as_float(raw[8 * q4 + 2 * e]) + shift_r[8 * q4 + 2 * e], 0.f);
            const float x1 = fmaxf(__uint_as_float(raw[8 * q4 + 2 * e + 1]) + shift_r[8 * q4 + 2 * e + 1], 0.f);
            h2[e] = __floats2bfloat162_rn(x0, x1);
          }
          if (!rvalid) pkd[q4] = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
        }
        // tile k writes buffer k % 3 (last read by the pooling of tile k-3) and column 0 of buffer (k+1) % 3
        // (last read by the pooling of tile k-2): wait for tile k-2 to have been pooled
        if (k >= 2) {
          PROF_WAIT_B(mbar_wait(ld_bar0 + 8u * (3 + nbuf), done_phase))
          if (nbuf == 2) done_phase ^= 1u;
        }
        if (epi.flags & 32) {              // debug: drain without the shared-memory stores (bottleneck probing)
          __syncwarp();
          if (lane == 0) mbar_arrive(ld_bar0 + 8u * buf);
          continue;
        }
        // buffer [g][c][8 chunks of 16 B], chunk position XOR-ed with (c - 1) & 7 so that the eight pixels of a
        // tile row (128 B apart) do not share banks
        const uint32_t sbuf = epi_base + (uint32_t)buf * Cfg::POOL_BUF_BYTES;
        const uint32_t prow = sbuf + (uint32_t)((g * 9 + j + 1) * 128);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) sts128(prow + ((uint32_t)((half * 4 + q4) ^ j) << 4), pkd[q4]);
        if (j == 7) {        // the tile's last column is column 0 of the next tile's buffer
          const uint32_t nrow = epi_base + (uint32_t)nbuf * Cfg::POOL_BUF_BYTES + (uint32_t)(g * 9 * 128);
          if (last_of_strip) {     // the next tile starts a strip: its column -1 is padding
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) pkd[q4] = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) sts128(nrow + ((uint32_t)((half * 4 + q4) ^ 7) << 4), pkd[q4]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(ld_bar0 + 8u * buf);
      }
      PROF_END(5 + ew)
    } else {
    epi_dispatch(a.fast ? epi_mode(epi) : EPI_HAS_ALL, epi_loop);
    if (a.fast) {
      if (elect_one()) tma_store_wait<0>();
      __syncwarp();
    }
    }
  }

teardown:
  tc_fence_before();
  if (PAIR) cluster_sync_all();      // neither CTA may exit while the other can still signal it
  else __syncthreads();
  if (warp == Cfg::MMA_WARP) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

static int make_tiled_map_4d(CUtensorMap* map, const void* base, int N, int H, int W, int C, int box_c,
                             int box_w, int box_h, int swizzle_bytes);

template <int CK, bool PAIR>
static int launch_halo(const vdqn_conv_desc* d, cudaStream_t stream) {
  using Cfg = HaloCfg<CK, PAIR>;
  DeviceInfo* dev = device_info();
  if (dev == nullptr) return VDQN_ERR_CUDA;
  static bool attr_set = false;
  auto kfn = halo_conv_kernel<CK, PAIR>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess)
      return set_error(VDQN_ERR_CUDA, "cudaFuncSetAttribute(halo_conv): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  CUtensorMap tmX, tmW;
  int rc = make_tiled_map_4d(&tmX, d->x, d->N, d->H, d->W, d->Cin, CK, Cfg::TW + d->S - 1,
                             Cfg::TH + d->R - 1, CK == 64 ? 128 : 32);
  if (rc != VDQN_OK) return rc;
  rc = make_tiled_map_2d(&tmW, d->w, (uint64_t)d->R * d->S * d->Cin, d->Cout, CK, Cfg::W_ROWS, CK == 64 ? 128 : 32);
  if (rc != VDQN_OK) return rc;
  CUtensorMap tmOut = tmX, tmRes = tmX, tmMask = tmX;
  const bool fast = fast_epilogue_ok(d) && d->ldc == d->Cout && (!d->residual || d->ldr == d->Cout) &&
                    (!d->mask_src || d->ldm == d->Cout);
  if (fast) {
    // per-warp tiles: 32 channels x 8 x 4 pixels, 64-byte rows
    rc = make_tiled_map_nhwc(&tmOut, d->out, d->N, d->H, d->W, d->Cout, 32, 8, 4, 64);
    if (rc == VDQN_OK && d->residual) rc = make_tiled_map_nhwc(&tmRes, d->residual, d->N, d->H, d->W, d->Cout, 32, 8, 4, 64);
    if (rc == VDQN_OK && d->mask_src) rc = make_tiled_map_nhwc(&tmMask, d->mask_src, d->N, d->H, d->W, d->Cout, 32, 8, 4, 64);
    if (rc != VDQN_OK) return rc;
  }
  HaloArgs a{};
  a.fast = fast ? 1 : 0;
  a.stages = Cfg::stages_for(fast ? (d->residual != nullptr ? 1 : 0) + (d->mask_src != nullptr ? 1 : 0) : 0);
  {
    static const int cap = [] { const char* e = getenv("VDQN_HALO_STAGES"); return e != nullptr ? atoi(e) : 0; }();
    if (CK == 64 && cap >= 2 && cap < a.stages) a.stages = cap;      // (experiments)
  }
  a.x = static_cast<const __nv_bfloat16*>(d->x);
  a.N = d->N; a.H = d->H; a.W = d->W; a.Cout = d->Cout;
  a.R = d->R; a.S = d->S; a.pad_lo = d->pad_lo;
  a.tiles_w = (d->W + Cfg::TW - 1) / Cfg::TW;
  a.tiles_h = (d->H + Cfg::TH - 1) / Cfg::TH;
  a.num_tiles = d->N * a.tiles_w * a.tiles_h;
  a.epi = make_epi_args(d);
  a.alias_from = d->x_alias_from; a.alias_shift = d->x_alias_shift;
  if (a.alias_from > 0 && (CK != 16 || a.alias_shift <= 0 || a.alias_shift > a.alias_from))
    return set_error(VDQN_ERR_ARG, "halo_conv: input aliasing is for the packed stem only (0 < shift <= first)");
  const int sms = d->max_ctas > 0 && d->max_ctas < dev->num_sms ? d->max_ctas : dev->num_sms;
  // schedule units: CTAs over tiles, or (PAIR) CTA pairs over pairs of tiles
  constexpr int MT = PAIR ? 2 : 1;
  const int units = a.num_tiles / MT, slots = sms / MT;
  const int grid = units < slots ? units : slots;
  CUtensorMap tmW2 = tmW;
  a.split_tile = 0; a.split_cta = 0; a.shift2 = d->shift2;
  if (d->split_n > 0) {
    if (d->w2 == nullptr || d->split_n >= d->N || grid < 2)
      return set_error(VDQN_ERR_SHAPE, "halo_conv: bad dual-network launch");
    rc = make_tiled_map_2d(&tmW2, d->w2, (uint64_t)d->R * d->S * d->Cin, d->Cout, CK, Cfg::W_ROWS, CK == 64 ? 128 : 32);
    if (rc != VDQN_OK) return rc;
    a.split_tile = d->split_n * a.tiles_w * a.tiles_h;
    int g0 = (int)((long)grid * a.split_tile / a.num_tiles);
    if (g0 < 1) g0 = 1;
    if (g0 > grid - 1) g0 = grid - 1;
    a.split_cta = g0 * MT;
  }
  if (PAIR)
    launch_kernel_cluster(kfn, grid * 2, Cfg::THREADS, Cfg::SMEM_BYTES, stream, 2, tmX, tmW, tmW2, tmOut, tmRes,
                          tmMask, a);
  else
    launch_kernel(kfn, grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream, tmX, tmW, tmW2, tmOut, tmRes, tmMask, a);
  VDQN_CHECK_LAUNCH("halo_conv launch");
  return VDQN_OK;
}

static int make_tiled_map_4d(CUtensorMap* map, const void* base, int N, int H, int W, int C, int box_c,
                             int box_w, int box_h, int swizzle_bytes) {
  return make_tiled_map_nhwc(map, base, N, H, W, C, box_c, box_w, box_h, swizzle_bytes);
}

// packed stem + max-pool in one kernel (HaloCfg POOL): d->pool_out receives max_pool2d(relu(conv + shift), 3, 2, 1)
static int launch_halo_pool(const vdqn_conv_desc* d, cudaStream_t stream) {
  using Cfg = HaloCfg<16, false, true>;
  DeviceInfo* dev = device_info();
  if (dev == nullptr) return VDQN_ERR_CUDA;
  if (d->H % Cfg::POOL_ROWS != 0 || d->W % Cfg::TW != 0 || (d->H & 1) || (d->W & 1) || !(d->flags & VDQN_EPI_RELU) ||
      d->residual != nullptr || d->mask_src != nullptr || d->colsum != nullptr)
    return set_error(VDQN_ERR_SHAPE, "halo_conv: fused pooling needs H %% 14 == 0, W %% 8 == 0 and a plain ReLU epilogue");
  static bool attr_set = false;
  auto kfn = halo_conv_kernel<16, false, true>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return set_error(VDQN_ERR_CUDA, "cudaFuncSetAttribute(halo_conv pool): %s", cudaGetErrorString(e));
    }
    attr_set = true;
  }
  CUtensorMap tmX, tmW;
  int rc = make_tiled_map_4d(&tmX, d->x, d->N, d->H, d->W, d->Cin, 16, Cfg::TW + d->S - 1, Cfg::TH + d->R - 1, 32);
  if (rc != VDQN_OK) return rc;
  rc = make_tiled_map_2d(&tmW, d->w, (uint64_t)d->R * d->S * d->Cin, d->Cout, 16, Cfg::W_ROWS, 32);
  if (rc != VDQN_OK) return rc;
  HaloArgs a{};
  a.fast = 0;
  a.x = static_cast<const __nv_bfloat16*>(d->x);
  a.N = d->N; a.H = d->H; a.W = d->W; a.Cout = d->Cout;
  a.R = d->R; a.S = d->S; a.pad_lo = d->pad_lo;
  a.tiles_w = d->W / Cfg::TW;
  a.tiles_h = d->H / Cfg::POOL_ROWS;                  // strips per image
  a.num_tiles = d->N * a.tiles_w * a.tiles_h;
  a.epi = make_epi_args(d);
  a.alias_from = d->x_alias_from; a.alias_shift = d->x_alias_shift;
  if (a.alias_from > 0 && (a.alias_shift <= 0 || a.alias_shift > a.alias_from))
    return set_error(VDQN_ERR_ARG, "halo_conv: input aliasing needs 0 < shift <= first");
  a.pool_out = static_cast<__nv_bfloat16*>(d->pool_out);
  a.pool_idx = d->pool_idx;
  a.idx_images = d->pool_idx_images;
  const int sms = d->max_ctas > 0 && d->max_ctas < dev->num_sms ? d->max_ctas : dev->num_sms;
  const int strips = d->N * a.tiles_h;
  const int grid = strips < sms ? strips : sms;
  CUtensorMap tmW2 = tmW;
  a.split_tile = 0; a.split_cta = 0; a.shift2 = d->shift2;
  if (d->split_n > 0) {
    if (d->w2 == nullptr || d->split_n >= d->N || grid < 2) return set_error(VDQN_ERR_SHAPE, "halo_conv: bad dual-network launch");
    rc = make_tiled_map_2d(&tmW2, d->w2, (uint64_t)d->R * d->S * d->Cin, d->Cout, 16, Cfg::W_ROWS, 32);
    if (rc != VDQN_OK) return rc;
    a.split_tile = d->split_n * a.tiles_w * a.tiles_h;
    int g0 = (int)((long)grid * d->split_n / d->N);
    if (g0 < 1) g0 = 1;
    if (g0 > grid - 1) g0 = grid - 1;
    a.split_cta = g0;
  }
  launch_kernel(kfn, grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream, tmX, tmW, tmW2, tmX, tmX, tmX, a);
  VDQN_CHECK_LAUNCH("halo_conv (stem + pool) launch");
  return VDQN_OK;
}

// Shapes this kernel takes: stride-1 "same" convolutions with Cout = 64 whose whole filter fits in
// shared memory: 3x3 over 64 channels (layer1 fwd/dgrad) and the packed 4x4 x 16-channel stem.
bool halo_conv_supported(const vdqn_conv_desc* d) {
  if (d->stride != 1 || d->dil != 1 || d->Cout != 64 || d->out_scatter >= 2 || d->out2 != nullptr)
    return false;
  if (d->pad_hi_w >= 0 && d->pad_hi_w != d->pad_hi) return false;
  if (d->Cin == 64 && d->R == 3 && d->S == 3 && d->pad_lo == 1 && d->pad_hi == 1) return true;
  if (d->Cin == 16 && d->R == 4 && d->S == 4 && d->pad_lo == 2 && d->pad_hi == 1) return true;
  return false;
}

int halo_conv_launch(const vdqn_conv_desc* d, cudaStream_t stream) {
  if (d->pool_out != nullptr) {
    if (d->Cin != 16) return set_error(VDQN_ERR_SHAPE, "halo_conv: fused pooling is for the packed stem");
    return launch_halo_pool(d, stream);
  }
  if (d->Cin != 64) return launch_halo<16, false>(d, stream);
  // CTA pairs need an even number of tiles (in both image ranges of a dual-network launch)
  const int per_img = ((d->W + 7) / 8) * ((d->H + 15) / 16);
  const bool even = ((long)d->N * per_img) % 2 == 0 && (d->split_n <= 0 || ((long)d->split_n * per_img) % 2 == 0);
  const bool pair = even && d->N * per_img >= 2 && d->algo != 4 && (d->algo == 3 || pair_default());
  if (d->algo == 3 && !pair)
    return set_error(VDQN_ERR_SHAPE, "halo_conv: CTA-pair kernel requested for an unsupported shape");
  return pair ? launch_halo<64, true>(d, stream) : launch_halo<64, false>(d, stream);
}

}  // namespace vdqn

VDQN_DEFINE_ROLE_PROFILE_READER(vdqn_debug_role_profile)
