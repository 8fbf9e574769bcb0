// Implicit-GEMM convolution on the 5th-gen tensor cores (sm_100a), used for the forward and
// data-gradient passes of every convolution of the Q-network trunk + head
// (reference call sites: archs/HabitatDQNMultiAction.py:49-51 -> torchvision BasicBlock,
//  SURVEY.md 2c rows K1-K5, K7).
//
//   D[m, co] = sum_{r,s,ci} X[n, p*stride + r*dil - pad, q*stride + s*dil - pad, ci] * W[co, r, s, ci]
//   m = (n, p, q) linearised; X is NHWC bf16; W is [Cout][R][S][Cin] bf16; fp32 accumulation.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0    : TMA producer.  A tile = 128 output pixels x CK channels of one filter tap, gathered
//               by the TMA unit in IM2COL mode (padding halo zero-filled, stride handled by the
//               traversal stride); B tile = BN x CK slab of the weight matrix (tiled TMA).
//   warp 1    : allocates TMEM, issues tcgen05.mma (M=128, N=BN, K=16) from one elected lane,
//               releases smem stages / publishes accumulators with tcgen05.commit.
//   warps 2-9 : epilogue, two warps per TMEM lane quadrant, each owning every other 32-column chunk
//               of the tile.  tcgen05.ld the fp32 accumulator (a ring in TMEM so the next
//               tile's MMAs overlap), then + per-channel shift (folded BN / bias), + residual,
//               ReLU or ReLU-mask (backward), optional per-channel column sums (d beta), bf16/fp32
//               store, optional stride-2 scatter (zero-dilated gradient for strided dgrad).
#include <stdlib.h>

#include <cstdlib>

#include "epilogue.cuh"
#include "ptx.cuh"
#include "vdqn_internal.h"
#include "role_profile.cuh"

namespace vdqn {

struct IgemmArgs {
  int M_total, Ho, Wo, Cout;
  int R, S, Cin, stride, dil, lower_h, lower_w;
  int num_m_tiles, num_n_tiles;
  EpiArgs epi;
  int out_scatter;    // 1: opix = m;  2: opix = (n*2Ho + 2p + off_h)*2Wo + 2q + off_w
  int off_h, off_w;
  int scatter_inputs; // residual / mask indexed by the scattered pixel
  int fast;           // staged TMA epilogue (BN <= 128, bf16 compact output)
  int alias_out;      // BN = 128 staged epilogue: the out tile is written in place of the mask (else residual) tile
  int direct_pre;     // BN = 256: direct epilogue with residual / mask rows loaded one chunk ahead (bf16 compact output)
  // dual-network launch: m-tiles [split_m_tile, num_m_tiles) use the second weight set; CTAs
  // [0, split_cta) work on the first range, the rest on the second (split_cta == 0: off)
  int split_m_tile, split_cta;
  const float* shift2;
  int stages;         // pipeline depth of this launch (IgemmCfg::stages_for)
  // Second K segment: after the R*S*Cin/CK k-blocks of the main operand, seg2_kb more whose A tiles come
  // from ANOTHER tensor (tmA2: a 1x1 window at stride2 over the same output grid) and whose B columns simply
  // continue in the weight matrix.  Forward: the 1x1/2 downsample of a residual block accumulated into its
  // conv2 (archs/HabitatDQNMultiAction.py -> torchvision BasicBlock: out = bn2(conv2(..)) + downsample(x));
  // backward: the downsample's data gradient accumulated into conv1's.
  int seg2_kb, stride2;
  // out_scatter == 3: the GEMM's Cout = 4 Cq columns are the 2x2 output-parity classes (a, b, c) of a stride-2
  // data gradient: row (n, p, q), column (a, b, c) is pixel (2p + a, 2q + b), channel c of dX [N][2Ho][2Wo][Cq]
  int Cq;
};

// PAIR: the CTA is one half of a cta_group::2 pair -- the pair computes a 256 x BN tile, this CTA
// stages its own 128 A rows and BN/2 of the B rows (see ptx.cuh)
template <int BN, int CK, bool PAIR = false>
struct IgemmCfg {
  static constexpr int BM = 128;
  static constexpr int KSUB = (CK == 64) ? 1 : 4;  // TMA sub-blocks per pipeline stage
  static constexpr int A_SUB_BYTES = BM * CK * 2;
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;
  static constexpr int B_SUB_BYTES = B_ROWS * CK * 2;
  static constexpr int STAGE_BYTES = KSUB * (A_SUB_BYTES + B_SUB_BYTES);
  // staged epilogue tiles (per epilogue warp: out / residual / mask, 4 KB per 64-column group);
  // BN = 256 keeps the direct epilogue (its long K loops hide it) and all its smem for the pipeline
  static constexpr bool FAST_EPI = BN <= 128;
  static constexpr int GROUPS = BN / 64;
  // BN = 64: the out tile is double-buffered so a tile never waits for the previous tile's TMA store
  static constexpr int OUT_BUFS = (GROUPS == 1) ? 2 : 1;
  // eight epilogue warps; a warp's staging tiles are [32 pixels][32 channels] (64-byte rows, 64B
  // swizzle), one set (residual, mask, OUT_BUFS x out) per 32-column chunk it owns (GROUPS of them)
  static constexpr int EPI_WARPS = 8;
  static constexpr int THREADS = (2 + EPI_WARPS) * 32;
  // Shared memory is split AT LAUNCH between the pipeline and the epilogue staging: a launch only
  // reserves the staging tiles it uses (out, + residual, + mask), the rest becomes pipeline stages.
  // The BN = 128 layers are bound by bytes in flight (5 -> 3 stages costs 27 %): a plain forward conv
  // gets 8 stages instead of the 5 a worst-case static split would leave.
  static constexpr int SMEM_LAYOUT = 224 * 1024;
#ifndef VDQN_MAX_STAGES
#define VDQN_MAX_STAGES 8          // (experiments: -DVDQN_MAX_STAGES=n caps the pipeline depth)
#endif
  static constexpr int MAX_STAGES = VDQN_MAX_STAGES;
  __host__ __device__ static constexpr int epi_warp_bytes(int n_in) { return FAST_EPI ? GROUPS * (n_in + OUT_BUFS) * 2048 : 0; }
  __host__ __device__ static constexpr int stages_for(int n_in) {
    return (SMEM_LAYOUT - EPI_WARPS * epi_warp_bytes(n_in)) / STAGE_BYTES > MAX_STAGES
               ? MAX_STAGES
               : (SMEM_LAYOUT - EPI_WARPS * epi_warp_bytes(n_in)) / STAGE_BYTES;
  }
  // accumulator ring over all 512 TMEM columns (8 / 4 / 2 accumulators for BN = 64 / 128 / 256): the
  // MMA warp may run that many tiles ahead of the epilogue, which hides the mbarrier hand-off
  // latencies that otherwise bound launches with few k-blocks per tile (1x1, parity-class convs)
  static constexpr int NACC = 512 / BN;
  static constexpr int TMEM_COLS = 512;
  static constexpr int SMEM_BYTES = SMEM_LAYOUT + 1024 /*align*/ + 512 /*barriers*/;
  static constexpr uint64_t SWZ = (CK == 64) ? kSwz128 : kSwz32;
  static constexpr int ROW_BYTES = CK * 2;          // bytes per smem row (= swizzle span)
  static constexpr int SBO = 8 * ROW_BYTES;         // 8-row group pitch
};

template <int BN, int CK, bool PAIR>
__global__ void __launch_bounds__(320, 1)
igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
             const __grid_constant__ CUtensorMap tmMask, const __grid_constant__ CUtensorMap tmA2, const IgemmArgs a) {
  using Cfg = IgemmCfg<BN, CK, PAIR>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int nstages = a.stages;                      // pipeline depth of this launch (host: stages_for)
  const uint32_t epi_base = smem_base + nstages * Cfg::STAGE_BYTES;
  const uint32_t bar_base = smem_base + Cfg::SMEM_LAYOUT;
  // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], then tmem ptr slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::MAX_STAGES + s); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * Cfg::MAX_STAGES + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * Cfg::MAX_STAGES + Cfg::NACC + i); };
  const uint32_t ld_bar0 = bar_base + 8u * (2 * Cfg::MAX_STAGES + 2 * Cfg::NACC);     // one per epilogue warp
  const uint32_t tmem_slot = ld_bar0 + 8u * Cfg::EPI_WARPS;
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  // warp index via a shuffle broadcast so the compiler keeps role dispatch (and the TMA / MMA
  // operands computed under it) on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::MAX_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int i = 0; i < Cfg::NACC; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), (PAIR ? 2 : 1) * Cfg::EPI_WARPS);   // one arrive per epilogue warp (of both CTAs of a pair)
    }
    for (int i = 0; i < Cfg::EPI_WARPS; ++i) mbar_init(ld_bar0 + 8u * i, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
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
  pdl_wait();      // the previous kernel's global writes are visible from here on

  // this CTA's slice of the tile space (whole launch, or one of the two networks' image ranges).
  // PAIR: the schedule runs over CTA pairs and pairs of m-tiles; rank r takes m-tile 2*m_pair + r.
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  constexpr int MT = PAIR ? 2 : 1;
  const int cta = (int)blockIdx.x / MT, ncta = (int)gridDim.x / MT;
  const int split_cta = a.split_cta / MT, split_mt = a.split_m_tile / MT, n_mt = a.num_m_tiles / MT;
  const bool second = split_cta > 0 && cta >= split_cta;
  const int tile0 = (second ? split_mt * a.num_n_tiles : 0) + cta - (second ? split_cta : 0);
  const int tstep = split_cta > 0 ? (second ? ncta - split_cta : split_cta) : ncta;
  const int num_tiles = (split_cta > 0 && !second ? split_mt : n_mt) * a.num_n_tiles;
  const CUtensorMap* tmBp = second ? &tmB2 : &tmB;
  const int cblks = a.Cin / CK;
  const int num_sub = a.R * a.S * cblks;
  const int num_kb = num_sub / Cfg::KSUB;
  const int total_kb = num_kb + a.seg2_kb;          // seg2_kb > 0 only on the 64-channel-block path (KSUB == 1)
  const int HoWo = a.Ho * a.Wo;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    PROF_BEGIN
    for (int t = tile0; t < num_tiles; t += tstep) {
      PROF_TILE
      const int n_t = t % a.num_n_tiles, m_t = (t / a.num_n_tiles) * MT + rank;
      const int m0 = m_t * Cfg::BM;
      const int img = m0 / HoWo;
      const int rem = m0 - img * HoWo;
      const int p0 = rem / a.Wo, q0 = rem - p0 * a.Wo;
      const int cw = q0 * a.stride + a.lower_w, ch = p0 * a.stride + a.lower_h;
      // (tap row, tap column, channel block) advance incrementally: the producer is one thread and
      // an integer division per k-block would sit on its critical path
      int c0 = 0, off_w = 0, off_h = 0, s_i = 0, jk = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        PROF_WAIT_A(mbar_wait(empty_bar(stage), phase ^ 1))
        const bool leader = elect_one();
        // PAIR: the leader CTA's barrier collects the bytes of both CTAs' loads
        if (leader && rank == 0) mbar_expect_tx(full_bar(stage), MT * Cfg::STAGE_BYTES);
        const uint32_t fbar = PAIR ? mapa_u32(full_bar(stage), 0) : full_bar(stage);
        const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
        const uint32_t sB = sA + Cfg::KSUB * Cfg::A_SUB_BYTES;
#pragma unroll
        for (int sub = 0; sub < Cfg::KSUB; ++sub) {
          if (leader) {
            if (PAIR) {
              tma_load_im2col_4d_pair(sA + sub * Cfg::A_SUB_BYTES, &tmA, fbar, c0, cw, ch, img,
                                      (uint16_t)off_w, (uint16_t)off_h);
              tma_load_2d_pair(sB + sub * Cfg::B_SUB_BYTES, tmBp, fbar, jk, n_t * BN + rank * Cfg::B_ROWS);
            } else {
              tma_load_im2col_4d(sA + sub * Cfg::A_SUB_BYTES, &tmA, fbar, c0, cw, ch, img,
                                 (uint16_t)off_w, (uint16_t)off_h);
              tma_load_2d(sB + sub * Cfg::B_SUB_BYTES, tmBp, fbar, jk, n_t * BN);
            }
          }
          jk += CK;
          c0 += CK;
          if (c0 == a.Cin) {
            c0 = 0;
            off_w += a.dil;
            if (++s_i == a.S) { s_i = 0; off_w = 0; off_h += a.dil; }
          }
        }
        __syncwarp();
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
      if constexpr (CK == 64) {
        // second segment: 1x1 window over the other tensor, B columns continue at jk
        const int cw2 = q0 * a.stride2, ch2 = p0 * a.stride2;
        for (int kb = 0; kb < a.seg2_kb; ++kb) {
          PROF_WAIT_A(mbar_wait(empty_bar(stage), phase ^ 1))
          const bool leader = elect_one();
          if (leader && rank == 0) mbar_expect_tx(full_bar(stage), MT * Cfg::STAGE_BYTES);
          const uint32_t fbar = PAIR ? mapa_u32(full_bar(stage), 0) : full_bar(stage);
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::KSUB * Cfg::A_SUB_BYTES;
          if (leader) {
            if (PAIR) {
              tma_load_im2col_4d_pair(sA, &tmA2, fbar, kb * CK, cw2, ch2, img, 0, 0);
              tma_load_2d_pair(sB, tmBp, fbar, jk, n_t * BN + rank * Cfg::B_ROWS);
            } else {
              tma_load_im2col_4d(sA, &tmA2, fbar, kb * CK, cw2, ch2, img, 0, 0);
              tma_load_2d(sB, tmBp, fbar, jk, n_t * BN);
            }
          }
          jk += CK;
          __syncwarp();
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
    PROF_END(0)
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (PAIR: leader CTA only)
    constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, BN, 0, 0);
    if (PAIR && rank != 0) goto teardown;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    PROF_BEGIN
    for (int t = tile0; t < num_tiles; t += tstep, ++it) {
      PROF_TILE
      const int acc = it % Cfg::NACC;
      const uint32_t acc_phase = (it / Cfg::NACC) & 1;
      PROF_WAIT_A(mbar_wait(tempty_bar(acc), acc_phase ^ 1))
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < total_kb; ++kb) {
        PROF_WAIT_B(mbar_wait(full_bar(stage), phase))
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sA = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sB = sA + Cfg::KSUB * Cfg::A_SUB_BYTES;
#pragma unroll
          for (int sub = 0; sub < Cfg::KSUB; ++sub) {
#pragma unroll
            for (int k = 0; k < CK / 16; ++k) {
              const uint64_t ad =
                  make_smem_desc(sA + sub * Cfg::A_SUB_BYTES + k * 32, 16, Cfg::SBO, Cfg::SWZ);
              const uint64_t bd =
                  make_smem_desc(sB + sub * Cfg::B_SUB_BYTES + k * 32, 16, Cfg::SBO, Cfg::SWZ);
              if (PAIR) umma_f16_pair(d_tmem, ad, bd, idesc, (kb | sub | k) != 0);
              else umma_f16(d_tmem, ad, bd, idesc, (kb | sub | k) != 0);
            }
          }
          if (PAIR) {
            umma_commit_pair(empty_bar(stage));
            if (kb == total_kb - 1) umma_commit_pair(tfull_bar(acc));
          } else {
            umma_commit(empty_bar(stage));
            if (kb == total_kb - 1) umma_commit(tfull_bar(acc));
          }
        }
        __syncwarp();
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
    }
    PROF_END(4)
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    const int ew = warp - 2;                   // 0..7
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may touch
    const int half = ew >> 2;                  // it owns the 32-column chunks half, half + 2, ...
    const int row = quad * 32 + lane;
    constexpr int NCH = BN / 64;               // chunks per warp
    // per-warp staging: res[NCH] | mask[NCH] | out[OUT_BUFS][NCH], 2 KB tiles
    // per-warp staging of this launch: [residual x NCH] [mask x NCH] [out x OUT_BUFS x NCH], absent
    // inputs take no room
    const int n_in = (a.epi.residual != nullptr ? 1 : 0) + (a.epi.mask_src != nullptr ? 1 : 0);
    // alias_out: no out tiles of their own -- a lane overwrites the mask (else residual) row it has just consumed
    // with its output row, and the next tile's inputs are only requested once the TMA store has read the tile.
    // The shared memory saved becomes pipeline stages (mask only: 6 -> 8, residual + mask: 5 -> 6), which is
    // what the 128-wide launches are short of; their long k-loops hide the later prefetch.
    const bool alias = Cfg::FAST_EPI && a.alias_out != 0;
    const uint32_t stg_in = epi_base + ew * (uint32_t)(alias ? n_in * NCH * 2048 : Cfg::epi_warp_bytes(n_in));
    const uint32_t stg_mask0 = stg_in + (a.epi.residual != nullptr ? NCH * 2048 : 0);
    const uint32_t stg_out_base = alias ? (a.epi.mask_src != nullptr ? stg_mask0 : stg_in) : stg_in + n_in * NCH * 2048;
    const uint32_t ld_bar = ld_bar0 + 8u * ew;
    EpiArgs epi = a.epi;
    if (second) epi.shift = a.shift2;
    const bool fast = Cfg::FAST_EPI && a.fast;
    // the tile loop is instantiated once per combination of optional epilogue steps; the launch picks
    // its specialisation (epilogue.cuh, epi_dispatch)
    auto epi_loop = [&](auto mode_tag) {
    constexpr int EPI = decltype(mode_tag)::value;
    const bool has_res = (EPI & EPI_HAS_RES) && epi.residual != nullptr;
    const bool has_mask = (EPI & EPI_HAS_MASK) && epi.mask_src != nullptr;
    const bool has_in = has_res || has_mask;
    uint32_t ld_parity = 0;
    float csum[NCH];                     // per-lane column sums of this warp's chunks of the current column tile
#pragma unroll
    for (int i = 0; i < NCH; ++i) csum[i] = 0.f;
    // staged path: each thread keeps running sums of its own row (32 per chunk) and the warp
    // transpose-reduce runs once per column tile instead of once per chunk per tile
    constexpr bool ROWACC = Cfg::FAST_EPI && (EPI & EPI_HAS_COLSUM) != 0;
    float row_acc[ROWACC ? NCH * 32 : 1];
#pragma unroll
    for (int i = 0; i < (ROWACC ? NCH * 32 : 1); ++i) row_acc[i] = 0.f;
    int cs_nt = -1;
    auto flush_colsum = [&]() {
      if (a.epi.colsum != nullptr && cs_nt >= 0) {
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
          if constexpr (ROWACC) {
            if (fast) {
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) { v[j] = row_acc[i * 32 + j]; row_acc[i * 32 + j] = 0.f; }
              csum[i] += warp_transpose_reduce(v, lane);
            }
          }
          const int cidx = cs_nt * BN + (half + 2 * i) * 32 + lane;
          atomicAdd(a.epi.colsum + (a.out_scatter == 3 ? cidx % a.Cq : cidx), csum[i]);
          csum[i] = 0.f;
        }
      }
    };
    // accumulator drained: tell the MMA issuer (PAIR: the leader CTA's barrier, counted over both CTAs)
    auto release_acc = [&](int acc_i) {
      if (PAIR && rank != 0) mbar_arrive_cluster(mapa_u32(tempty_bar(acc_i), 0));
      else mbar_arrive(tempty_bar(acc_i));
    };
    // scatter launches (zero-dilated destination: parity-class data gradients): the same staged
    // tiles, but rows live at per-pixel addresses a tensor map cannot describe, so the LSU moves
    // them -- four lanes per 64-byte row (whole sectors), cp.async for the inputs
    const bool gather = a.out_scatter >= 2;
    const bool blocks2 = a.out_scatter == 3;
    // out_scatter 2: pixel (2p + off_h, 2q + off_w) of the zero-dilated destination, row pitch ld.
    // out_scatter 3: the column tile starting at col0 lies in image row 2p + col0 / (2 Cq), starting at pixel 2q:
    //   element offset = pixel * Cq + col % (2 Cq)  (the two pixels 2q, 2q+1 are adjacent in memory)
    auto scatter_pix = [&](int mm, int nn_t) -> long {
      if (mm >= a.M_total) return -1;
      const int img = mm / HoWo;
      const int rem = mm - img * HoWo;
      const int p = rem / a.Wo, q = rem - p * a.Wo;
      if (blocks2) return ((long)img * (2 * a.Ho) + 2 * p + (nn_t * BN) / (2 * a.Cq)) * (2 * a.Wo) + 2 * q;
      return ((long)img * (2 * a.Ho) + 2 * p + a.off_h) * (2 * a.Wo) + 2 * q + a.off_w;
    };
    // element offset of column `col` of a row whose scatter_pix is `pix` in a tensor of row pitch `ld`
    auto scatter_off = [&](long pix, int col, int ld) -> long {
      return blocks2 ? pix * a.Cq + (col % (2 * a.Cq)) : pix * ld + col;
    };
    auto tile_mn = [&](int tt, int& nn_t, int& mm_t) {
      nn_t = tt % a.num_n_tiles;
      mm_t = (tt / a.num_n_tiles) * MT + rank;
    };
    // residual / mask tiles are prefetched one tile ahead (see halo_conv.cu)
    auto issue_inputs = [&](int tt) {
      int nn_t, mm_t;
      tile_mn(tt, nn_t, mm_t);
      if (gather) {
        const long pix = scatter_pix(mm_t * Cfg::BM + row, nn_t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = 8 * i + (lane >> 2);
          const long pr = __shfl_sync(0xffffffffu, pix, r);
          const uint32_t dst = (uint32_t)r * 64u + ((uint32_t)((lane & 3) ^ ((r >> 1) & 3)) << 4);
          const uint32_t nb = pr >= 0 ? 16u : 0u;
          const long prc = pr >= 0 ? pr : 0;
#pragma unroll
          for (int ci = 0; ci < NCH; ++ci) {
            const int col = nn_t * BN + (half + 2 * ci) * 32 + (lane & 3) * 8;
            if (has_res) cp_async_16(stg_in + ci * 2048 + dst, epi.residual + scatter_off(prc, col, epi.ldr), nb);
            if (has_mask) cp_async_16(stg_mask0 + ci * 2048 + dst, epi.mask_src + scatter_off(prc, col, epi.ldm), nb);
          }
        }
        cp_async_commit();
        return;
      }
      const int r0 = mm_t * Cfg::BM + quad * 32;
      if (elect_one()) {
        mbar_expect_tx(ld_bar, NCH * ((has_res ? 2048u : 0u) + (has_mask ? 2048u : 0u)));
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int col = nn_t * BN + (half + 2 * ci) * 32;
          if (has_res) tma_load_2d(stg_in + ci * 2048, &tmRes, ld_bar, col, r0);
          if (has_mask) tma_load_2d(stg_mask0 + ci * 2048, &tmMask, ld_bar, col, r0);
        }
      }
      __syncwarp();
    };
    if (fast && has_in && tile0 < num_tiles) issue_inputs(tile0);
    EpiRegs pre;
    if constexpr (!Cfg::FAST_EPI) {
      if (a.direct_pre && tile0 < num_tiles) {
        int n0, m0;
        tile_mn(tile0, n0, m0);
        const long mrow = (long)m0 * Cfg::BM + row;
        epilogue_load_regs<EPI>(epi, mrow < a.M_total, mrow, n0 * BN + half * 32, pre);
      }
    }
    int it = 0;
    PROF_BEGIN
    for (int t = tile0; t < num_tiles; t += tstep, ++it) {
      PROF_TILE
      int n_t, m_t;
      tile_mn(t, n_t, m_t);
      if (n_t != cs_nt) { flush_colsum(); cs_nt = n_t; }
      const int acc = it % Cfg::NACC;
      const uint32_t acc_phase = (it / Cfg::NACC) & 1;
      const int m = m_t * Cfg::BM + row;
      const bool valid = m < a.M_total;
      if (fast) {
        const int row0 = m_t * Cfg::BM + quad * 32;            // this warp's 32 output rows
        PROF_WAIT_A(mbar_wait(tfull_bar(acc), acc_phase))
        tc_fence_after();
        const uint32_t stg = alias ? stg_out_base : stg_out_base + (uint32_t)((it % Cfg::OUT_BUFS) * NCH) * 2048u;
        if (!alias) {
          PROF_WAIT_B(if (elect_one()) tma_store_wait_read<Cfg::OUT_BUFS - 1>(); __syncwarp())   // this out buffer's last store has been read
        }
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int chunk = half + 2 * ci;
          uint32_t raw[32];
          tmem_ld_32x32(tmem_base + acc * BN + chunk * 32 + ((uint32_t)(quad * 32) << 16), raw);
          tmem_ld_wait();
          if (ci == 0 && has_in) {
            if (gather) { PROF_WAIT_B(cp_async_wait_all(); __syncwarp()) }
            else { PROF_WAIT_B(mbar_wait(ld_bar, ld_parity)) }
          }
          float* racc = ROWACC ? row_acc + ci * 32 : nullptr;      // ci is a compile-time constant here
          const float cs = epilogue_half_staged<64, EPI>(epi, raw, valid, n_t * BN + chunk * 32, 0, lane, stg + ci * 2048,
                                                    stg_in + ci * 2048, stg_mask0 + ci * 2048, racc);
#pragma unroll
          for (int i = 0; i < NCH; ++i)
            if (i == ci) csum[i] += cs;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc(acc);
        if (has_in) {
          ld_parity ^= 1;
          if (!alias && t + tstep < num_tiles) issue_inputs(t + tstep);
        }
        if (gather) {
          __syncwarp();                       // the staged tiles are complete
          const long pix = scatter_pix(m, n_t);
          __nv_bfloat16* obase = static_cast<__nv_bfloat16*>(epi.out);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = 8 * i + (lane >> 2);
            const long pr = __shfl_sync(0xffffffffu, pix, r);
            const uint32_t src = (uint32_t)r * 64u + ((uint32_t)((lane & 3) ^ ((r >> 1) & 3)) << 4);
#pragma unroll
            for (int ci = 0; ci < NCH; ++ci) {
              const uint4 v4 = lds128(stg + ci * 2048 + src);
              if (pr >= 0)
                *reinterpret_cast<uint4*>(obase + scatter_off(pr, n_t * BN + (half + 2 * ci) * 32 + (lane & 3) * 8, epi.ldc)) = v4;
            }
          }
          __syncwarp();
          continue;
        }
        fence_proxy_async();
        __syncwarp();
        if (elect_one()) {
#pragma unroll
          for (int ci = 0; ci < NCH; ++ci)
            tma_store_2d(&tmOut, stg + ci * 2048, n_t * BN + (half + 2 * ci) * 32, row0);
          tma_store_commit();
        }
        __syncwarp();
        if (alias) {
          // the tile the store reads is the next tile's input buffer
          PROF_WAIT_B(if (elect_one()) tma_store_wait_read<0>(); __syncwarp())
          if (t + tstep < num_tiles) issue_inputs(t + tstep);
        }
        continue;
      }
      if constexpr (!Cfg::FAST_EPI) {
        if (a.direct_pre) {
          // inputs of chunk ci were loaded while chunk ci - 1 (or the previous tile's last chunk) was processed
          mbar_wait(tfull_bar(acc), acc_phase);
          tc_fence_after();
#pragma unroll
          for (int ci = 0; ci < NCH; ++ci) {
            const int chunk = half + 2 * ci;
            uint32_t raw[32];
            tmem_ld_32x32(tmem_base + acc * BN + chunk * 32 + ((uint32_t)(quad * 32) << 16), raw);
            EpiRegs nxt;
            if (ci + 1 < NCH) {
              epilogue_load_regs<EPI>(epi, valid, (long)m, n_t * BN + (chunk + 2) * 32, nxt);
            } else if (t + tstep < num_tiles) {
              int n2, m2;
              tile_mn(t + tstep, n2, m2);
              const long mrow = (long)m2 * Cfg::BM + row;
              epilogue_load_regs<EPI>(epi, mrow < a.M_total, mrow, n2 * BN + half * 32, nxt);
            }
            tmem_ld_wait();
            const float cs = epilogue_chunk_regs<EPI>(epi, raw, valid, (long)m, n_t * BN + chunk * 32, lane, pre);
            pre = nxt;
#pragma unroll
            for (int i = 0; i < NCH; ++i)
              if (i == ci) csum[i] += cs;
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc(acc);
          continue;
        }
      }
      long opix = m;
      long opix2 = 0;
      if (a.out_scatter == 2 || a.epi.out2 != nullptr) {
        const int img = m / HoWo;
        const int rem = m - img * HoWo;
        const int p = rem / a.Wo, q = rem - p * a.Wo;
        const long dil_pix = ((long)img * (2 * a.Ho) + 2 * p) * (2 * a.Wo) + 2 * q;
        if (a.out_scatter == 2) opix = ((long)img * (2 * a.Ho) + 2 * p + a.off_h) * (2 * a.Wo) + 2 * q + a.off_w;
        opix2 = dil_pix;
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int ci = 0; ci < NCH; ++ci) {
        const int chunk = half + 2 * ci;
        uint32_t raw[32];
        tmem_ld_32x32(tmem_base + acc * BN + chunk * 32 + ((uint32_t)(quad * 32) << 16), raw);
        tmem_ld_wait();
        const float cs = epilogue_chunk(epi, raw, valid, a.scatter_inputs ? opix : (long)m, opix, opix2,
                                        n_t * BN + chunk * 32, lane);
#pragma unroll
        for (int i = 0; i < NCH; ++i)
          if (i == ci) csum[i] += cs;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) release_acc(acc);
    }
    PROF_END(5 + ew)
    flush_colsum();
    };
    if constexpr (Cfg::FAST_EPI) epi_dispatch(fast ? epi_mode(epi) : EPI_HAS_ALL, epi_loop);
    else epi_dispatch(a.direct_pre ? epi_mode(epi) : EPI_HAS_ALL, epi_loop);
    if (fast) {
      if (elect_one()) tma_store_wait<0>();
      __syncwarp();
    }
  }

teardown:
  tc_fence_before();
  if (PAIR) cluster_sync_all();      // neither CTA may exit while the other can still signal it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------- host side

template <int BN, int CK, bool PAIR = false>
static int launch_igemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmB2,
                        const CUtensorMap* epi_maps, IgemmArgs& a, int num_sms, cudaStream_t stream) {
  const CUtensorMap& tmA2 = epi_maps[3];
  using Cfg = IgemmCfg<BN, CK, PAIR>;
  static bool attr_set = false;
  auto kfn = igemm_kernel<BN, CK, PAIR>;
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(VDQN_ERR_CUDA, "cudaFuncSetAttribute(igemm): %s",
                                           cudaGetErrorString(e));
    attr_set = true;
  }
  a.stages = Cfg::stages_for(a.fast ? (a.epi.residual != nullptr ? 1 : 0) + (a.epi.mask_src != nullptr ? 1 : 0) : 0);
  if (a.alias_out) {
    const int n_in = (a.epi.residual != nullptr ? 1 : 0) + (a.epi.mask_src != nullptr ? 1 : 0);
    const int st = (Cfg::SMEM_LAYOUT - Cfg::EPI_WARPS * n_in * Cfg::GROUPS * 2048) / Cfg::STAGE_BYTES;
    a.stages = st > Cfg::MAX_STAGES ? Cfg::MAX_STAGES : st;
  }
  if (!a.fast) a.stages = Cfg::SMEM_LAYOUT / Cfg::STAGE_BYTES > Cfg::MAX_STAGES ? Cfg::MAX_STAGES
                                                                                : Cfg::SMEM_LAYOUT / Cfg::STAGE_BYTES;
  // schedule units: CTAs over tiles, or (PAIR) CTA pairs over pairs of m-tiles
  constexpr int MT = PAIR ? 2 : 1;
  const int tiles = (a.num_m_tiles / MT) * a.num_n_tiles;
  const int slots = num_sms / MT;
  int grid = tiles < slots ? tiles : slots;
  if (a.split_m_tile > 0) {
    // two image ranges (two weight sets): give each its share of the SMs
    const int t0 = (a.split_m_tile / MT) * a.num_n_tiles, t1 = tiles - t0;
    if (t1 <= 0 || grid < 2) return set_error(VDQN_ERR_SHAPE, "conv_gemm: empty second image range");
    int g0 = (int)((long)grid * t0 / tiles);
    if (g0 < 1) g0 = 1;
    if (g0 > grid - 1) g0 = grid - 1;
    a.split_cta = g0 * MT;
  }
  if (PAIR)
    launch_kernel_cluster(kfn, grid * 2, Cfg::THREADS, Cfg::SMEM_BYTES, stream, 2, tmA, tmB, tmB2, epi_maps[0], epi_maps[1],
                          epi_maps[2], tmA2, a);
  else
    launch_kernel(kfn, grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream, tmA, tmB, tmB2, epi_maps[0], epi_maps[1], epi_maps[2], tmA2, a);
  VDQN_CHECK_LAUNCH("igemm launch");
  return VDQN_OK;
}

}  // namespace vdqn

using namespace vdqn;

extern "C" int vdqn_conv_gemm(const vdqn_conv_desc* d, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (d == nullptr) return set_error(VDQN_ERR_ARG, "conv_gemm: null descriptor");
  const int CK = (d->Cin % 64 == 0) ? 64 : 16;
  if (d->Cin % CK != 0) return set_error(VDQN_ERR_SHAPE, "conv_gemm: Cin=%d not a multiple of 16", d->Cin);
  if (CK == 16 && (d->R * d->S * (d->Cin / 16)) % 4 != 0)
    return set_error(VDQN_ERR_SHAPE, "conv_gemm: 16-channel path needs R*S*Cin/16 %% 4 == 0");
  const int Ho = (d->H + d->pad_lo + d->pad_hi - (d->R - 1) * d->dil - 1) / d->stride + 1;
  const int pad_hi_w = d->pad_hi_w >= 0 ? d->pad_hi_w : d->pad_hi;
  const int Wo = (d->W + d->pad_lo + pad_hi_w - (d->S - 1) * d->dil - 1) / d->stride + 1;
  if (Ho <= 0 || Wo <= 0) return set_error(VDQN_ERR_SHAPE, "conv_gemm: empty output");
  int BN = d->Cout >= 256 ? 256 : d->Cout;
  if (d->tile_n > 0) BN = d->tile_n;
  if (!(BN == 64 || BN == 128 || BN == 256) || d->Cout % BN != 0)
    return set_error(VDQN_ERR_SHAPE, "conv_gemm: Cout=%d unsupported (tile %d)", d->Cout, BN);
  if (CK == 16 && BN != 64) return set_error(VDQN_ERR_SHAPE, "conv_gemm: 16-channel path is Cout=64 only");
  if (d->ldc % 8 != 0 || (d->residual && d->ldr % 8 != 0) || (d->mask_src && d->ldm % 8 != 0))
    return set_error(VDQN_ERR_SHAPE, "conv_gemm: leading dimensions must be multiples of 8");
  if ((reinterpret_cast<uintptr_t>(d->x) | reinterpret_cast<uintptr_t>(d->w) |
       reinterpret_cast<uintptr_t>(d->out)) & 15)
    return set_error(VDQN_ERR_ARG, "conv_gemm: pointers must be 16-byte aligned");

  DeviceInfo* dev = device_info();
  if (dev == nullptr) return VDQN_ERR_CUDA;
  if (d->algo == 2 && !halo_conv_supported(d))
    return set_error(VDQN_ERR_SHAPE, "conv_gemm: halo algorithm requested for an unsupported shape");
  if (d->algo != 1 && d->tile_n == 0 && d->x2 == nullptr && halo_conv_supported(d)) return halo_conv_launch(d, stream);
  if (d->x_alias_from > 0)
    return set_error(VDQN_ERR_ARG, "conv_gemm: input aliasing needs the packed-stem halo kernel");

  CUtensorMap tmA, tmB;
  int rc = make_im2col_map(&tmA, d->x, d->N, d->H, d->W, d->Cin, CK, 128, d->stride,
                           -d->pad_lo, -d->pad_lo,
                           d->pad_hi - (d->R - 1) * d->dil, pad_hi_w - (d->S - 1) * d->dil,
                           CK == 64 ? 128 : 32);
  if (rc != VDQN_OK) return rc;
  // CTA pairs (cta_group::2, 256 x BN tiles): wide tiles of the 64-channel-block path, an even
  // number of m-tiles (in both image ranges of a dual-network launch).  algo 3 demands it, 4 forbids.
  const long M_total_l = (long)d->N * Ho * Wo;
  const int m_tiles = (int)((M_total_l + 127) / 128);
  bool pair = CK == 64 && BN >= 128 && m_tiles % 2 == 0 && d->algo != 4 && (d->algo == 3 || pair_default());
  if (pair && d->split_n > 0 && (((long)d->split_n * Ho * Wo) / 128) % 2 != 0) pair = false;
  if (d->algo == 3 && !pair)
    return set_error(VDQN_ERR_SHAPE, "conv_gemm: CTA-pair kernel requested for an unsupported shape");
  const int b_rows = pair ? BN / 2 : BN;
  // second K segment (x2: a 1x1 / stride2 window over another tensor, same output grid): the weight rows are
  // [R*S*Cin | Cin2] long
  const bool seg2 = d->x2 != nullptr;
  if (seg2) {
    if (CK != 64 || d->Cin2 % 64 != 0 || d->stride2 < 1 || (reinterpret_cast<uintptr_t>(d->x2) & 15))
      return set_error(VDQN_ERR_SHAPE, "conv_gemm: the second operand needs 64-channel blocks");
    if ((d->H2 - 1) / d->stride2 + 1 != Ho || (d->W2 - 1) / d->stride2 + 1 != Wo)
      return set_error(VDQN_ERR_SHAPE, "conv_gemm: the second operand does not cover the same output grid");
  }
  const uint64_t k_total = (uint64_t)d->R * d->S * d->Cin + (seg2 ? d->Cin2 : 0);
  rc = make_tiled_map_2d(&tmB, d->w, k_total, d->Cout, CK, b_rows, CK == 64 ? 128 : 32);
  if (rc != VDQN_OK) return rc;

  IgemmArgs a{};
  a.M_total = d->N * Ho * Wo;
  a.Ho = Ho; a.Wo = Wo; a.Cout = d->Cout;
  a.R = d->R; a.S = d->S; a.Cin = d->Cin; a.stride = d->stride; a.dil = d->dil;
  a.lower_h = -d->pad_lo; a.lower_w = -d->pad_lo;
  a.num_m_tiles = (a.M_total + 127) / 128;
  a.num_n_tiles = d->Cout / BN;
  a.epi = make_epi_args(d);
  a.out_scatter = (d->out_scatter == 2 || d->out_scatter == 3) ? d->out_scatter : 1;
  a.off_h = d->scatter_off_h; a.off_w = d->scatter_off_w;
  a.scatter_inputs = (d->out_scatter >= 2 && (d->flags & VDQN_EPI_SCATTER_INPUTS)) ? 1 : 0;
  a.seg2_kb = seg2 ? d->Cin2 / 64 : 0;
  a.stride2 = seg2 ? d->stride2 : 1;
  a.Cq = d->Cout / 4;
  if (d->out_scatter == 3 && (BN > 2 * a.Cq || (2 * a.Cq) % BN != 0 || d->Cout % 4 != 0))
    return set_error(VDQN_ERR_SHAPE, "conv_gemm: 2x2-block scatter needs column tiles inside one image row (tile %d, Cout %d)", BN, d->Cout);

  // staged epilogue: 2-D maps over the [M][ld] output / residual / mask matrices, boxes of
  // 32 channels x 32 pixels (64-byte rows)
  CUtensorMap epi_maps[4] = {tmB, tmB, tmB, tmA};
  a.fast = (BN <= 128 && fast_epilogue_ok(d)) ? 1 : 0;
  {
    static const bool pre_on = [] { const char* e = getenv("VDQN_DIRECT_PRE"); return e == nullptr || atoi(e) != 0; }();
    a.direct_pre = (pre_on && BN == 256 && d->out_scatter <= 1 && !(d->flags & VDQN_EPI_SCATTER_INPUTS) && fast_epilogue_ok(d)) ? 1 : 0;
  }
  {
    // 128-wide tiles with inputs and a k-loop long enough to hide the later prefetch (>= 9 k-blocks of 64)
    static const bool alias_on = [] { const char* e = getenv("VDQN_ALIAS_OUT"); return e == nullptr || atoi(e) != 0; }();
    a.alias_out = (alias_on && BN == 128 && CK == 64 && a.fast && d->out_scatter < 2 &&
                   (d->residual != nullptr || d->mask_src != nullptr) && k_total >= 9 * 64) ? 1 : 0;
  }
  if (d->out_scatter == 3 && !a.fast)
    return set_error(VDQN_ERR_SHAPE, "conv_gemm: 2x2-block scatter needs the staged epilogue (bf16 output, tile <= 128)");
  if (seg2) {
    rc = make_im2col_map(&epi_maps[3], d->x2, d->N, d->H2, d->W2, d->Cin2, 64, 128, d->stride2, 0, 0, 0, 0, 128);
    if (rc != VDQN_OK) return rc;
  }
  if (a.fast && d->out_scatter < 2) {      // scatter launches move the staged tiles with the LSU
    rc = make_tiled_map_2d(&epi_maps[0], d->out, d->Cout, a.M_total, 32, 32, 64, d->ldc);
    if (rc == VDQN_OK && d->residual)
      rc = make_tiled_map_2d(&epi_maps[1], d->residual, d->Cout, a.M_total, 32, 32, 64, d->ldr);
    if (rc == VDQN_OK && d->mask_src)
      rc = make_tiled_map_2d(&epi_maps[2], d->mask_src, d->Cout, a.M_total, 32, 32, 64, d->ldm);
    if (rc != VDQN_OK) return rc;
  }
  CUtensorMap tmB2 = tmB;
  a.split_m_tile = 0; a.split_cta = 0; a.shift2 = d->shift2;
  if (d->split_n > 0) {
    const long split_m = (long)d->split_n * Ho * Wo;
    if (d->w2 == nullptr || d->split_n >= d->N || split_m % 128 != 0)
      return set_error(VDQN_ERR_SHAPE, "conv_gemm: dual-network launch needs w2 and split_n*Ho*Wo %% 128 == 0");
    rc = make_tiled_map_2d(&tmB2, d->w2, k_total, d->Cout, CK, b_rows, CK == 64 ? 128 : 32);
    if (rc != VDQN_OK) return rc;
    a.split_m_tile = (int)(split_m / 128);
  }
  const int sms = d->max_ctas > 0 && d->max_ctas < dev->num_sms ? d->max_ctas : dev->num_sms;
  if (CK == 16) return launch_igemm<64, 16>(tmA, tmB, tmB2, epi_maps, a, sms, stream);
  if (pair) {
    if (BN == 128) return launch_igemm<128, 64, true>(tmA, tmB, tmB2, epi_maps, a, sms, stream);
    return launch_igemm<256, 64, true>(tmA, tmB, tmB2, epi_maps, a, sms, stream);
  }
  switch (BN) {
    case 64: return launch_igemm<64, 64>(tmA, tmB, tmB2, epi_maps, a, sms, stream);
    case 128: return launch_igemm<128, 64>(tmA, tmB, tmB2, epi_maps, a, sms, stream);
    default: return launch_igemm<256, 64>(tmA, tmB, tmB2, epi_maps, a, sms, stream);
  }
}

VDQN_DEFINE_ROLE_PROFILE_READER(vdqn_debug_role_profile_igemm)
