// Shared accumulator epilogue of the convolution kernels: one thread owns one output pixel (TMEM
// lane) and 32 consecutive output channels per call.
//   v = acc + shift[c] + residual[m,c];  ReLU;  v = mask_src[m,c] > 0 ? v : 0;  store bf16/fp32;
//   colsum[c] += sum over the warp's 32 pixels (warp transpose-reduce, one atomic per channel).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/vdqn.h"
#include "ptx.cuh"

namespace vdqn {

struct EpiArgs {
  void* out;                      // [opix][ldc] bf16 or fp32
  void* out2;                     // optional second bf16 destination (zero-dilated copy)
  const float* shift;             // [Cout] or null
  const __nv_bfloat16* residual;  // [M][ldr] or null
  const __nv_bfloat16* mask_src;  // [M][ldm] or null
  float* colsum;                  // [Cout] or null
  int ldc, ldr, ldm, out2_ld;
  int flags;                      // VDQN_EPI_*
};

inline EpiArgs make_epi_args(const vdqn_conv_desc* d) {
  EpiArgs e;
  e.out = d->out; e.out2 = d->out2; e.shift = d->shift;
  e.residual = static_cast<const __nv_bfloat16*>(d->residual);
  e.mask_src = static_cast<const __nv_bfloat16*>(d->mask_src);
  e.colsum = d->colsum;
  e.ldc = d->ldc; e.ldr = d->ldr; e.ldm = d->ldm; e.out2_ld = d->out2_ld;
  e.flags = d->flags;
  return e;
}

// warp transpose-reduce: lane L ends up with the sum over the 32 lanes of v[L]
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// raw: 32 fp32 accumulators of this thread's pixel, channels [c0, c0+32).  `m` indexes
// residual / mask_src (compact pixel index), `opix` / `opix2` the destinations.
// Returns this lane's share of the per-channel column sum (channel c0 + lane, summed over the warp's
// 32 pixels) when a.colsum != nullptr; the caller accumulates it across tiles and flushes with
// ONE atomic per channel per CTA (same-address atomics per tile made the data-gradient epilogues
// 2.5x slower than the forward ones).
__device__ __forceinline__ float epilogue_chunk(const EpiArgs& a, const uint32_t (&raw)[32], bool valid,
                                                long m, long opix, long opix2, int c0, int lane) {
  const bool out_f32 = a.flags & VDQN_EPI_OUT_F32;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
  if (a.shift != nullptr) {
    const float4* sp = reinterpret_cast<const float4*>(a.shift + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 s4 = __ldg(sp + j);
      v[4 * j + 0] += s4.x; v[4 * j + 1] += s4.y; v[4 * j + 2] += s4.z; v[4 * j + 3] += s4.w;
    }
  }
  if (a.residual != nullptr && valid) {
    const uint4* rp = reinterpret_cast<const uint4*>(a.residual + m * a.ldr + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 r4 = __ldg(rp + j);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        v[8 * j + 2 * e] += f.x;
        v[8 * j + 2 * e + 1] += f.y;
      }
    }
  }
  if (a.flags & VDQN_EPI_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  if (a.mask_src != nullptr && valid) {
    const uint4* mp = reinterpret_cast<const uint4*>(a.mask_src + m * a.ldm + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 r4 = __ldg(mp + j);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        if (!(f.x > 0.f)) v[8 * j + 2 * e] = 0.f;
        if (!(f.y > 0.f)) v[8 * j + 2 * e + 1] = 0.f;
      }
    }
  }
  if (!valid) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
  }
  if (valid) {
    if (out_f32) {
      float4* op = reinterpret_cast<float4*>(static_cast<float*>(a.out) + opix * a.ldc + c0);
#pragma unroll
      for (int j = 0; j < 8; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
      uint4 pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk[j]);
#pragma unroll
        for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
      }
      uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(a.out) + opix * a.ldc + c0);
#pragma unroll
      for (int j = 0; j < 4; ++j) op[j] = pk[j];
      if (a.out2 != nullptr) {
        uint4* op2 = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(a.out2) + opix2 * a.out2_ld + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j) op2[j] = pk[j];
      }
    }
  }
  if (a.colsum != nullptr) {
    // round to the stored precision first so d beta matches what the weight gradient consumes
    if (!out_f32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
    }
    // warp transpose-reduce: afterwards lane L holds the column-(c0+L) sum over 32 pixels
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const float send = upper ? v[i] : v[i + off];
        const float keep = upper ? v[i + off] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
      }
    }
    return v[0];
  }
  return 0.f;
}

// ---------------------------------------------------------------------------------------------
// Staged ("fast") epilogue: global traffic goes through per-warp shared-memory tiles of
// 32 pixels x 64 channels (bf16, 128-byte rows, 128B swizzle) moved by TMA -- residual / mask
// tiles are TMA-loaded, the output tile is TMA-stored -- so the LSU only sees conflict-free
// 16-byte shared accesses instead of 32 different cache lines per instruction.
// One call handles 32 accumulator columns [half*32, half*32+32) of the warp's current 64-column
// group: `stg_*` are the warp's staging tiles (1024-byte aligned), `c0` the first global channel.
//
// ROW_BYTES = 128: 64-column tiles (128B swizzle), `half` selects the 32 columns inside the row.
// ROW_BYTES = 64: the tile IS 32 columns wide (64-byte rows, 64B swizzle; `half` unused) -- used
// when every epilogue warp owns its own 32-column slice (two warps per TMEM lane quadrant).
// `row_acc` (optional, 32 floats per thread): column sums are then NOT reduced per call -- the
// thread adds its own row's (rounded) values into row_acc and the caller reduces once at the end of
// the kernel with warp_transpose_reduce (the 31 shuffles per tile were a third of the data-gradient
// epilogue).
// EPI: which optional steps are COMPILED IN (EPI_HAS_* bits; each still checks its pointer / flag at
// run time).  ptxas if-converts the optional blocks into predicated instructions, so a forward
// launch compiled with everything present issues the whole residual + mask code with the predicate
// off (measured: 310 instead of ~130 instructions per 32 columns, and the epilogue warps are the
// longest role of the small-K kernels).  Callers therefore dispatch ONCE per kernel to a loop
// specialised for the steps the launch actually has (epi_mode()).
enum : int { EPI_HAS_RES = 1, EPI_HAS_MASK = 2, EPI_HAS_COLSUM = 4, EPI_HAS_SHIFT_RELU = 8, EPI_HAS_ALL = 15 };

__device__ __forceinline__ int epi_mode(const EpiArgs& a) {
  return (a.residual != nullptr ? EPI_HAS_RES : 0) | (a.mask_src != nullptr ? EPI_HAS_MASK : 0) |
         (a.colsum != nullptr ? EPI_HAS_COLSUM : 0) |
         ((a.shift != nullptr || (a.flags & VDQN_EPI_RELU)) ? EPI_HAS_SHIFT_RELU : 0);
}

template <int V>
struct EpiMode { static constexpr int value = V; };

// run `f(EpiMode<m>{})` for the specialisation matching `mode` (the common combinations of this
// network; anything else takes the generic all-steps version)
template <typename F>
__device__ __forceinline__ void epi_dispatch(int mode, F&& f) {
  switch (mode) {
    case EPI_HAS_SHIFT_RELU: f(EpiMode<EPI_HAS_SHIFT_RELU>{}); break;                                  // conv + BN (+ ReLU)
    case EPI_HAS_SHIFT_RELU | EPI_HAS_RES: f(EpiMode<EPI_HAS_SHIFT_RELU | EPI_HAS_RES>{}); break;      // + identity
    case EPI_HAS_MASK | EPI_HAS_COLSUM: f(EpiMode<EPI_HAS_MASK | EPI_HAS_COLSUM>{}); break;            // data gradient
    case EPI_HAS_RES | EPI_HAS_MASK | EPI_HAS_COLSUM: f(EpiMode<EPI_HAS_RES | EPI_HAS_MASK | EPI_HAS_COLSUM>{}); break;
    case 0: f(EpiMode<0>{}); break;                                                                    // plain GEMM
    default: f(EpiMode<EPI_HAS_ALL>{}); break;
  }
}

template <int ROW_BYTES = 128, int EPI = EPI_HAS_ALL>
__device__ __forceinline__ float epilogue_half_staged(const EpiArgs& a, const uint32_t (&raw)[32], bool valid,
                                                      int c0, int half, int lane, uint32_t stg_out,
                                                      uint32_t stg_res, uint32_t stg_mask,
                                                      float* row_acc = nullptr, const float* shift_r = nullptr) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
  if ((EPI & EPI_HAS_SHIFT_RELU) && shift_r != nullptr) {
    // the caller keeps its 32 shift values in registers across tiles (zeros when there is no shift)
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += shift_r[j];
  } else if ((EPI & EPI_HAS_SHIFT_RELU) && a.shift != nullptr) {
    const float4* sp = reinterpret_cast<const float4*>(a.shift + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 s4 = __ldg(sp + j);
      v[4 * j + 0] += s4.x; v[4 * j + 1] += s4.y; v[4 * j + 2] += s4.z; v[4 * j + 3] += s4.w;
    }
  }
  const uint32_t row_off = (uint32_t)lane * (uint32_t)ROW_BYTES;
  const uint32_t sw = ROW_BYTES == 128 ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);
  if (ROW_BYTES == 64) half = 0;
  if ((EPI & EPI_HAS_RES) && a.residual != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 r4 = lds128(stg_res + row_off + ((((uint32_t)(half * 4 + j)) ^ sw) << 4));
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r4);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        v[8 * j + 2 * e] += f.x;
        v[8 * j + 2 * e + 1] += f.y;
      }
    }
  }
  if ((EPI & EPI_HAS_SHIFT_RELU) && (a.flags & VDQN_EPI_RELU)) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  // rows outside the image / matrix are never stored (TMA clips, the LSU copies check bounds): they
  // only have to be zero for the column sums.  Few tiles have such rows: warp-uniform branch.
  if ((EPI & EPI_HAS_COLSUM) && __any_sync(0xffffffffu, !valid)) {
    if (!valid) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint4 pk;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
    // the ReLU mask is applied to the PACKED values: one packed compare (0xFFFF per half where the stored
    // activation is > 0) and one AND per channel pair instead of two unpacks, two compares and two selects
    if ((EPI & EPI_HAS_MASK) && a.mask_src != nullptr) {
      const uint4 r4 = lds128(stg_mask + row_off + ((((uint32_t)(half * 4 + j)) ^ sw) << 4));
      const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&r4);
      const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
      pk.x &= __hgt2_mask(mh[0], zero2); pk.y &= __hgt2_mask(mh[1], zero2);
      pk.z &= __hgt2_mask(mh[2], zero2); pk.w &= __hgt2_mask(mh[3], zero2);
    }
    sts128(stg_out + row_off + ((((uint32_t)(half * 4 + j)) ^ sw) << 4), pk);
    if ((EPI & EPI_HAS_COLSUM) && a.colsum != nullptr) {
      // column sums use the values as stored (rounded), so d beta matches what the weight gradient consumes
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        v[8 * j + 2 * e] = f.x;
        v[8 * j + 2 * e + 1] = f.y;
      }
    }
  }
  if ((EPI & EPI_HAS_COLSUM) && a.colsum != nullptr) {
    if (row_acc != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; ++j) row_acc[j] += v[j];
      return 0.f;
    }
    return warp_transpose_reduce(v, lane);
  }
  return 0.f;
}

// Direct epilogue with its inputs already in registers (BN = 256 tiles, whose shared memory all goes to the
// pipeline): the caller loads a chunk's residual / mask rows (64 bytes per lane each) one chunk AHEAD, so the
// global-load latency that the plain direct epilogue exposes per chunk (role profile: 23.7 k cycles per
// 128 x 256 tile against 18.4 k of MMA issue) overlaps the previous chunk's math.  bf16 compact output only.
struct EpiRegs {
  uint4 res[4], mask[4];
};

template <int EPI>
__device__ __forceinline__ void epilogue_load_regs(const EpiArgs& a, bool valid, long m, int c0, EpiRegs& r) {
  if ((EPI & EPI_HAS_RES) && a.residual != nullptr) {
    const uint4* rp = reinterpret_cast<const uint4*>(a.residual + m * a.ldr + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) r.res[j] = valid ? __ldg(rp + j) : make_uint4(0u, 0u, 0u, 0u);
  }
  if ((EPI & EPI_HAS_MASK) && a.mask_src != nullptr) {
    const uint4* mp = reinterpret_cast<const uint4*>(a.mask_src + m * a.ldm + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) r.mask[j] = valid ? __ldg(mp + j) : make_uint4(0u, 0u, 0u, 0u);
  }
}

template <int EPI>
__device__ __forceinline__ float epilogue_chunk_regs(const EpiArgs& a, const uint32_t (&raw)[32], bool valid,
                                                     long opix, int c0, int lane, const EpiRegs& r) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
  if ((EPI & EPI_HAS_SHIFT_RELU) && a.shift != nullptr) {
    const float4* sp = reinterpret_cast<const float4*>(a.shift + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 s4 = __ldg(sp + j);
      v[4 * j + 0] += s4.x; v[4 * j + 1] += s4.y; v[4 * j + 2] += s4.z; v[4 * j + 3] += s4.w;
    }
  }
  if ((EPI & EPI_HAS_RES) && a.residual != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.res[j]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        v[8 * j + 2 * e] += f.x;
        v[8 * j + 2 * e + 1] += f.y;
      }
    }
  }
  if ((EPI & EPI_HAS_SHIFT_RELU) && (a.flags & VDQN_EPI_RELU)) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
  }
  uint4 pk[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk[j]);
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[8 * j + 2 * e], v[8 * j + 2 * e + 1]);
    if ((EPI & EPI_HAS_MASK) && a.mask_src != nullptr) {
      const __nv_bfloat162* mh = reinterpret_cast<const __nv_bfloat162*>(&r.mask[j]);
      const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
      pk[j].x &= __hgt2_mask(mh[0], zero2); pk[j].y &= __hgt2_mask(mh[1], zero2);
      pk[j].z &= __hgt2_mask(mh[2], zero2); pk[j].w &= __hgt2_mask(mh[3], zero2);
    }
  }
  if (valid) {
    uint4* op = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(a.out) + opix * a.ldc + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) op[j] = pk[j];
  }
  if ((EPI & EPI_HAS_COLSUM) && a.colsum != nullptr) {
    // the values as stored (rounded); rows outside the matrix contribute nothing
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk[j]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        v[8 * j + 2 * e] = valid ? f.x : 0.f;
        v[8 * j + 2 * e + 1] = valid ? f.y : 0.f;
      }
    }
    return warp_transpose_reduce(v, lane);
  }
  return 0.f;
}

// bf16 output: the staged epilogue applies -- to a compact destination through TMA, to a
// zero-dilated (scatter) destination through per-row LSU copies, provided residual / mask are
// indexed by the scattered pixel as well (or absent)
inline bool fast_epilogue_ok(const vdqn_conv_desc* d) {
  const bool has_in = d->residual != nullptr || d->mask_src != nullptr;
  if (d->out_scatter >= 2 && has_in && !(d->flags & VDQN_EPI_SCATTER_INPUTS)) return false;
  return !(d->flags & VDQN_EPI_OUT_F32) && d->out2 == nullptr &&
         d->ldc % 8 == 0 && (d->residual == nullptr || d->ldr % 8 == 0) &&
         (d->mask_src == nullptr || d->ldm % 8 == 0);
}

}  // namespace vdqn
