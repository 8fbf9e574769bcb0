// Hardware probe (run on a B200): can a tcgen05.mma A-operand descriptor start at an arbitrary
// 128-byte row inside a TMA-written, 128B-swizzled "halo" tile, with an 8-row group pitch (SBO)
// that is not a multiple of 1024 bytes?  If yes, a 3x3 convolution can read its nine shifted
// windows out of ONE shared-memory copy of the input tile instead of nine im2col copies from L2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o umma_probe umma_probe.cu
#include "../video_dqn_b200/csrc/ptx.cuh"

#include <cuda_bf16.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

using namespace vdqn;

constexpr int HW_H = 18, HW_W = 10, C = 64, NOUT = 64;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
             float* out, int r, int s, int use_base_offset, int row_pitch_pixels) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;                       // 18*10*128 = 23040 B
  const uint32_t sB = base + 24 * 1024;           // 64 x 128 B
  const uint32_t bar = base + 40 * 1024;
  const uint32_t mma_bar = bar + 8;
  const uint32_t slot = bar + 16;
  uint32_t* slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mma_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, HW_H * HW_W * 128 + NOUT * 128);
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(sA),
        "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(bar), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
    tma_load_2d(sB, &tmW, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t start = sA + (r * row_pitch_pixels + s) * 128;
    const uint32_t idesc = make_idesc_bf16(128, NOUT, 0, 0);
    for (int k = 0; k < 4; ++k) {
      uint64_t ad = make_smem_desc(start + k * 32, 16, row_pitch_pixels * 128, kSwz128);
      if (use_base_offset) ad |= (uint64_t)((start >> 7) & 7) << 49;
      const uint64_t bd = make_smem_desc(sB + k * 32, 16, 1024, kSwz128);
      umma_f16(tmem, ad, bd, idesc, k != 0);
    }
    umma_commit(mma_bar);
  }
  __syncthreads();
  mbar_wait(mma_bar, 0);
  tc_fence_after();
  uint32_t raw[32];
  for (int chunk = 0; chunk < 2; ++chunk) {
    tmem_ld_32x32(tmem + chunk * 32 + ((uint32_t)(warp * 32) << 16), raw);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * NOUT + chunk * 32 + j] = __uint_as_float(raw[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  std::vector<__nv_bfloat16> hx(HW_H * HW_W * C), hw(NOUT * C);
  std::vector<float> fx(hx.size()), fw(hw.size());
  srand(1);
  for (size_t i = 0; i < hx.size(); ++i) { float v = (rand() % 17 - 8) / 8.f; hx[i] = __float2bfloat16(v); fx[i] = v; }
  for (size_t i = 0; i < hw.size(); ++i) { float v = (rand() % 9 - 4) / 4.f; hw[i] = __float2bfloat16(v); fw[i] = v; }
  __nv_bfloat16 *dx, *dw; float* dout;
  cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dw, hw.size() * 2); cudaMalloc(&dout, 128 * NOUT * 4);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmX, tmW;
  {
    cuuint64_t dims[4] = {C, HW_W, HW_H, 1};
    cuuint64_t strides[3] = {C * 2, HW_W * C * 2, (cuuint64_t)HW_H * HW_W * C * 2};
    cuuint32_t box[4] = {C, HW_W, HW_H, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode X: %d\n", (int)r);
    cuuint64_t d2[2] = {C, NOUT}; cuuint64_t s2[1] = {C * 2}; cuuint32_t b2[2] = {C, NOUT}; cuuint32_t e2[2] = {1, 1};
    r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dw, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode W: %d\n", (int)r);
  }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  std::vector<float> hout(128 * NOUT);
  for (int ubo = 0; ubo < 2; ++ubo)
    for (int r = 0; r < 3; ++r)
      for (int s = 0; s < 3; ++s) {
        probe_kernel<<<1, 128, 48 * 1024>>>(tmX, tmW, dout, r, s, ubo, HW_W);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("tap (%d,%d) base_offset=%d: CUDA error %s\n", r, s, ubo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0;
        for (int m = 0; m < 128; ++m) {
          const int g = m / 8, j = m % 8;
          for (int n = 0; n < NOUT; ++n) {
            double acc = 0;
            for (int c = 0; c < C; ++c) acc += (double)fx[((g + r) * HW_W + (j + s)) * C + c] * fw[n * C + c];
            const double err = fabs(acc - hout[m * NOUT + n]);
            if (err > maxerr) maxerr = err;
            if (err > 1e-3) ++bad;
          }
        }
        printf("tap (%d,%d) base_offset=%d: max err %.4g, bad %d/%d %s\n", r, s, ubo, maxerr, bad, 128 * NOUT,
               bad == 0 ? "OK" : "MISMATCH");
      }
  return 0;
}
