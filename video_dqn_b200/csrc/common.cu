// Error reporting, device query and TMA tensor-map construction for libvdqn.so.
// The driver entry points are resolved through cudaGetDriverEntryPoint so the library has no
// link-time dependency on libcuda (it must load on a machine without a driver).
#include "vdqn_internal.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>

namespace vdqn {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// VDQN_PAIR=0 in the environment keeps every launch on the single-CTA kernels
bool pair_default() {
  static const bool on = [] {
    const char* e = getenv("VDQN_PAIR");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("VDQN_PDL");
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static std::mutex g_mu;
static bool g_ready = false;
static DeviceInfo g_dev{-1, 0};
static EncodeTiledFn g_encode_tiled = nullptr;
static EncodeIm2colFn g_encode_im2col = nullptr;

static int do_init(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_ready && (device < 0 || device == g_dev.device)) return VDQN_OK;
  cudaError_t e;
  if (device >= 0) {
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return set_error(VDQN_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  }
  int cur = 0;
  e = cudaGetDevice(&cur);
  if (e != cudaSuccess) return set_error(VDQN_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, cur);
  if (e != cudaSuccess) return set_error(VDQN_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return set_error(VDQN_ERR_CUDA, "device %d is sm_%d%d; libvdqn is built for sm_100a only (no fallback)",
                     cur, prop.major, prop.minor);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess)
    return set_error(VDQN_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  g_encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  fn = nullptr;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess)
    return set_error(VDQN_ERR_DRIVER, "cuTensorMapEncodeIm2col not available from the driver");
  g_encode_im2col = reinterpret_cast<EncodeIm2colFn>(fn);
  g_dev.device = cur;
  g_dev.num_sms = prop.multiProcessorCount;
  g_ready = true;
  return VDQN_OK;
}

DeviceInfo* device_info() {
  if (!g_ready && do_init(-1) != VDQN_OK) return nullptr;
  return &g_dev;
}

static CUtensorMapSwizzle swz_enum(int bytes) {
  switch (bytes) {
    case 128: return CU_TENSOR_MAP_SWIZZLE_128B;
    case 64: return CU_TENSOR_MAP_SWIZZLE_64B;
    case 32: return CU_TENSOR_MAP_SWIZZLE_32B;
    default: return CU_TENSOR_MAP_SWIZZLE_NONE;
  }
}

int make_im2col_map(CUtensorMap* map, const void* base, int N, int H, int W, int C,
                    int channels_per_pixel, int pixels_per_column, int stride, int lower_h,
                    int lower_w, int upper_h, int upper_w, int swizzle_bytes) {
  if (device_info() == nullptr) return VDQN_ERR_CUDA;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const int lower[2] = {lower_w, lower_h};
  const int upper[2] = {upper_w, upper_h};
  const cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_encode_im2col(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base),
                               dims, strides, lower, upper, (cuuint32_t)channels_per_pixel,
                               (cuuint32_t)pixels_per_column, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swz_enum(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(VDQN_ERR_DRIVER,
                     "cuTensorMapEncodeIm2col failed (%d): N=%d H=%d W=%d C=%d cpp=%d ppc=%d stride=%d "
                     "lower=(%d,%d) upper=(%d,%d)",
                     (int)r, N, H, W, C, channels_per_pixel, pixels_per_column, stride, lower_h, lower_w,
                     upper_h, upper_w);
  return VDQN_OK;
}

int make_tiled_map_nhwc(CUtensorMap* map, const void* base, int N, int H, int W, int C, int box_c,
                        int box_w, int box_h, int swizzle_bytes) {
  if (device_info() == nullptr) return VDQN_ERR_CUDA;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims,
                              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz_enum(swizzle_bytes),
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(VDQN_ERR_DRIVER, "cuTensorMapEncodeTiled(4d) failed (%d): N=%d H=%d W=%d C=%d box=(%d,%d,%d)",
                     (int)r, N, H, W, C, box_c, box_w, box_h);
  return VDQN_OK;
}

int make_tiled_map_2d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows,
                      uint32_t box_cols, uint32_t box_rows, int swizzle_bytes,
                      uint64_t row_stride_elems) {
  if (device_info() == nullptr) return VDQN_ERR_CUDA;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {(row_stride_elems ? row_stride_elems : cols) * 2};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_tiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              swz_enum(swizzle_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(VDQN_ERR_DRIVER, "cuTensorMapEncodeTiled failed (%d): cols=%llu rows=%llu box=(%u,%u)",
                     (int)r, (unsigned long long)cols, (unsigned long long)rows, box_cols, box_rows);
  return VDQN_OK;
}

}  // namespace vdqn

extern "C" const char* vdqn_last_error(void) { return vdqn::g_err; }
extern "C" int vdqn_abi_version(void) { return VDQN_ABI_VERSION; }
extern "C" int vdqn_init(int device) { return vdqn::do_init(device); }
extern "C" int vdqn_num_sms(void) {
  vdqn::DeviceInfo* d = vdqn::device_info();
  return d ? d->num_sms : -1;
}
extern "C" long long vdqn_launch_count(void) { return vdqn::g_launches.load(); }

extern "C" int vdqn_zero(void* ptr, int64_t bytes, void* stream_v) {
  if (ptr == nullptr || bytes < 0) return vdqn::set_error(VDQN_ERR_ARG, "zero: bad argument");
  if (bytes == 0) return VDQN_OK;
  cudaError_t e = cudaMemsetAsync(ptr, 0, (size_t)bytes, static_cast<cudaStream_t>(stream_v));
  if (e != cudaSuccess) return vdqn::set_error(VDQN_ERR_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
  return VDQN_OK;
}
