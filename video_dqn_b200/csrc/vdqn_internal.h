// Internal helpers shared by the translation units of libvdqn.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vdqn.h"

namespace vdqn {

struct DeviceInfo {
  int device;
  int num_sms;
};

// printf-style; stores a thread-local message and returns `code`.
int set_error(int code, const char* fmt, ...);
// Lazily initialised (vdqn_init on the current device); nullptr + error on failure.
DeviceInfo* device_info();

// NHWC bf16 activation tensor, TMA im2col mode.  lower/upper corners are the bounding box of
// the filter window's base pixel (lower = -pad, upper = pad_hi - (R-1)*dil).
int make_im2col_map(CUtensorMap* map, const void* base, int N, int H, int W, int C,
                    int channels_per_pixel, int pixels_per_column, int stride,
                    int lower_h, int lower_w, int upper_h, int upper_w, int swizzle_bytes);
// Row-major bf16 matrix [rows][cols] (cols contiguous), box [box_rows][box_cols].
int make_tiled_map_2d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows,
                      uint32_t box_cols, uint32_t box_rows, int swizzle_bytes,
                      uint64_t row_stride_elems = 0);

// NHWC bf16 activation tensor, plain tiled mode, box = [1][box_h][box_w][box_c] (out-of-image
// coordinates, including negative ones, are zero-filled).
int make_tiled_map_nhwc(CUtensorMap* map, const void* base, int N, int H, int W, int C, int box_c,
                        int box_w, int box_h, int swizzle_bytes);

// halo_conv.cu
bool halo_conv_supported(const vdqn_conv_desc* d);
int halo_conv_launch(const vdqn_conv_desc* d, cudaStream_t stream);

// halo_wgrad.cu: 64->64 3x3 weight gradient from one smem copy of the input window;
// writes part[splits][64][576] with splits = number of CTAs launched
bool halo_wgrad_supported(const vdqn_wgrad_desc* d);
int halo_wgrad_launch(const vdqn_wgrad_desc* d, cudaStream_t stream);

// halo_wgrad_stem.cu: packed-stem weight gradient; writes part[splits][64][256]
bool halo_wgrad_stem_supported(const vdqn_wgrad_desc* d);
int halo_wgrad_stem_launch(const vdqn_wgrad_desc* d, cudaStream_t stream);

// td_bulk.cu: streaming TD epilogue over the first n_chunks * 1024 elements (Bellman branch)
bool td_bulk_supported(const vdqn_td_desc* d);
int td_bulk_launch(const vdqn_td_desc* d, long n_chunks, cudaStream_t stream);

inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

// every kernel launch of the library is counted (bench.py reports it as `gpu_launches`)
void count_launch();

#define VDQN_CHECK_LAUNCH(what)                                                         \
  do {                                                                                  \
    ::vdqn::count_launch();                                                             \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess)                                                             \
      return ::vdqn::set_error(VDQN_ERR_CUDA, what ": %s", cudaGetErrorString(e__));    \
  } while (0)

// Programmatic dependent launch (opt-in: VDQN_PDL=1 in the environment): the kernel may start its
// prologue before the previous kernel on the stream has finished; every kernel calls pdl_wait()
// (ptx.cuh) before its first global-memory access, so ordering is unchanged.  Measured on the
// B = 256 step (CUDA-graph replay): 7.66 ms with, 7.55 ms without -- the persistent kernels fill
// every SM, so an early-resident successor only takes issue slots from the tail; hence off by default.
bool pdl_enabled();
// CTA-pair (tcgen05 cta_group::2) kernels are used where a shape allows it unless VDQN_PAIR=0
bool pair_default();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// thread-block cluster of `cluster_x` consecutive CTAs (CTA pairs for cta_group::2 kernels)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t stream, unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace vdqn
