// Gradient exchange of the data-parallel step as ONE kernel over NVLink / NVSwitch peer memory (SURVEY 8e;
// the reference itself is single-GPU, train_q_network.py:275).
//
// Every rank keeps its flat fp32 gradient arena in symmetric memory (the same allocation mapped into every
// process).  After the backward pass each rank launches this kernel; rank r owns elements
// [r n/W, (r+1) n/W):
//   barrier A   every rank's gradients are complete (flags written into the peers' memory, release/acquire
//               at system scope)
//   reduce      NVSwitch multicast: one `multimem.ld_reduce` per 16 bytes (the switch adds the W copies);
//               plain peer-to-peer otherwise: W loads, summed in rank order
//   broadcast   `multimem.st` (or W peer stores) of the sum into every rank's arena
//   barrier B   all sums have landed everywhere; the Adam kernel that follows reads the local arena
// Each element is summed by exactly one rank and copied, so all ranks hold bit-identical gradients.
// 49.7 MB per rank and step: ~50 MB in and out per GPU instead of the 2 (W-1)/W x 49.7 MB x latency-bound
// ring / tree steps NCCL schedules between the SMs that persistent conv kernels leave free; and no
// collective node inside the captured step graph.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"
#include "vdqn_internal.h"

namespace vdqn {

constexpr int kNvlMaxWorld = 16;

struct NvlArgs {
  float* peer[kNvlMaxWorld];          // this process' mapping of every rank's arena
  uint32_t* flags[kNvlMaxWorld];      // ... of every rank's flag block: [0, W) barrier A, [W, 2W) barrier B
  float* mc;                          // multicast address of the arena, or nullptr
  uint32_t* epoch;                    // local: number of exchanges done
  uint32_t* counter;                  // local: blocks that finished the copy phase
  long n4;                            // float4 elements per rank slice
  long first;                         // first element of the exchanged range
  int rank, world;
  int flag0;                          // first flag slot of this exchange's channel (2 W slots per channel)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_v4(float* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 multimem_ld_reduce_v4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_v4(float* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// bounded: a rank that never arrives must surface as a trapped launch, not as a hung GPU
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t epoch) {
  uint32_t spins = 0;
  while ((int32_t)(ld_acquire_sys(p) - epoch) < 0) {
    if (++spins > (1u << 26)) {
      printf("vdqn: nvl_allreduce: a peer did not arrive (block %d thread %d epoch %u)\n", blockIdx.x, threadIdx.x, epoch);
      __trap();
    }
    __nanosleep(64);
  }
}

template <bool MULTIMEM>
__global__ void __launch_bounds__(512) nvl_allreduce_kernel(const NvlArgs a) {
  // (launched with 512 threads when it has the GPU to itself, with 128 when it runs next to the backward pass:
  //  a CTA of 128 threads x ~40 registers fits beside a persistent conv CTA)
  __shared__ uint32_t s_epoch, s_last;
  if (threadIdx.x == 0) s_epoch = *a.epoch + 1;
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const int W = a.world;
  // ---- barrier A: block 0 tells every peer "my gradients are complete", every block waits for all peers
  if (blockIdx.x == 0 && threadIdx.x < W) {
    __threadfence_system();
    st_release_sys(a.flags[threadIdx.x] + a.flag0 + a.rank, epoch);
  }
  if (threadIdx.x < W) wait_flag(a.flags[a.rank] + a.flag0 + threadIdx.x, epoch);
  __syncthreads();
  // ---- reduce this rank's slice, broadcast the sums.  The loop is bound by bytes in flight over a ~3 us
  // NVLink round trip: every thread keeps UNROLL (x W for the peer-to-peer form) 16-byte loads outstanding.
  const long base = a.first / 4 + (long)a.rank * a.n4;
  const long stride = (long)gridDim.x * blockDim.x;
  constexpr int UNROLL = MULTIMEM ? 8 : 2;
  for (long i0 = blockIdx.x * (long)blockDim.x + threadIdx.x; i0 < a.n4; i0 += stride * UNROLL) {
    if (MULTIMEM) {
      float4 s[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (i0 + u * stride < a.n4) s[u] = multimem_ld_reduce_v4(a.mc + (base + i0 + u * stride) * 4);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (i0 + u * stride < a.n4) multimem_st_v4(a.mc + (base + i0 + u * stride) * 4, s[u]);
    } else {
      float4 v[UNROLL][kNvlMaxWorld];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int r = 0; r < kNvlMaxWorld; ++r)
          if (r < W && i0 + u * stride < a.n4) v[u][r] = ld_relaxed_sys_v4(a.peer[r] + (base + i0 + u * stride) * 4);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (i0 + u * stride < a.n4) {
          float4 s = v[u][0];
#pragma unroll
          for (int r = 1; r < kNvlMaxWorld; ++r)
            if (r < W) { s.x += v[u][r].x; s.y += v[u][r].y; s.z += v[u][r].z; s.w += v[u][r].w; }
#pragma unroll
          for (int r = 0; r < kNvlMaxWorld; ++r)
            if (r < W) st_relaxed_sys_v4(a.peer[r] + (base + i0 + u * stride) * 4, s);
        }
      }
    }
  }
  // ---- barrier B: the last block of this rank to finish signals the peers and waits for theirs
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x < W) {
      __threadfence_system();
      st_release_sys(a.flags[threadIdx.x] + a.flag0 + W + a.rank, epoch);
      wait_flag(a.flags[a.rank] + a.flag0 + W + threadIdx.x, epoch);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      *a.counter = 0u;
      *a.epoch = epoch;
    }
  }
}

}  // namespace vdqn

using namespace vdqn;

extern "C" int vdqn_nvl_allreduce(const vdqn_nvl_desc* d, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (d == nullptr || d->peer_bufs == nullptr || d->peer_flags == nullptr || d->epoch == nullptr || d->counter == nullptr)
    return set_error(VDQN_ERR_ARG, "nvl_allreduce: null pointer");
  if (d->world < 2 || d->world > kNvlMaxWorld || d->rank < 0 || d->rank >= d->world)
    return set_error(VDQN_ERR_ARG, "nvl_allreduce: 2 <= world <= 16");
  if (d->n < 1 || d->n % (4L * d->world) != 0 || d->first < 0 || d->first % 4 != 0)
    return set_error(VDQN_ERR_SHAPE, "nvl_allreduce: the range length must be a multiple of 4 * world, its start of 4");
  if (d->channel < 0 || d->channel > 3) return set_error(VDQN_ERR_ARG, "nvl_allreduce: channel 0..3");
  DeviceInfo* dev = device_info();
  if (dev == nullptr) return VDQN_ERR_CUDA;
  NvlArgs a{};
  for (int r = 0; r < d->world; ++r) {
    if (d->peer_bufs[r] == nullptr || d->peer_flags[r] == nullptr || (reinterpret_cast<uintptr_t>(d->peer_bufs[r]) & 15))
      return set_error(VDQN_ERR_ARG, "nvl_allreduce: bad peer pointer %d", r);
    a.peer[r] = static_cast<float*>(d->peer_bufs[r]);
    a.flags[r] = static_cast<uint32_t*>(d->peer_flags[r]);
  }
  a.mc = static_cast<float*>(d->multicast_ptr);
  a.epoch = d->epoch; a.counter = d->counter;
  a.n4 = d->n / (4L * d->world);
  a.first = d->first;
  a.rank = d->rank; a.world = d->world;
  a.flag0 = d->channel * 2 * d->world;
  const int threads = d->threads == 128 || d->threads == 256 ? d->threads : 512;
  int grid = d->max_ctas > 0 && d->max_ctas < dev->num_sms ? d->max_ctas : dev->num_sms;
  const long need = (a.n4 + threads - 1) / threads;
  if (grid > need) grid = (int)need;
  if (a.mc != nullptr) launch_kernel(nvl_allreduce_kernel<true>, grid, threads, 0, stream, a);
  else launch_kernel(nvl_allreduce_kernel<false>, grid, threads, 0, stream, a);
  VDQN_CHECK_LAUNCH("nvl_allreduce");
  return VDQN_OK;
}
