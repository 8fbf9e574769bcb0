// TD epilogue, streaming form (large batches): the loss arithmetic of process_batch
// (train_q_network.py:134-180) with every array moved by the bulk-copy engine.
//
// The per-thread form (elementwise.cu) is LSU-bound, not DRAM-bound: six concurrent per-thread streams
// (three Q arrays, two 8-byte label arrays, dQ) reached 63-68 % of the measured copy bandwidth at
// B = 2^20.  Here a persistent CTA walks chunks of 1024 (sample, class) elements; one thread issues ONE
// `cp.async.bulk` per array and chunk into a three-stage shared-memory ring (mbarrier complete_tx), all
// threads compute from shared memory (a thread's A actions are a stride-A pattern: conflict-free for the
// odd A of this network), dQ goes back through a double-buffered shared tile and one bulk store per
// chunk.  ~50 KB in flight per SM without a single LSU global access except `act` (8 bytes per sample,
// read directly: its chunk start is only 8-byte aligned).
//
// Element -> arithmetic is that of td_epilogue_kernel (same fp32 operations in the same order), so y,
// arg-max and dQ are bit-identical to the per-thread kernels; the loss is the same set of terms summed
// in another order.  Full chunks only: the caller runs the per-thread kernel on the tail.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"
#include "vdqn_internal.h"

namespace vdqn {

constexpr int kTdBulkChunk = 1024;      // elements (b, c) per chunk
constexpr int kTdBulkThreads = 1024;    // one element per thread and chunk: 32 warps hide the shared-memory / act latencies
constexpr int kTdBulkStages = 3;
constexpr int kTdBulkMaxSmem = 224 * 1024;   // dynamic part; the kernel also has ~1.3 KB of static shared memory

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(dst)),
               "r"(src), "r"(bytes)
               : "memory");
}

struct TdBulkLayout {
  int q_bytes, l_bytes, stage_bytes, n_q, n_l;
};

__host__ __device__ inline TdBulkLayout td_bulk_layout(const vdqn_td_desc& d) {
  TdBulkLayout L;
  L.q_bytes = kTdBulkChunk * d.A * 4;
  L.l_bytes = kTdBulkChunk * (d.labels_f32 ? 4 : 8);
  L.n_q = d.double_dqn ? 3 : 2;
  L.n_l = d.use_valid ? 3 : 2;
  L.stage_bytes = L.n_q * L.q_bytes + L.n_l * L.l_bytes;
  return L;
}

// A_T: the number of actions as a compile-time constant (0: read d.A).  With a run-time A the arg-max / dQ
// loops stay loops and the first version of this kernel issued ~150 instructions per element from 8 warps:
// ncu showed 28 % issue utilisation, DRAM 42 % busy -- bound by its own instruction latencies, not by HBM.
template <int A_T>
__global__ void __launch_bounds__(kTdBulkThreads, 1) td_epilogue_bulk_kernel(const vdqn_td_desc d, const long n_chunks) {
  extern __shared__ __align__(128) uint8_t td_smem[];
  __shared__ __align__(8) uint64_t full_bar[kTdBulkStages];
  __shared__ float red[kTdBulkThreads / 32];
  const TdBulkLayout L = td_bulk_layout(d);
  const int A = A_T > 0 ? A_T : d.A;
  const uint32_t smem0 = smem_u32(td_smem);
  const uint32_t out0 = smem0 + kTdBulkStages * L.stage_bytes;           // two dQ tiles
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTdBulkStages; ++s) mbar_init(smem_u32(&full_bar[s]), 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();

  auto issue = [&](long chunk, int s) {           // one thread
    const uint32_t bar = smem_u32(&full_bar[s]);
    const long e0 = chunk * kTdBulkChunk;
    uint32_t dst = smem0 + s * L.stage_bytes;
    mbar_expect_tx(bar, (uint32_t)L.stage_bytes);
    bulk_load(dst, d.q_s + e0 * A, L.q_bytes, bar); dst += L.q_bytes;
    bulk_load(dst, d.q_next_target + e0 * A, L.q_bytes, bar); dst += L.q_bytes;
    if (d.double_dqn) { bulk_load(dst, d.q_next_online + e0 * A, L.q_bytes, bar); dst += L.q_bytes; }
    const long lo = e0 * (d.labels_f32 ? 4 : 8);
    bulk_load(dst, reinterpret_cast<const uint8_t*>(d.rew) + lo, L.l_bytes, bar); dst += L.l_bytes;
    bulk_load(dst, reinterpret_cast<const uint8_t*>(d.term) + lo, L.l_bytes, bar); dst += L.l_bytes;
    if (d.use_valid) bulk_load(dst, reinterpret_cast<const uint8_t*>(d.valid) + lo, L.l_bytes, bar);
  };
  const long first = blockIdx.x, stride = gridDim.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kTdBulkStages; ++s)
      if (first + s * stride < n_chunks) issue(first + s * stride, s);
  }
  float local = 0.f;
  long it = 0;
  for (long chunk = first; chunk < n_chunks; chunk += stride, ++it) {
    const int s = (int)(it % kTdBulkStages);
    const uint32_t parity = (uint32_t)((it / kTdBulkStages) & 1);
    const long e0 = chunk * kTdBulkChunk;
    // `act` straight from global memory, in flight while the stage lands
    int act[kTdBulkChunk / kTdBulkThreads];
    const long b0 = e0 / d.C;                              // one 64-bit division per chunk, the rest in 32 bits
    const uint32_t r0 = (uint32_t)(e0 - b0 * d.C);
#pragma unroll
    for (int k = 0; k < kTdBulkChunk / kTdBulkThreads; ++k)
      act[k] = (int)d.act[b0 + (r0 + (uint32_t)(k * kTdBulkThreads) + threadIdx.x) / (uint32_t)d.C];
    if (threadIdx.x == 0) tma_store_wait_read<1>();        // the dQ tile of chunk it-2 has left shared memory
    mbar_wait(smem_u32(&full_bar[s]), parity);
    const uint8_t* st = td_smem + s * L.stage_bytes;
    const float* sm_qs = reinterpret_cast<const float*>(st);
    const float* sm_qt = reinterpret_cast<const float*>(st + L.q_bytes);
    const float* sm_qo = d.double_dqn ? reinterpret_cast<const float*>(st + 2 * L.q_bytes) : sm_qt;
    const uint8_t* sm_rew = st + L.n_q * L.q_bytes;
    const uint8_t* sm_term = sm_rew + L.l_bytes;
    const uint8_t* sm_valid = sm_term + L.l_bytes;
    float g[kTdBulkChunk / kTdBulkThreads];
#pragma unroll
    for (int k = 0; k < kTdBulkChunk / kTdBulkThreads; ++k) {
      const int e = k * kTdBulkThreads + threadIdx.x;
      const float* qs = sm_qs + e * A;
      const float* qt = sm_qt + e * A;
      const float* qsel = sm_qo + e * A;
      int best = 0;
      float bv = qsel[0];
#pragma unroll
      for (int a = 1; a < A; ++a) {
        const float v = qsel[a];
        if (v > bv) { bv = v; best = a; }            // strict > : first maximum wins (torch.argmax)
      }
      float term, rew, mask = 1.f;
      if (d.labels_f32) {
        term = reinterpret_cast<const float*>(sm_term)[e];
        rew = reinterpret_cast<const float*>(sm_rew)[e];
        if (d.use_valid) mask = reinterpret_cast<const float*>(sm_valid)[e];
      } else {
        term = (float)reinterpret_cast<const int64_t*>(sm_term)[e];
        rew = (float)reinterpret_cast<const int64_t*>(sm_rew)[e];
        if (d.use_valid) mask = (float)reinterpret_cast<const int64_t*>(sm_valid)[e];
      }
      const float q_a = qt[best] * (1.f - term);
      float y = d.linear ? rew + (q_a - 0.1f) : rew + d.gamma * q_a;
      if (d.clip_rect) y = fminf(fmaxf(y, 0.f), 1.f);
      const float diff = qs[act[k]] - y;
      float l = 0.5f * diff * diff;
      if (d.use_valid) l *= mask;
      local += l;
      g[k] = diff * mask * d.inv_count;
      if (d.best_out != nullptr) d.best_out[e0 + e] = best;
      if (d.y_out != nullptr) d.y_out[e0 + e] = y;
    }
    __syncthreads();      // every read of stage s is done; thread 0 is past its wait on the dQ tile
    if (threadIdx.x == 0 && chunk + kTdBulkStages * stride < n_chunks) issue(chunk + kTdBulkStages * stride, s);
    if (d.dq != nullptr) {
      float* so = reinterpret_cast<float*>(td_smem + kTdBulkStages * L.stage_bytes + (it & 1) * L.q_bytes);
#pragma unroll
      for (int k = 0; k < kTdBulkChunk / kTdBulkThreads; ++k) {
        float* o = so + (k * kTdBulkThreads + threadIdx.x) * A;
#pragma unroll
        for (int a = 0; a < A; ++a) o[a] = (a == act[k]) ? g[k] : 0.f;
      }
      fence_proxy_async();
      __syncthreads();
      if (threadIdx.x == 0) {
        bulk_store(d.dq + e0 * A, out0 + (uint32_t)(it & 1) * L.q_bytes, L.q_bytes);
        tma_store_commit();
      }
    }
  }
  if (threadIdx.x == 0) tma_store_wait<0>();
  for (int off = 16; off; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (kTdBulkThreads >> 5) ? red[threadIdx.x] : 0.f;
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (threadIdx.x == 0 && d.loss_out != nullptr) atomicAdd(d.loss_out, v * d.inv_count);
  }
}

// Bellman branch, every streamed array 16-byte aligned, chunk byte counts multiples of 16
bool td_bulk_supported(const vdqn_td_desc* d) {
  if (d->ground_truth || d->A < 1 || d->A > 8) return false;
  auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const TdBulkLayout L = td_bulk_layout(*d);
  if ((size_t)kTdBulkStages * L.stage_bytes + 2 * (size_t)L.q_bytes > (size_t)kTdBulkMaxSmem) return false;
  return al16(d->q_s) && al16(d->q_next_target) && al16(d->q_next_online) && al16(d->dq) && al16(d->rew) &&
         al16(d->term) && al16(d->valid);
}

int td_bulk_launch(const vdqn_td_desc* d, long n_chunks, cudaStream_t stream) {
  DeviceInfo* dev = device_info();
  if (dev == nullptr) return VDQN_ERR_CUDA;
  const TdBulkLayout L = td_bulk_layout(*d);
  const size_t smem = (size_t)kTdBulkStages * L.stage_bytes + 2 * (size_t)L.q_bytes;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaSuccess;
    for (auto fn : {td_epilogue_bulk_kernel<0>, td_epilogue_bulk_kernel<1>, td_epilogue_bulk_kernel<3>})
      if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kTdBulkMaxSmem);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return set_error(VDQN_ERR_CUDA, "cudaFuncSetAttribute(td_bulk): %s", cudaGetErrorString(e));
    }
    attr_smem = kTdBulkMaxSmem;
  }
  const int grid = (int)(n_chunks < dev->num_sms ? n_chunks : dev->num_sms);
  // the shipped configuration has 3 actions (configs/experiments/real_data/config.yml:5), VALUE_LEARNING 1
  auto kfn = d->A == 3 ? td_epilogue_bulk_kernel<3> : d->A == 1 ? td_epilogue_bulk_kernel<1> : td_epilogue_bulk_kernel<0>;
  launch_kernel(kfn, grid, kTdBulkThreads, smem, stream, *d, n_chunks);
  VDQN_CHECK_LAUNCH("td_epilogue_bulk");
  return VDQN_OK;
}

}  // namespace vdqn
