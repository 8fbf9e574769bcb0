/* vdqn.h -- C ABI of the B200-native Q-learning hot path (libvdqn.so).
 *
 * The reference (uiuc-robovision/video-dqn) has NO native code and no FFI: every FLOP of
 * its training step is dispatched by PyTorch to cuDNN/cuBLAS/ATen.  The entry points below
 * are therefore the boundary a maintainer would bind in place of those ATen calls; each one
 * names the reference call site (file:line, relative to the reference root) whose work it
 * replaces.  All functions:
 *   - take plain pointers / sizes only (device pointers unless stated), no torch types;
 *   - enqueue work on the caller's stream (`stream` is a cudaStream_t passed as void*),
 *     never synchronise, never allocate device memory;
 *   - return VDQN_OK (0) or a negative error code; `vdqn_last_error()` returns a
 *     thread-local message.  There is no CPU fallback.
 * Activations are NHWC bf16; accumulation is fp32.
 */
#ifndef VDQN_H_
#define VDQN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VDQN_OK 0
#define VDQN_ERR_ARG (-1)
#define VDQN_ERR_SHAPE (-2)
#define VDQN_ERR_CUDA (-3)
#define VDQN_ERR_DRIVER (-4)

#define VDQN_ABI_VERSION 1

/* conv epilogue flags */
#define VDQN_EPI_RELU 1     /* out = max(v, 0) */
#define VDQN_EPI_OUT_F32 2  /* store fp32 instead of bf16 */
#define VDQN_EPI_SCATTER_INPUTS 4 /* with out_scatter == 2: residual / mask_src are read at the scattered pixel too */
/* debugging only (bottleneck probing of the halo kernel, tools/gpu_diag.py): results are wrong */
#define VDQN_EPI_DEBUG_NO_STORE 8
#define VDQN_EPI_DEBUG_NO_MMA 16
#define VDQN_EPI_DEBUG_NO_MATH 32

const char* vdqn_last_error(void);
int vdqn_abi_version(void);
/* Resolve driver entry points (tensor-map encoders), query the device.  Idempotent. */
int vdqn_init(int device);
int vdqn_num_sms(void);
/* Number of kernels this library has launched (or captured into a graph) so far in this process. */
long long vdqn_launch_count(void);
/* Zero `bytes` bytes of device memory on `stream` (cudaMemsetAsync: a memset node when captured in a CUDA
 * graph, no kernel).  The fused step clears its gradient arena and loss accumulator with it, so every KERNEL of
 * a step is one of this library's. */
int vdqn_zero(void* ptr, int64_t bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Implicit-GEMM convolution, forward and data-gradient (tcgen05 + TMA im2col).
 * Replaces: F.conv2d + eval-mode batch_norm (+ residual add) (+ relu) of
 *   archs/HabitatDQNMultiAction.py:49-51 (self.features(...) = torchvision ResNet-18
 *   BasicBlock.forward, torchvision/models/resnet.py:89-105) and features[8] (:30),
 *   and, for the backward pass, the cudnn_convolution_backward_input + threshold_backward +
 *   native_batch_norm_backward(d beta) launched by loss.backward() (train_q_network.py:226).
 *
 *   v[m, co] = sum_{r,s,ci} x[n, p*stride + r*dil - pad_lo, q*stride + s*dil - pad_lo, ci]
 *                           * w[co, r, s, ci]          (m = (n,p,q) row-major)
 *   v += shift[co]; v += residual[m, co]; if RELU v = max(v,0); if mask_src: v = mask_src[m,co] > 0 ? v : 0
 *   out[opix(m), co] = v;  colsum[co] += sum_m v (rounded to the stored precision)
 * BN is folded: w = gamma*rstd*W, shift = beta - mean*gamma*rstd (see vdqn_weight_prep).
 * A data-gradient is the same computation on dY with the spatially flipped, channel-
 * transposed filter and pad_lo = R-1-pad.
 */
typedef struct vdqn_conv_desc {
  const void* x;       /* bf16 [N][H][W][Cin] */
  const void* w;       /* bf16 [Cout][R][S][Cin] */
  void* out;           /* bf16 (or fp32) [opix][ldc] */
  void* out2;          /* optional bf16 second copy, written zero-dilated (stride-2 scatter) */
  const float* shift;  /* [Cout] or NULL */
  const void* residual; /* bf16 [M][ldr] or NULL */
  const void* mask_src; /* bf16 [M][ldm] or NULL */
  float* colsum;       /* [Cout], accumulated with atomics, or NULL */
  int32_t N, H, W, Cin, Cout, R, S;
  int32_t stride, dil, pad_lo, pad_hi;
  int32_t ldc, ldr, ldm, out2_ld;
  int32_t out_scatter; /* 1: opix = m; 2: opix = (n*2Ho + 2p)*2Wo + 2q */
  int32_t flags;       /* VDQN_EPI_* */
  int32_t tile_n;      /* 0 = auto (64/128/256) */
  int32_t max_ctas;    /* 0 = one per SM */
  int32_t algo;        /* 0 = auto, 1 = im2col-TMA kernel, 2 = halo-tile kernel (Cout = 64 layers),
                        * 3 = im2col-TMA kernel on CTA pairs (tcgen05 cta_group::2, 256 x tile_n tiles;
                        *     needs Cout >= 128 and an even number of 128-pixel tiles), 4 = never pairs */
  int32_t pad_hi_w;    /* upper padding along W when it differs from pad_hi (H); -1 = same */
  int32_t scatter_off_h, scatter_off_w; /* out_scatter == 2: opix = (n*2Ho + 2p + off_h)*2Wo + 2q + off_w */
  /* Two networks in one launch (online + target forward): images [split_n, N) use w2 / shift2, the
   * SMs are partitioned between the two image ranges in proportion to their tiles.  split_n == 0:
   * off.  Requires split_n*Ho*Wo % 128 == 0 for the im2col kernel. */
  const void* w2;
  const float* shift2;
  int32_t split_n;
  /* Input aliasing (packed-stem halo kernel only): images n >= x_alias_from are read from image
   * n - x_alias_shift of x.  The fused step runs online [s ; s'] and target [s'] as one 3B pass; the
   * third range re-reads the packed s' frames instead of keeping a copy.  0 / 0: off. */
  int32_t x_alias_from, x_alias_shift;
  /* Packed stem fused with torchvision's resnet.maxpool (max_pool2d(3, 2, 1)): when pool_out != NULL the
   * conv + shift + ReLU result is NOT stored (`out` is ignored); pool_out receives the pooled tensor
   * bf16 [N][H/2][W/2][Cout] and, for images n < pool_idx_images, pool_idx (uint8, same shape, may be NULL)
   * the arg-max slot r*3+s of every window (first maximum in scan order, as torch picks it) that
   * vdqn_maxpool_bwd consumes.  Needs H %% 14 == 0, W %% 8 == 0, flags = VDQN_EPI_RELU, no residual / mask. */
  void* pool_out;
  uint8_t* pool_idx;
  int32_t pool_idx_images;
  /* Second operand accumulated into the same output tile (im2col kernel, 64-channel blocks): after the
   * R*S*Cin reduction over x, Cin2 more channels are reduced over x2 (bf16 [N][H2][W2][Cin2], read through a
   * 1x1 window at stride2: pixel (p*stride2, q*stride2) feeds output pixel (p, q)); each row of `w` (and w2)
   * is [R*S*Cin | Cin2] long.  Forward: the 1x1/2 downsample branch of a residual block inside its conv2
   * (torchvision BasicBlock.forward: out = bn2(conv2(..)) + downsample(x)); backward: the downsample's data
   * gradient inside conv1's.  NULL: off. */
  const void* x2;
  int32_t Cin2, H2, W2, stride2;
  /* out_scatter == 3: stride-2 data gradient as ONE dense GEMM.  The Cout = 4 Cq columns are ordered
   * (a, b, c): row (n, p, q), column (a, b, c) is dX[n][2p + a][2q + b][c] of `out` [N][2Ho][2Wo][Cq] (the
   * filter holds the taps every output-parity class uses, zeros elsewhere); residual / mask_src / colsum are
   * indexed like the output (VDQN_EPI_SCATTER_INPUTS), colsum has Cq entries.  Needs tile_n = 128. */
} vdqn_conv_desc;
int vdqn_conv_gemm(const vdqn_conv_desc* d, void* stream);

/* ---------------------------------------------------------------------------------------
 * Convolution weight gradient (tcgen05, both operands MN-major, split over pixels).
 * Replaces: cudnn_convolution_backward_weight launched by loss.backward()
 *   (train_q_network.py:226) for every Conv2d of the Q-network.
 *   part[split][co][(r,s,ci)] = sum_{m in split} dy[m, co] * x[n, p*stride + r*dil - pad_lo, ..., ci]
 */
typedef struct vdqn_wgrad_desc {
  const void* x;   /* bf16 [N][H][W][Cin] */
  const void* dy;  /* bf16 [M][ldy], M = N*Ho*Wo */
  float* part;     /* fp32 [splits][Cout][R*S*Cin] */
  int32_t N, H, W, Cin, Cout, R, S;
  int32_t stride, dil, pad_lo, pad_hi;
  int32_t ldy;
  int32_t splits;  /* number of pixel-range partitions (>=1); algo 2: number of CTAs (<= SMs, <= tiles) */
  int32_t max_ctas;
  int32_t algo;    /* 0/1 = im2col-TMA kernel; 2 = halo-tile kernel (64->64 3x3 stride 1 only) */
} vdqn_wgrad_desc;
int vdqn_conv_wgrad(const vdqn_wgrad_desc* d, void* stream);

/* Reduce the split partials and turn them into the reference's parameter gradients:
 *   g = sum_split part;  dW[co,ci,r,s] = scale[co] * g[co,(r,s,ci)]   (torch OIHW layout, fp32)
 *   dgamma[co] += rstd[co] * (sum_k W[co,k] * g[co,k] - mean[co] * dbeta[co])   (atomics: zero it first)
 * (native_batch_norm_backward in eval mode)  `kmap` selects how the
 * GEMM-K index maps to (ci,r,s): 0 = (r,s,ci) with ci < Cin; 1 = space-to-depth stem
 * (4x4 taps x 16 packed channels -> 7x7 x 3). */
typedef struct vdqn_wgrad_fin_desc {
  const float* part;   /* [splits][Cout][K] */
  const float* w;      /* fp32 master weights, OIHW */
  const float* gamma;  /* BN weight or NULL (no BN: scale = 1) */
  const float* var;    /* running_var */
  const float* mean;   /* running_mean */
  const float* dbeta;  /* [Cout] column sums of dy, or NULL */
  float* dw;           /* fp32 OIHW */
  float* dgamma;       /* [Cout] or NULL */
  int32_t splits, Cout, Cin, R, S, K, kmap;
  float eps;
} vdqn_wgrad_fin_desc;
int vdqn_wgrad_finalize(const vdqn_wgrad_fin_desc* d, void* stream);

/* Many reductions in one launch.  `items` is a DEVICE array of n entries; build each entry on the host
 * with vdqn_wgrad_finalize_plan (returns the number of thread blocks the tensor needs, 0 if it has to
 * go through vdqn_wgrad_finalize instead -- the packed stem), set first_block to the running sum of
 * those counts, copy the array to the device once and reuse it (pointers are static across steps). */
typedef struct vdqn_wgrad_fin_item {
  vdqn_wgrad_fin_desc d;
  int32_t cn, items, TX, TY, nchunks, first_block;
} vdqn_wgrad_fin_item;
int vdqn_wgrad_finalize_plan(const vdqn_wgrad_fin_desc* d, vdqn_wgrad_fin_item* out);
int vdqn_wgrad_finalize_multi(const vdqn_wgrad_fin_item* items_dev, int32_t n, int32_t total_blocks, void* stream);

/* ---------------------------------------------------------------------------------------
 * fp32 master weights -> bf16 GEMM operands with BN folded in.
 * Replaces nothing in the reference (it keeps fp32 weights and runs BN as a separate op,
 * torch native_batch_norm(training=False)); it is the precision/layout boundary of this build.
 *   w_fwd[co][r][s][ci]   = bf16(scale[co] * W[co][ci][r][s])
 *   w_dgrad[ci][R-1-r][S-1-s][co] = same value (flipped + transposed, for the data gradient)
 *   shift[co] = beta[co] - mean[co]*scale[co]   (or the conv bias when gamma == NULL)
 * kmap as in vdqn_wgrad_finalize (1 = stem space-to-depth packing; w_dgrad unused).
 */
typedef struct vdqn_wprep_desc {
  const float* w;     /* OIHW fp32 */
  const float* gamma; const float* beta; const float* mean; const float* var; /* NULL if no BN */
  const float* bias;  /* conv bias or NULL */
  void* w_fwd;        /* bf16 [Cout][K] */
  void* w_dgrad;      /* bf16 [Cin][R*S*Cout] or NULL; dgrad_parity: four parity-class filters (see below) */
  float* shift;       /* [Cout] */
  int32_t Cout, Cin, R, S, K, kmap;
  float eps;
  /* 1 (3x3 stride-2 convs): w_dgrad holds the data-gradient filters of the four output-parity
   * classes (h%2, w%2) back to back -- class (a,b) is [Cin][na][nb][Cout] with na = 1 + a,
   * nb = 1 + b taps (rows r = 1 | {2, 0}, same for columns), class offsets 0, 1, 3, 5 taps x Cin*Cout. */
  int32_t dgrad_parity;
  /* Concatenated layouts (0 = the plain ones above).  ldw_fwd / fwd_col0: row pitch and first column of
   * this tensor inside a wider forward matrix (the 1x1 downsample appended to conv2's rows); gamma_b ..
   * var_b: a second BatchNorm whose shift beta_b - mean_b * gamma_b / sqrt(var_b + eps) is ADDED to this
   * tensor's (conv2 carries the downsample's).  ldw_dgrad / dgrad_col0: the same for the data-gradient
   * matrix.  dgrad_parity == 2: the data-gradient filter of a 3x3 stride-2 conv as one [4 Cin][2][2][Cout]
   * matrix (row (a*2 + b)*Cin + ci, column (u*2 + v)*Cout + co; taps a class does not use stay zero: the
   * buffer must be zero-initialised once). */
  int32_t ldw_fwd, fwd_col0, ldw_dgrad, dgrad_col0;
  const float* gamma_b; const float* beta_b; const float* mean_b; const float* var_b;
} vdqn_wprep_desc;
int vdqn_weight_prep(const vdqn_wprep_desc* d, void* stream);
/* Same for `n` tensors in one launch: `descs_dev` / `offsets_dev` are DEVICE arrays (offsets[t] = sum of
 * Cout*K of tensors before t; total = sum over all). */
int vdqn_weight_prep_multi(const vdqn_wprep_desc* descs_dev, const int64_t* offsets_dev, int32_t n,
                           int64_t total, void* stream);
/* Tiled variant (coalesced reads and writes through a shared-memory transpose) for tensors with
 * kmap == 0, Cin % 32 == 0, Cout % 32 == 0 and R*S <= 9: one block per 32x32 channel tile;
 * tile_offsets_dev[t] = first block of tensor t (device array), total_tiles = number of blocks. */
int vdqn_weight_prep_tiled(const vdqn_wprep_desc* descs_dev, const int32_t* tile_offsets_dev, int32_t n,
                           int32_t total_tiles, void* stream);

/* ---------------------------------------------------------------------------------------
 * Input staging: NCHW fp32 (dataloaders/q_learning_real.py:75-76 output) or uint8 HWC
 * (util/torch.py:26-36 `to_imgnet` fused) -> space-to-depth NHWC bf16 [N][H/2][W/2][16]
 * with channel (ph*2+pw)*3 + c, channels 12..15 zero.  */
int vdqn_stem_pack_f32(const float* x_nchw, void* out, int32_t N, int32_t H, int32_t W, void* stream);
int vdqn_stem_pack_u8(const uint8_t* x_nhwc, void* out, int32_t N, int32_t H, int32_t W, void* stream);

/* max_pool2d(3, 2, 1) on NHWC bf16 (torchvision resnet.maxpool); idx (uint8 window slot 0..8)
 * may be NULL for inference.  Backward scatters dy to the arg-max and applies the stem ReLU mask,
 * taken from the pooled output y (window max > 0); dx is [N][H][W][C]. */
int vdqn_maxpool_fwd(const void* x, void* y, uint8_t* idx, int32_t N, int32_t H, int32_t W, int32_t C, void* stream);
int vdqn_maxpool_bwd(const void* dy, const uint8_t* idx, const void* y, void* dx, float* colsum,
                     int32_t N, int32_t H, int32_t W, int32_t C, void* stream);

/* ---------------------------------------------------------------------------------------
 * Q-head MLP in fp32 (archs/HabitatDQNMultiAction.py:31,53: Linear 1600-512-256-15 + ReLU).
 *   y[b,o] = act(sum_k x[b,k] * w[o,k] + bias[o])
 *   bwd:  dx[b,k] = sum_o dy[b,o] w[o,k];  dw[o,k] = sum_b dy[b,o] x[b,k];  db[o] = sum_b dy[b,o]
 *   (dy is pre-masked by the caller's activation: vdqn_linear_bwd applies `y > 0` when relu != 0) */
int vdqn_linear_fwd(const float* x, const float* w, const float* bias, float* y,
                    int32_t B, int32_t K, int32_t O, int32_t relu, void* stream);
int vdqn_linear_bwd(const float* x, const float* w, const float* y, float* dy /* overwritten: masked */,
                    float* dx, float* dw, float* db,
                    int32_t B, int32_t K, int32_t O, int32_t relu, void* stream);
/* First half of the backward of y = relu(x W^T + b) when the weight / data gradients run on the
 * tensor-core conv kernels (top.0 seen as a 5x5 valid convolution over the head output,
 * archs/HabitatDQNMultiAction.py:31): dy <- dy * (y > 0) in place, db[o] = sum_b dy[b,o], and an
 * optional bf16 copy of the masked gradient (dy_bf16, [B][O]). */
int vdqn_relu_mask_colsum(float* dy, const float* y, void* dy_bf16, float* db, int32_t B, int32_t O,
                          int32_t relu, void* stream);

/* head conv output bf16 [B][P][C] (NHWC) <-> fp32 [B][C*P] (torch Flatten order); the backward
 * also applies the head ReLU mask and accumulates the conv-bias gradient. */
int vdqn_head_flatten_fwd(const void* h_nhwc, float* flat, int32_t B, int32_t P, int32_t C, void* stream);
/* out[n][c] = mean over the P pixels of x[n][p][c] (NHWC bf16 -> fp32): the AdaptiveAvgPool2d(1) that
 * ends the trunk of the `basic` architecture (archs/HabitatDQNMultiAction.py:33, extra_capacity=False). */
int vdqn_avgpool_fwd(const void* x_nhwc, float* out, int32_t N, int32_t P, int32_t C, void* stream);

int vdqn_head_flatten_bwd(const float* dflat, const void* h_nhwc, void* dh_nhwc, float* dbias,
                          int32_t B, int32_t P, int32_t C, void* stream);

/* ---------------------------------------------------------------------------------------
 * Q-head MLP on the tensor cores at fp32-grade accuracy (the fast path of the fused step; replaces the
 * cuBLAS sgemm calls PyTorch dispatches for top.0 / top.2 / top.4 and their backward,
 * archs/HabitatDQNMultiAction.py:31,53).  Every fp32 operand is held as TWO bf16 matrices hi = bf16(x),
 * lo = bf16(x - hi); a GEMM accumulates up to three operand pairs ("segments": hi*hi, lo*hi, hi*lo) into
 * one fp32 tile:
 *     D[m,n] = sum_s sum_k A_s[m,k] * B_s[n,k]
 * Operands are row-major bf16 matrices with 16-byte aligned rows, read either K-major (matrix [M or N][K])
 * or, with a_mn / b_mn != 0, MN-major (matrix [K][M or N]) -- the same buffers serve x W^T, dy W and dy^T x.
 * Epilogue, in this order: + bias[n]; ReLU; zero where mask_f32 / mask_bf16 [m][n] <= 0 (ReLU mask of a stored
 * activation); then any of: fp32 store (optionally un-permuting columns: n = f*c*p + pp*c + cc goes to
 * f*c*p + cc*p + pp, for d top.0.weight), hi/lo bf16 store (columns up to ld_hl, zeros beyond N), bf16
 * store, per-column sums accumulated with atomics into colsum[n % colsum_mod] (bias gradients).
 * split_m > 0 (multiple of 128): rows >= split_m use the second operand set b2 / bias2 (target network). */
typedef struct vdqn_mlp_operand {
  const void* ptr;          /* bf16 */
  int32_t rows, cols, ld;   /* the matrix as stored; ld in elements, multiple of 8 */
} vdqn_mlp_operand;
typedef struct vdqn_mlp_gemm_desc {
  vdqn_mlp_operand a[3], b[3], b2[3];
  int32_t K[3];
  int32_t nseg, M, N, BN /* 64, 128 or 256 */, a_mn, b_mn, split_m;
  const float* bias; const float* bias2;
  int32_t relu;
  const float* mask_f32; const void* mask_bf16; int32_t ldmask;
  float* out_f32; int32_t ld_f32, perm_c, perm_p;
  void* out_hi; void* out_lo; int32_t ld_hl;
  void* out_bf16; int32_t ld_bf16;
  float* colsum; int32_t colsum_mod;
} vdqn_mlp_gemm_desc;
int vdqn_mlp_gemm(const vdqn_mlp_gemm_desc* d, void* stream);
/* up to 4 independent GEMMs in ONE launch (each is latency-bound: the three weight gradients and the head
 * conv's dy of the backward pass share a launch); only descs[0] may use split_m */
int vdqn_mlp_gemm_grouped(const vdqn_mlp_gemm_desc* descs, int32_t n, void* stream);
/* x fp32 [rows][cols] -> hi / lo bf16 [rows][ld_out] (columns >= cols zero); perm_c != 0 re-orders the
 * columns of every (perm_c x perm_p) block from c*perm_p + p to p*perm_c + c (top.0.weight -> the NHWC order
 * of the head-conv output); colsum (optional) accumulates the per-column sums of x. */
int vdqn_split_bf16(const float* x, void* hi, void* lo, int32_t rows, int32_t cols, int32_t ld_out,
                    int32_t perm_c, int32_t perm_p, float* colsum, void* stream);

/* ---------------------------------------------------------------------------------------
 * Data-parallel gradient exchange as one kernel over NVLink / NVSwitch peer memory (the reference is
 * single-GPU, train_q_network.py:275).  The flat fp32 gradient arena of every rank lives in
 * symmetric memory: peer_bufs[r] / peer_flags[r] are THIS process' device pointers to rank r's arena and to
 * rank r's flag block (uint32[8 * world], zero-initialised once); multicast_ptr is the NVSwitch multicast
 * address of the arena or NULL (then plain peer loads / stores are used).  After the call, enqueued on every
 * rank's stream, every arena holds the element-wise SUM over ranks, bit-identical on all ranks (each element is
 * summed by one rank and copied).  epoch / counter: two zero-initialised uint32 in LOCAL device memory that the
 * kernel maintains (CUDA-graph replay safe).  n (elements) must be a multiple of 4 * world. */
typedef struct vdqn_nvl_desc {
  void* const* peer_bufs;
  void* const* peer_flags;
  void* multicast_ptr;
  uint32_t* epoch;
  uint32_t* counter;
  int64_t n;                /* elements exchanged, starting at element `first` of the arena (multiple of 4) */
  int32_t rank, world, max_ctas;
  int64_t first;
  /* Several exchanges may be in flight (an early one next to the backward pass, the rest after it): each uses
   * its own channel 0..3 = its own flag slots [channel * 2 world, ...) of the flag block (uint32[8 * world])
   * and its own epoch / counter pair.  threads: 128 / 256 / 512 per CTA (0 = 512): 128-thread CTAs fit on an
   * SM beside a persistent conv CTA. */
  int32_t channel, threads;
} vdqn_nvl_desc;
int vdqn_nvl_allreduce(const vdqn_nvl_desc* d, void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused TD epilogue: replaces the ~24 ATen launches of process_batch
 * (train_q_network.py:134-180): repeat/gather/argmax/gather/detach/mul/add/clamp/sub/pow/mean
 * and their backward.  q_* are fp32 [B][C][A]; act int64 [B]; rew/term/valid int64 [B][C].
 *   a* = argmax_a q_sel[b,c,:] (first max), q_sel = q_next_online if double_dqn else q_next_target
 *   y  = rew + gamma * q_next_target[b,c,a*] * (1 - term)   (linear: rew + (q - 0.1))
 *   y  = clamp(y, 0, 1) if clip_rect;  l = 0.5 (q_s[b,c,act[b]] - y)^2 * (valid if use_valid)
 *   loss_sum += sum l   (caller divides by B*C*world or passes inv_count);  dq[b,c,a] = (q_b - y)*mask*inv_count on a == act[b]
 * loss_out[0] accumulates sum(l)*inv_count with atomics: zero it first. best/y_out optional. */
typedef struct vdqn_td_desc {
  const float* q_s; const float* q_next_online; const float* q_next_target;
  const int64_t* act; const int64_t* rew; const int64_t* term; const int64_t* valid;
  float* dq; float* loss_out; int64_t* best_out; float* y_out;
  int32_t B, C, A;
  float gamma; float inv_count;
  int32_t double_dqn, clip_rect, linear, use_valid;
  /* TRAIN_ON_GROUND_TRUTH (process_batch(compare_ground_truth=True), train_q_network.py:170-178):
   * ground_truth != 0 regresses q_s[b,c,act[b]] onto gt[b,c] (float64, gamma^steps_to_reward from
   * dataloaders/q_learning_real.py:86-89) instead of the Bellman target; q_next_*, rew, term are not
   * read.  value_learning != 0 masks the NaN entries: l = 0.5 (q_b*mask - gt0)^2, mask = !isnan(gt);
   * otherwise l = 0.5 (q_b - gt)^2 and NaNs propagate as in the reference. */
  const double* gt;
  int32_t ground_truth, value_learning;
  /* CONFIDENCE_REWARD (train_q_network.py:101, dataloaders/q_learning_real.py:76-77): the loader then
   * yields the detector scores (floating point) as reward and terminal, which process_batch casts with
   * `.float()` (:158-160).  labels_f32 != 0: rew / term / valid point to float32 arrays (the cast done). */
  int32_t labels_f32;
} vdqn_td_desc;
int vdqn_td_epilogue(const vdqn_td_desc* d, void* stream);

/* value[r] = max_a q[r][a] and (optional) its first arg-max, r over B*C rows: the
 * `model(images).max(2).values` of visualize_value.py:96-97 and `[0, class, :].max()` of
 * evaluation/evaluate.py:110-114. */
int vdqn_q_max(const float* q, float* value, int64_t* arg, int64_t rows, int32_t A, void* stream);

/* ---------------------------------------------------------------------------------------
 * Training step of the inverse-dynamics model (train_inverse_model.py:85-110): the pieces the
 * Q-learning path does not already have.
 *
 * Softmax cross-entropy with mean reduction, `nn.CrossEntropyLoss()(y, act)` (:100-102) and the
 * accuracy count `y.argmax(1) == act` (:105-106), fused with the gradient:
 *   loss_out[0] += inv_count * sum_b (logsumexp(y[b,:]) - y[b,act[b]])     (zero it before the call)
 *   dlogits[b,c] = inv_count * (softmax(y[b,:])[c] - [c == act[b]])        (may be NULL: validation)
 *   correct_out[0] += #{b : first arg-max of y[b,:] == act[b]}             (may be NULL)
 * logits fp32 [B,C] row-major, labels int64 [B], 1 <= C <= 32. */
int vdqn_cross_entropy(const float* logits, const int64_t* labels, float* dlogits, float* loss_out,
                       int32_t* correct_out, int32_t B, int32_t C, float inv_count, void* stream);
/* Element dropout of `nn.Dropout2d(0.5)` applied to the [B,128] fc1 output (:44,78; on a 2-D input it
 * drops single elements).  keep[i] in {0,1}:
 *   vdqn_dropout_mask : keep[i] = u(seed, counter, i) >= p, u uniform in [0,1) from a counter-based
 *                       generator (splitmix64 of seed, counter, i) -- stateless, graph-replay safe
 *   vdqn_dropout_apply: y[i] = keep[i] ? x[i] * scale : 0   (scale = 1/(1-p); y may alias x; the same
 *                       call is the backward pass on the gradient) */
int vdqn_dropout_mask(uint8_t* keep, int64_t n, float p, uint64_t seed, uint64_t counter, void* stream);
int vdqn_dropout_apply(const float* x, const uint8_t* keep, float scale, float* y, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------
 * Train-mode BatchNorm2d for the `basic` architecture (ARCHITECTURE != 'extra_capacity'): `set_train()`
 * leaves its trunk BatchNorms in train mode (archs/HabitatDQNMultiAction.py:37-40), so `model(before)`
 * and `model(after)` (train_q_network.py:131,142) normalise with BATCH statistics and update the running
 * ones (torch.nn.BatchNorm2d: momentum 0.1, eps 1e-5, unbiased running variance).  Activations are
 * NHWC bf16 viewed as [M = N*H*W, C], C % 8 == 0; statistics fp32, their sums fp64.
 *
 *   vdqn_bn_stats       sums[0:C] = sum_m x[m,c],  sums[C:2C] = sum_m x[m,c]^2   (the call zeroes sums)
 *   vdqn_bn_finalize    mean = sums/M, var = sumsq/M - mean^2 (biased), rstd = 1/sqrt(var + eps);
 *                       scale = gamma*rstd, shift = beta - mean*scale;
 *                       if running_mean != NULL: running_mean = (1-mom)*running_mean + mom*mean,
 *                       running_var = (1-mom)*running_var + mom*var*M/(M-1); if nbt != NULL: ++*nbt
 *   vdqn_bn_apply       y = x*scale[c] + shift[c] (+ residual) (ReLU if relu), bf16
 *   vdqn_bn_bwd_reduce  sums[0:C] = sum_m dy, sums[C:2C] = sum_m dy * xhat,  xhat = (x - mean)*rstd
 *   vdqn_bn_bwd_apply   dgamma = sums[C:2C], dbeta = sums[0:C] (fp32, overwritten);
 *                       dx = gamma*rstd * (dy - dbeta/M - xhat*dgamma/M), bf16
 * dy is the gradient w.r.t. the BatchNorm OUTPUT (already masked by the following ReLU).
 *   vdqn_avgpool_bwd    dfeat[n,p,c] = feat[n,p,c] > 0 ? dpooled[n,c] / P : 0  (AdaptiveAvgPool2d(1) after the
 *                       last block's ReLU, archs/HabitatDQNMultiAction.py:33) */
int vdqn_bn_stats(const void* x, double* sums, int64_t M, int32_t C, void* stream);
int vdqn_bn_finalize(const double* sums, int64_t M, int32_t C, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                     float eps, float* mean, float* rstd, float* scale, float* shift, void* stream);
int vdqn_bn_apply(const void* x, const float* scale, const float* shift, const void* residual, int32_t relu,
                  void* y, int64_t M, int32_t C, void* stream);
int vdqn_bn_bwd_reduce(const void* dy, const void* x, const float* mean, const float* rstd, double* sums,
                       int64_t M, int32_t C, void* stream);
int vdqn_bn_bwd_apply(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma,
                      const double* sums, float* dgamma, float* dbeta, void* dx, int64_t M, int32_t C,
                      void* stream);
int vdqn_avgpool_bwd(const float* dpooled, const void* feat, void* dfeat, int32_t N, int32_t P, int32_t C,
                     void* stream);

/* ---------------------------------------------------------------------------------------
 * Fused multi-tensor Adam (+ optional hard target-network sync in the same pass).
 * Replaces optim.Adam.step (train_q_network.py:124,227; ~550 launches under torch 1.3.1)
 * and target_net.load_state_dict(model.state_dict()) (:215-216).
 * p, g, m, v are flat fp32 arenas of n elements; grad_scale multiplies g (1/world for DDP).
 *   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 * if target != NULL: target[i] = p_new[i]. */
int vdqn_adam_fused(float* p, const float* g, float* m, float* v, float* target, int64_t n,
                    double lr, double beta1, double beta2, double eps, int32_t step, float grad_scale,
                    void* stream);
/* Same update for CUDA-graph replay: the step counter lives in HBM (`step_dev`, incremented by the
 * call) and the bias-correction scalars are recomputed on the device into `scalars_dev[2]`. */
int vdqn_adam_fused_graph(float* p, const float* g, float* m, float* v, float* target, int64_t n,
                          double lr, double beta1, double beta2, double eps, float grad_scale,
                          int32_t* step_dev, float* scalars_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VDQN_H_ */
