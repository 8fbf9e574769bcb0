"""Tensor-level wrappers over the C ABI (include/vdqn.h).  Each takes CUDA torch tensors, checks
shapes/dtypes (ValueError before launch, mirroring the reference's `Exception("bad shape")`
style of failing early) and enqueues on torch's current stream.  No fallback paths."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

bf16 = torch.bfloat16

# When set to a list, conv_gemm / conv_wgrad / adam_fused / td_epilogue bracket their launch with
# CUDA events on the launching stream and append (kind, tag, start_event, end_event).  bench.py uses
# this for the per-kernel roofline; it is None (zero overhead) otherwise.
PROFILE = None


class _Prof:
    def __init__(self, kind, tag):
        self.kind, self.tag = kind, tag

    def __enter__(self):
        if PROFILE is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e.record()
            PROFILE.append((self.kind, self.tag, self.s, self.e))
        return False


def _req(cond, msg):
    if not cond:
        raise ValueError(msg)


def _cuda(t, dtype, name):
    _req(t.is_cuda, f"{name} must be a CUDA tensor")
    _req(t.dtype == dtype, f"{name} must be {dtype}, got {t.dtype}")
    _req(t.is_contiguous(), f"{name} must be contiguous")
    return t


_NUM_SMS = None


def num_sms() -> int:
    global _NUM_SMS
    if _NUM_SMS is None:
        n = L.load().vdqn_num_sms()
        if n <= 0:
            L.check(-3, "num_sms")
        _NUM_SMS = n
    return _NUM_SMS


def zero_(t):
    """t <- 0 through cudaMemsetAsync on the current stream (no kernel launch)."""
    _req(t.is_cuda and t.is_contiguous(), "zero_: contiguous CUDA tensor")
    L.check(L.load().vdqn_zero(t.data_ptr(), t.numel() * t.element_size(), L.stream_ptr()), "zero")
    return t


def conv_out_hw(H, W, R, S, stride, pad_lo, pad_hi, dil=1):
    return ((H + pad_lo + pad_hi - (R - 1) * dil - 1) // stride + 1,
            (W + pad_lo + pad_hi - (S - 1) * dil - 1) // stride + 1)


def conv_gemm(x, w, stride=1, pad_lo=0, pad_hi=None, *, shift=None, residual=None, mask_src=None,
              relu=False, out_f32=False, colsum=None, out=None, out2=None, out_scatter=1,
              tile_n=0, max_ctas=0, dil=1, algo=0, pad_hi_w=-1, scatter_off=(0, 0), scatter_inputs=False,
              w2=None, shift2=None, split_n=0, x_alias=None, pool_out=None, pool_idx=None, pool_idx_images=0,
              debug_flags=0, x2=None, stride2=1, ksize=None):
    """x [N,H,W,Cin] bf16, w [Cout,R,S,Cin] bf16 -> out [N,Ho,Wo,Cout] (or zero-dilated
    [N,2Ho,2Wo,Cout] when out_scatter == 2; `out` must then be pre-zeroed).
    x_alias = (first, shift): images n >= first are read from image n - shift (packed stem only)."""
    lib = L.load()
    _cuda(x, bf16, "x"); _cuda(w, bf16, "w")
    N, H, W_, Cin = x.shape
    if ksize is not None:
        # `w` as a matrix [Cout, R*S*Cin (+ Cin2)]: the rows may carry the columns of a second operand x2 (a
        # 1x1 / stride2 window over another tensor, accumulated into the same output tile)
        R, S = ksize
        Cout = w.shape[0]
        _req(x.dim() == 4 and w.dim() == 2 and w.shape[1] == R * S * Cin + (x2.shape[3] if x2 is not None else 0),
             "bad shape")
    else:
        _req(x.dim() == 4 and w.dim() == 4 and x.shape[3] == w.shape[3] and x2 is None, "bad shape")
        Cout, R, S, _ = w.shape
    pad_hi = pad_lo if pad_hi is None else pad_hi
    Ho = conv_out_hw(H, W_, R, S, stride, pad_lo, pad_hi, dil)[0]
    Wo = conv_out_hw(H, W_, R, S, stride, pad_lo, pad_hi if pad_hi_w < 0 else pad_hi_w, dil)[1]
    odt = torch.float32 if out_f32 else bf16
    if pool_out is not None:
        # packed stem + max_pool2d(3, 2, 1) in one kernel: only the pooled tensor (and arg-max slots) is stored
        _cuda(pool_out, bf16, "pool_out")
        _req(pool_out.numel() == N * (Ho // 2) * (Wo // 2) * Cout, "bad shape")
        if pool_idx is not None:
            _cuda(pool_idx, torch.uint8, "pool_idx")
            _req(pool_idx.numel() >= pool_idx_images * (Ho // 2) * (Wo // 2) * Cout, "bad shape")
        out = pool_out                      # placeholder: `out` is not written
    if out is None:
        if out_scatter == 2:
            out = torch.zeros(N, 2 * Ho, 2 * Wo, Cout, device=x.device, dtype=odt)
        else:
            out = torch.empty(N, Ho, Wo, Cout, device=x.device, dtype=odt)
    _cuda(out, odt, "out")
    d = L.ConvDesc()
    d.x, d.w, d.out = x.data_ptr(), w.data_ptr(), out.data_ptr()
    d.out2 = L.ptr(out2)
    d.shift = L.ptr(shift); d.residual = L.ptr(residual); d.mask_src = L.ptr(mask_src)
    d.colsum = L.ptr(colsum)
    if shift is not None:
        _cuda(shift, torch.float32, "shift"); _req(shift.numel() == Cout, "bad shape")
    n_in = N * Ho * Wo * Cout * (4 if (scatter_inputs and out_scatter == 2) else 1)
    if residual is not None:
        _cuda(residual, bf16, "residual"); _req(residual.numel() == n_in, "bad shape")
    if mask_src is not None:
        _cuda(mask_src, bf16, "mask_src"); _req(mask_src.numel() == n_in, "bad shape")
    if colsum is not None:
        _cuda(colsum, torch.float32, "colsum")
        _req(colsum.numel() == (Cout // 4 if out_scatter == 3 else Cout), "bad shape")
    if out2 is not None:
        _cuda(out2, bf16, "out2"); _req(out2.numel() == N * 4 * Ho * Wo * Cout, "bad shape")
    d.N, d.H, d.W, d.Cin, d.Cout, d.R, d.S = N, H, W_, Cin, Cout, R, S
    d.stride, d.dil, d.pad_lo, d.pad_hi = stride, dil, pad_lo, pad_hi
    d.ldc = d.ldr = d.ldm = d.out2_ld = Cout
    d.out_scatter = out_scatter
    d.flags = (L.EPI_RELU if relu else 0) | (L.EPI_OUT_F32 if out_f32 else 0) | \
        (L.EPI_SCATTER_INPUTS if scatter_inputs else 0) | int(debug_flags)
    d.pad_hi_w, d.scatter_off_h, d.scatter_off_w = pad_hi_w, scatter_off[0], scatter_off[1]
    if split_n:
        _cuda(w2, bf16, "w2"); _req(w2.shape == w.shape and shift2 is not None, "bad shape")
        d.w2, d.shift2, d.split_n = w2.data_ptr(), shift2.data_ptr(), split_n
    d.tile_n, d.max_ctas, d.algo = tile_n, max_ctas, algo
    if x_alias is not None:
        d.x_alias_from, d.x_alias_shift = x_alias
    if pool_out is not None:
        d.pool_out, d.pool_idx, d.pool_idx_images = pool_out.data_ptr(), L.ptr(pool_idx), int(pool_idx_images)
    if x2 is not None:
        _cuda(x2, bf16, "x2"); _req(x2.dim() == 4 and x2.shape[0] == N, "bad shape")
        d.x2, d.H2, d.W2, d.Cin2, d.stride2 = x2.data_ptr(), x2.shape[1], x2.shape[2], x2.shape[3], stride2
    with _Prof("igemm", (N, H, W_, Cin, Cout, R, stride)):
        L.check(lib.vdqn_conv_gemm(C.byref(d), L.stream_ptr()), "conv_gemm")
    return out


def conv_wgrad(x, dy, R, S, stride=1, pad_lo=0, pad_hi=None, *, splits=1, part=None, max_ctas=0, dil=1,
               algo=0):
    """x [N,H,W,Cin] bf16, dy [N,Ho,Wo,Cout] bf16 -> part [splits,Cout,R*S*Cin] fp32."""
    lib = L.load()
    _cuda(x, bf16, "x"); _cuda(dy, bf16, "dy")
    N, H, W_, Cin = x.shape
    pad_hi = pad_lo if pad_hi is None else pad_hi
    Ho, Wo = conv_out_hw(H, W_, R, S, stride, pad_lo, pad_hi, dil)
    Cout = dy.shape[-1]
    _req(dy.numel() == N * Ho * Wo * Cout, "bad shape")
    K = R * S * Cin
    if part is None:
        part = torch.empty(splits, Cout, K, device=x.device, dtype=torch.float32)
    _cuda(part, torch.float32, "part"); _req(part.numel() >= splits * Cout * K, "bad shape")
    d = L.WgradDesc()
    d.x, d.dy, d.part = x.data_ptr(), dy.data_ptr(), part.data_ptr()
    d.N, d.H, d.W, d.Cin, d.Cout, d.R, d.S = N, H, W_, Cin, Cout, R, S
    d.stride, d.dil, d.pad_lo, d.pad_hi = stride, dil, pad_lo, pad_hi
    d.ldy, d.splits, d.max_ctas, d.algo = Cout, splits, max_ctas, algo
    with _Prof("wgrad", (N, H, W_, Cin, Cout, R, stride)):
        L.check(lib.vdqn_conv_wgrad(C.byref(d), L.stream_ptr()), "conv_wgrad")
    return part


def wgrad_finalize_desc(part, w, dw, *, splits, Cout, Cin, R, S, K, kmap=0, gamma=None, var=None,
                        mean=None, dbeta=None, dgamma=None, eps=1e-5):
    d = L.WgradFinDesc()
    d.part, d.w, d.dw = part.data_ptr(), w.data_ptr(), dw.data_ptr()
    d.gamma, d.var, d.mean = L.ptr(gamma), L.ptr(var), L.ptr(mean)
    d.dbeta, d.dgamma = L.ptr(dbeta), L.ptr(dgamma)
    d.splits, d.Cout, d.Cin, d.R, d.S, d.K, d.kmap = splits, Cout, Cin, R, S, K, kmap
    d.eps = eps
    return d


def wgrad_finalize_table(descs, device):
    """Device table for wgrad_finalize_multi from a list of WgradFinDesc (all must be row-form
    capable).  Returns (table uint8 tensor, n, total_blocks)."""
    lib = L.load()
    items = (L.WgradFinItem * len(descs))()
    total = 0
    for it, d in zip(items, descs):
        nb = lib.vdqn_wgrad_finalize_plan(C.byref(d), C.byref(it))
        _req(nb > 0, "wgrad_finalize_multi: tensor not reducible by the row-form kernel")
        it.first_block = total
        total += nb
    tab = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8).clone().to(device)
    return tab, len(descs), total


def wgrad_finalize_multi(table):
    lib = L.load()
    tab, n, total = table
    with _Prof("wgrad_finalize", (n, total)):
        L.check(lib.vdqn_wgrad_finalize_multi(tab.data_ptr(), n, total, L.stream_ptr()), "wgrad_finalize_multi")


def wgrad_finalize(part, w, dw, *, splits, Cout, Cin, R, S, K, kmap=0, gamma=None, var=None,
                   mean=None, dbeta=None, dgamma=None, eps=1e-5):
    lib = L.load()
    d = wgrad_finalize_desc(part, w, dw, splits=splits, Cout=Cout, Cin=Cin, R=R, S=S, K=K, kmap=kmap,
                            gamma=gamma, var=var, mean=mean, dbeta=dbeta, dgamma=dgamma, eps=eps)
    with _Prof("wgrad_finalize", (Cout, K, splits)):
        L.check(lib.vdqn_wgrad_finalize(C.byref(d), L.stream_ptr()), "wgrad_finalize")
    return dw


def weight_prep(w, w_fwd, shift, *, w_dgrad=None, gamma=None, beta=None, mean=None, var=None,
                bias=None, kmap=0, eps=1e-5, dgrad_parity=False):
    """w OIHW fp32 -> w_fwd bf16 [Cout, K] (+ w_dgrad bf16 [Cin, R*S*Cout]) and shift [Cout]."""
    lib = L.load()
    _cuda(w, torch.float32, "w")
    Cout, Cin, R, S = w.shape
    K = w_fwd.numel() // Cout
    d = L.WprepDesc()
    d.w, d.w_fwd, d.shift, d.w_dgrad = w.data_ptr(), w_fwd.data_ptr(), shift.data_ptr(), L.ptr(w_dgrad)
    d.gamma, d.beta, d.mean, d.var, d.bias = (L.ptr(gamma), L.ptr(beta), L.ptr(mean), L.ptr(var),
                                              L.ptr(bias))
    d.Cout, d.Cin, d.R, d.S, d.K, d.kmap = Cout, Cin, R, S, K, kmap
    d.eps = eps
    d.dgrad_parity = int(dgrad_parity)
    L.check(lib.vdqn_weight_prep(C.byref(d), L.stream_ptr()), "weight_prep")


def stem_pack(x, out=None):
    """NCHW fp32 [N,3,H,W] or uint8 HWC [N,H,W,3] -> space-to-depth [N,H/2,W/2,16] bf16."""
    lib = L.load()
    _req(x.is_cuda and x.is_contiguous() and x.dim() == 4, "bad shape")
    if x.dtype == torch.uint8:
        N, H, W_, c = x.shape
        fn = lib.vdqn_stem_pack_u8
    else:
        _req(x.dtype == torch.float32, "frames must be fp32 NCHW or uint8 NHWC")
        N, c, H, W_ = x.shape
        fn = lib.vdqn_stem_pack_f32
    _req(c == 3, "bad shape")
    if out is None:
        out = torch.empty(N, H // 2, W_ // 2, 16, device=x.device, dtype=bf16)
    with _Prof("stem_pack", (N,)):
        L.check(fn(x.data_ptr(), out.data_ptr(), N, H, W_, L.stream_ptr()), "stem_pack")
    return out


def maxpool_fwd(x, y=None, idx=None, save_idx=False):
    lib = L.load()
    _cuda(x, bf16, "x")
    N, H, W_, Cc = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W_ - 1) // 2 + 1
    if y is None:
        y = torch.empty(N, Ho, Wo, Cc, device=x.device, dtype=bf16)
    if save_idx and idx is None:
        idx = torch.empty(N, Ho, Wo, Cc, device=x.device, dtype=torch.uint8)
    with _Prof("maxpool_fwd", (N,)):
        L.check(lib.vdqn_maxpool_fwd(x.data_ptr(), y.data_ptr(), L.ptr(idx), N, H, W_, Cc, L.stream_ptr()),
                "maxpool_fwd")
    return y, idx


def maxpool_bwd(dy, idx, y, dx=None, colsum=None):
    """dy, idx, y: pooled-resolution [N,Ho,Wo,C]; dx: [N,2Ho,2Wo,C] (masked by the stem ReLU via y > 0)."""
    lib = L.load()
    _cuda(dy, bf16, "dy"); _cuda(y, bf16, "y"); _cuda(idx, torch.uint8, "idx")
    N, Ho, Wo, Cc = y.shape
    H, W_ = 2 * Ho, 2 * Wo
    if dx is None:
        dx = torch.empty(N, H, W_, Cc, device=y.device, dtype=bf16)
    with _Prof("maxpool_bwd", (N,)):
        L.check(lib.vdqn_maxpool_bwd(dy.data_ptr(), idx.data_ptr(), y.data_ptr(), dx.data_ptr(),
                                     L.ptr(colsum), N, H, W_, Cc, L.stream_ptr()), "maxpool_bwd")
    return dx


def linear_fwd(x, w, bias, relu, y=None):
    lib = L.load()
    _cuda(x, torch.float32, "x"); _cuda(w, torch.float32, "w")
    B, K = x.shape
    O = w.shape[0]
    _req(w.shape[1] == K, "bad shape")
    if y is None:
        y = torch.empty(B, O, device=x.device, dtype=torch.float32)
    with _Prof("mlp", (B, K, O)):
        L.check(lib.vdqn_linear_fwd(x.data_ptr(), w.data_ptr(), L.ptr(bias), y.data_ptr(), B, K, O,
                                    int(relu), L.stream_ptr()), "linear_fwd")
    return y


def linear_bwd(x, w, y, dy, dw, db, relu, dx=None, need_dx=True):
    """dy is overwritten with the ReLU-masked gradient."""
    lib = L.load()
    B, K = x.shape
    O = w.shape[0]
    if need_dx and dx is None:
        dx = torch.empty(B, K, device=x.device, dtype=torch.float32)
    with _Prof("mlp", (B, K, O)):
        L.check(lib.vdqn_linear_bwd(x.data_ptr(), w.data_ptr(), L.ptr(y), dy.data_ptr(),
                                    L.ptr(dx) if need_dx else None, dw.data_ptr(), db.data_ptr(),
                                    B, K, O, int(relu), L.stream_ptr()), "linear_bwd")
    return dx


def split_bf16(x, hi, lo, *, perm=None, colsum=None):
    """x fp32 [rows, cols] -> hi = bf16(x), lo = bf16(x - hi), both [rows, ld] with ld >= cols (zero padded);
    perm = (c, p): every block of c*p columns is re-ordered from c-major to p-major (top.0.weight)."""
    lib = L.load()
    _cuda(x, torch.float32, "x"); _cuda(hi, bf16, "hi"); _cuda(lo, bf16, "lo")
    _req(x.dim() == 2 and hi.shape == lo.shape and hi.shape[0] == x.shape[0] and hi.shape[1] >= x.shape[1], "bad shape")
    pc, pp = perm if perm is not None else (0, 0)
    with _Prof("mlp", (x.shape[0], x.shape[1])):
        L.check(lib.vdqn_split_bf16(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.shape[0], x.shape[1], hi.shape[1],
                                    pc, pp, L.ptr(colsum), L.stream_ptr()), "split_bf16")


def mlp_problem(segments, M, N, *, a_mn=False, b_mn=False, segments2=None, split_m=0, bias=None, bias2=None,
                relu=False, mask_f32=None, mask_bf16=None, out_f32=None, perm=None, out_hi=None, out_lo=None,
                out_bf16=None, colsum=None, colsum_mod=0, BN=64):
    """One GEMM of the Q-head MLP as a descriptor: D[M, N] = sum over `segments` [(A, B, K), ...] of
    A[m, :K] . B[n, :K] (bf16 operands, fp32 accumulation; include/vdqn.h).  A / B are 2-D bf16 tensors:
    [M or N, >= K] K-major, or with a_mn / b_mn [>= K, M or N] MN-major.  `segments2` (+ bias2, split_m): the
    B operands of rows >= split_m.  Returns (descriptor, tensors it points into)."""
    d = L.MlpGemmDesc()
    _req(1 <= len(segments) <= 3, "mlp_gemm: 1..3 segments")
    keep = []

    def fill(op, t, name):
        _req(t.is_cuda and t.dtype == bf16 and t.dim() == 2 and t.stride(1) == 1, f"{name}: 2-D bf16, unit column stride")
        op.ptr, op.rows, op.cols, op.ld = t.data_ptr(), t.shape[0], t.shape[1], t.stride(0)
        keep.append(t)
    for i, (A, B, K) in enumerate(segments):
        fill(d.a[i], A, "A"); fill(d.b[i], B, "B")
        d.K[i] = K
        if segments2 is not None:
            fill(d.b2[i], segments2[i][1], "B2")
    d.nseg, d.M, d.N, d.BN = len(segments), M, N, BN
    d.a_mn, d.b_mn, d.split_m = int(a_mn), int(b_mn), int(split_m if segments2 is not None else 0)
    d.bias, d.bias2, d.relu = L.ptr(bias), L.ptr(bias2), int(relu)
    if mask_f32 is not None:
        d.mask_f32, d.ldmask = mask_f32.data_ptr(), mask_f32.stride(0)
    if mask_bf16 is not None:
        d.mask_bf16, d.ldmask = mask_bf16.data_ptr(), mask_bf16.stride(0)
    if out_f32 is not None:
        _req(out_f32.dtype == torch.float32 and out_f32.stride(-1) == 1, "out_f32")
        d.out_f32, d.ld_f32 = out_f32.data_ptr(), out_f32.stride(0)
        if perm is not None:
            d.perm_c, d.perm_p = perm
    if out_hi is not None:
        _req(out_hi.shape == out_lo.shape and out_hi.dtype == bf16, "out_hi / out_lo")
        d.out_hi, d.out_lo, d.ld_hl = out_hi.data_ptr(), out_lo.data_ptr(), out_hi.stride(0)
    if out_bf16 is not None:
        _req(out_bf16.dtype == bf16, "out_bf16")
        d.out_bf16, d.ld_bf16 = out_bf16.data_ptr(), out_bf16.stride(0)
    if colsum is not None:
        _cuda(colsum, torch.float32, "colsum")
        d.colsum, d.colsum_mod = colsum.data_ptr(), colsum_mod
    return d, keep


def mlp_gemm_grouped(problems):
    """Up to four independent `mlp_problem`s in one launch."""
    lib = L.load()
    arr = (L.MlpGemmDesc * len(problems))(*[p[0] for p in problems])
    with _Prof("mlp", tuple((p[0].M, p[0].N) for p in problems)):
        L.check(lib.vdqn_mlp_gemm_grouped(arr, len(problems), L.stream_ptr()), "mlp_gemm")


def mlp_gemm(segments, M, N, **kw):
    mlp_gemm_grouped([mlp_problem(segments, M, N, **kw)])


def relu_mask_colsum(dy, y, db, relu=True, out_bf16=None):
    """dy <- dy * (y > 0) in place (fp32 [B, O]); db[o] = sum_b dy[b, o]; optional bf16 copy."""
    lib = L.load()
    B, O = dy.shape
    with _Prof("mlp", (B, O)):
        L.check(lib.vdqn_relu_mask_colsum(dy.data_ptr(), L.ptr(y), L.ptr(out_bf16), db.data_ptr(), B, O,
                                          int(relu), L.stream_ptr()), "relu_mask_colsum")
    return dy


def avgpool_fwd(x, out=None):
    """x [N,H,W,C] bf16 -> [N,C] fp32 mean over pixels."""
    lib = L.load()
    _cuda(x, bf16, "x")
    N, Cc = x.shape[0], x.shape[-1]
    P = x.numel() // max(N * Cc, 1)
    if out is None:
        out = torch.empty(N, Cc, device=x.device, dtype=torch.float32)
    with _Prof("mlp", (N,)):
        L.check(lib.vdqn_avgpool_fwd(x.data_ptr(), out.data_ptr(), N, P, Cc, L.stream_ptr()), "avgpool_fwd")
    return out


def head_flatten_fwd(h, flat=None):
    lib = L.load()
    _cuda(h, bf16, "h")
    B = h.shape[0]
    Cc = h.shape[-1]
    P = h.numel() // (B * Cc)
    if flat is None:
        flat = torch.empty(B, Cc * P, device=h.device, dtype=torch.float32)
    with _Prof("mlp", (B,)):
        L.check(lib.vdqn_head_flatten_fwd(h.data_ptr(), flat.data_ptr(), B, P, Cc, L.stream_ptr()),
                "head_flatten_fwd")
    return flat


def head_flatten_bwd(dflat, h, dh=None, dbias=None):
    lib = L.load()
    B = h.shape[0]
    Cc = h.shape[-1]
    P = h.numel() // (B * Cc)
    if dh is None:
        dh = torch.empty_like(h)
    with _Prof("mlp", (B,)):
        L.check(lib.vdqn_head_flatten_bwd(dflat.data_ptr(), h.data_ptr(), dh.data_ptr(), L.ptr(dbias),
                                          B, P, Cc, L.stream_ptr()), "head_flatten_bwd")
    return dh


def td_epilogue(q_s, q_next_online, q_next_target, act, rew, term, valid=None, *, gamma=0.99,
                double_dqn=True, clip_rect=True, linear=False, use_valid=False, inv_count=None,
                dq=None, loss=None, best=None, y=None, want_aux=False, gt=None, value_learning=False):
    """Fused Double-DQN TD loss + gradient.  Returns (loss[1] fp32, dq [B,C,A] fp32, best, y).
    `gt` (float64 [B,C]) switches to the ground-truth regression of
    process_batch(compare_ground_truth=True) (train_q_network.py:170-178); q_next_*, rew, term may
    then be None."""
    lib = L.load()
    _cuda(q_s, torch.float32, "q_s")
    _req(q_s.dim() == 3, "bad shape")
    B, Cc, A = q_s.shape
    _cuda(act, torch.int64, "act")
    _req(act.numel() == B, "bad shape")
    if gt is None:
        _cuda(q_next_target, torch.float32, "q_next_target")
        _req(q_s.shape == q_next_target.shape, "bad shape")
        # int64 labels (thresholded detections) or, with CONFIDENCE_REWARD, the detector scores the
        # reference casts with `.float()` (train_q_network.py:158-160): fp32, all label arrays alike
        ldt = torch.float32 if rew.dtype == torch.float32 else torch.int64
        for t, n in ((rew, "rew"), (term, "term")):
            _cuda(t, ldt, n)
        _req(rew.numel() == B * Cc and term.numel() == B * Cc, "bad shape")
        if q_next_online is not None:
            _cuda(q_next_online, torch.float32, "q_next_online")
            _req(q_next_online.shape == q_s.shape, "bad shape")
    else:
        _cuda(gt, torch.float64, "ground_truth")
        _req(gt.numel() == B * Cc, "bad shape")
    labels_f32 = gt is None and rew.dtype == torch.float32
    if use_valid:
        _cuda(valid, torch.float32 if labels_f32 else torch.int64, "valid_mask")
        _req(valid.numel() == B * Cc, "bad shape")
    dev = q_s.device
    if dq is None:
        dq = torch.empty_like(q_s)
    if loss is None:
        loss = torch.zeros(1, device=dev, dtype=torch.float32)
    if want_aux:
        best = torch.empty(B, Cc, device=dev, dtype=torch.int64) if best is None else best
        y = torch.empty(B, Cc, device=dev, dtype=torch.float32) if y is None else y
    d = L.TdDesc()
    d.q_s, d.q_next_online, d.q_next_target = q_s.data_ptr(), L.ptr(q_next_online), L.ptr(q_next_target)
    d.act, d.rew, d.term, d.valid = act.data_ptr(), L.ptr(rew), L.ptr(term), L.ptr(valid)
    d.dq, d.loss_out, d.best_out, d.y_out = dq.data_ptr(), loss.data_ptr(), L.ptr(best), L.ptr(y)
    d.B, d.C, d.A = B, Cc, A
    d.gamma = gamma
    d.inv_count = (1.0 / (B * Cc)) if inv_count is None else inv_count
    d.double_dqn, d.clip_rect, d.linear, d.use_valid = int(double_dqn), int(clip_rect), int(linear), int(use_valid)
    d.gt, d.ground_truth, d.value_learning = L.ptr(gt), int(gt is not None), int(value_learning)
    d.labels_f32 = int(labels_f32)
    with _Prof("td", (B, Cc, A)):
        L.check(lib.vdqn_td_epilogue(C.byref(d), L.stream_ptr()), "td_epilogue")
    return loss, dq, best, y


def q_max(q, value=None, arg=None, want_arg=False):
    """q [B,C,A] fp32 -> value [B,C] = max over actions (and the first arg-max)."""
    lib = L.load()
    _cuda(q, torch.float32, "q")
    _req(q.dim() == 3, "bad shape")
    B, Cc, A = q.shape
    if value is None:
        value = torch.empty(B, Cc, device=q.device, dtype=torch.float32)
    if want_arg and arg is None:
        arg = torch.empty(B, Cc, device=q.device, dtype=torch.int64)
    L.check(lib.vdqn_q_max(q.data_ptr(), value.data_ptr(), L.ptr(arg), B * Cc, A, L.stream_ptr()), "q_max")
    return value, arg


class BnBatchStats:
    """Per-layer scratch of a train-mode BatchNorm: fp64 sums, fp32 mean / rstd / scale / shift."""

    def __init__(self, C, device):
        self.C = C
        self.sums = torch.zeros(2 * C, device=device, dtype=torch.float64)
        self.mean = torch.empty(C, device=device, dtype=torch.float32)
        self.rstd = torch.empty(C, device=device, dtype=torch.float32)
        self.scale = torch.empty(C, device=device, dtype=torch.float32)
        self.shift = torch.empty(C, device=device, dtype=torch.float32)
        self.bsums = torch.zeros(2 * C, device=device, dtype=torch.float64)     # backward: sum dy, sum dy*xhat


class BnSync:
    """SyncBatchNorm plumbing for the data-parallel `basic` architecture: the per-channel fp64
    sums of a train-mode BatchNorm are all-reduced over the ranks between the statistics kernel and the
    finalize kernel (forward) and between the reduction and the apply kernel (backward), so every rank
    normalises with the statistics of the GLOBAL batch (count M * world)."""

    def __init__(self, world: int, group=None):
        self.world, self.group = world, group

    def all_reduce(self, t):
        import torch.distributed as dist
        dist.all_reduce(t, group=self.group)


def bn_train_fwd(x, st: "BnBatchStats", gamma, beta, running_mean, running_var, nbt, y, *, residual=None,
                 relu=False, momentum=0.1, eps=1e-5, update_running=True, sync=None):
    """Train-mode BatchNorm2d (+ residual) (+ ReLU) on raw conv output x [.., C] bf16 -> y bf16; batch
    statistics stay in `st` for the backward pass; running statistics / num_batches_tracked are updated
    in place (torch.nn.BatchNorm2d.forward)."""
    lib = L.load()
    _cuda(x, bf16, "x"); _cuda(y, bf16, "y")
    Cc = x.shape[-1]
    M = x.numel() // Cc
    _req(Cc == st.C and y.numel() == x.numel() and M >= 1, "bad shape")
    for t, n in ((gamma, "gamma"), (beta, "beta"), (running_mean, "running_mean"), (running_var, "running_var")):
        _cuda(t, torch.float32, n); _req(t.numel() == Cc, "bad shape")
    if residual is not None:
        _cuda(residual, bf16, "residual"); _req(residual.numel() == x.numel(), "bad shape")
    if nbt is not None:
        _cuda(nbt, torch.int64, "num_batches_tracked")
    sp = L.stream_ptr()
    with _Prof("bn", (M, Cc)):
        L.check(lib.vdqn_bn_stats(x.data_ptr(), st.sums.data_ptr(), M, Cc, sp), "bn_stats")
        Mg = M
        if sync is not None:
            sync.all_reduce(st.sums)
            Mg = M * sync.world
        L.check(lib.vdqn_bn_finalize(st.sums.data_ptr(), Mg, Cc, gamma.data_ptr(), beta.data_ptr(),
                                     running_mean.data_ptr() if update_running else None,
                                     running_var.data_ptr() if update_running else None,
                                     L.ptr(nbt) if update_running else None, momentum, eps,
                                     st.mean.data_ptr(), st.rstd.data_ptr(), st.scale.data_ptr(),
                                     st.shift.data_ptr(), sp), "bn_finalize")
        L.check(lib.vdqn_bn_apply(x.data_ptr(), st.scale.data_ptr(), st.shift.data_ptr(), L.ptr(residual),
                                  int(relu), y.data_ptr(), M, Cc, sp), "bn_apply")
    return y


def bn_train_bwd(dy, x, st: "BnBatchStats", gamma, dgamma, dbeta, dx, sync=None):
    """dy = gradient w.r.t. the BatchNorm output (bf16, already masked by the following ReLU), x = the
    raw conv output the forward normalised.  Writes dgamma / dbeta (fp32, overwritten) and dx (bf16, may
    alias dy)."""
    lib = L.load()
    _cuda(dy, bf16, "dy"); _cuda(x, bf16, "x"); _cuda(dx, bf16, "dx")
    Cc = x.shape[-1]
    M = x.numel() // Cc
    _req(Cc == st.C and dy.numel() == x.numel() and dx.numel() == x.numel() and M >= 1, "bad shape")
    _cuda(dgamma, torch.float32, "dgamma"); _cuda(dbeta, torch.float32, "dbeta")
    _req(dgamma.numel() == Cc and dbeta.numel() == Cc, "bad shape")
    sp = L.stream_ptr()
    with _Prof("bn", (M, Cc)):
        L.check(lib.vdqn_bn_bwd_reduce(dy.data_ptr(), x.data_ptr(), st.mean.data_ptr(), st.rstd.data_ptr(),
                                       st.bsums.data_ptr(), M, Cc, sp), "bn_bwd_reduce")
        dg, db, Mg = dgamma, dbeta, M
        if sync is not None:
            # d gamma / d beta are this rank's OWN sums (the gradient exchange adds the ranks up afterwards); dx
            # needs the sums of the global batch
            dbeta.copy_(st.bsums[:Cc]); dgamma.copy_(st.bsums[Cc:])
            sync.all_reduce(st.bsums)
            if getattr(st, "scratch", None) is None:
                st.scratch = torch.empty(2 * Cc, device=x.device, dtype=torch.float32)
            dg, db, Mg = st.scratch[:Cc], st.scratch[Cc:], M * sync.world
        L.check(lib.vdqn_bn_bwd_apply(dy.data_ptr(), x.data_ptr(), st.mean.data_ptr(), st.rstd.data_ptr(),
                                      gamma.data_ptr(), st.bsums.data_ptr(), dg.data_ptr(), db.data_ptr(),
                                      dx.data_ptr(), Mg, Cc, sp), "bn_bwd_apply")
    return dx


def avgpool_bwd(dpooled, feat, dfeat=None):
    """dfeat[n,p,c] = feat > 0 ? dpooled[n,c] / P : 0 (bf16): gradient of mean-over-pixels through the last
    block's ReLU."""
    lib = L.load()
    _cuda(dpooled, torch.float32, "dpooled"); _cuda(feat, bf16, "feat")
    N, Cc = feat.shape[0], feat.shape[-1]
    P = feat.numel() // max(N * Cc, 1)
    _req(dpooled.numel() == N * Cc, "bad shape")
    if dfeat is None:
        dfeat = torch.empty_like(feat)
    with _Prof("mlp", (N,)):
        L.check(lib.vdqn_avgpool_bwd(dpooled.data_ptr(), feat.data_ptr(), dfeat.data_ptr(), N, P, Cc,
                                     L.stream_ptr()), "avgpool_bwd")
    return dfeat


def cross_entropy(logits, labels, *, dlogits=None, loss=None, correct=None, inv_count=None, want_grad=True):
    """`nn.CrossEntropyLoss()(logits, labels)` (mean) + gradient + number of correct arg-max predictions
    (train_inverse_model.py:100-106).  `loss` (fp32[1]) and `correct` (int32[1]) are ACCUMULATED into:
    zero them first.  Returns (loss, dlogits, correct)."""
    lib = L.load()
    _cuda(logits, torch.float32, "logits"); _cuda(labels, torch.int64, "labels")
    _req(logits.dim() == 2 and labels.numel() == logits.shape[0], "bad shape")
    B, Cc = logits.shape
    dev = logits.device
    if want_grad and dlogits is None:
        dlogits = torch.empty_like(logits)
    if loss is None:
        loss = torch.zeros(1, device=dev, dtype=torch.float32)
    if correct is None:
        correct = torch.zeros(1, device=dev, dtype=torch.int32)
    with _Prof("ce", (B, Cc)):
        L.check(lib.vdqn_cross_entropy(logits.data_ptr(), labels.data_ptr(), L.ptr(dlogits) if want_grad else None,
                                       loss.data_ptr(), correct.data_ptr(), B, Cc,
                                       (1.0 / max(B, 1)) if inv_count is None else inv_count, L.stream_ptr()),
                "cross_entropy")
    return loss, dlogits, correct


def dropout_mask(keep, p, seed, counter):
    """keep (uint8, any shape) <- Bernoulli(1 - p) from the counter-based generator."""
    lib = L.load()
    _cuda(keep, torch.uint8, "keep")
    L.check(lib.vdqn_dropout_mask(keep.data_ptr(), keep.numel(), float(p), int(seed) & (2 ** 64 - 1),
                                  int(counter), L.stream_ptr()), "dropout_mask")
    return keep


def dropout_apply(x, keep, scale, y=None):
    """y = keep ? x * scale : 0 (y may be x)."""
    lib = L.load()
    _cuda(x, torch.float32, "x"); _cuda(keep, torch.uint8, "keep")
    _req(keep.numel() == x.numel(), "bad shape")
    if y is None:
        y = torch.empty_like(x)
    L.check(lib.vdqn_dropout_apply(x.data_ptr(), keep.data_ptr(), float(scale), y.data_ptr(), x.numel(),
                                   L.stream_ptr()), "dropout_apply")
    return y


def adam_fused(p, g, m, v, *, lr, step=None, betas=(0.9, 0.999), eps=1e-8, target=None, grad_scale=1.0,
               step_dev=None, scalars_dev=None):
    """Fused Adam over flat fp32 arenas.  Either `step` (host int) or `step_dev` + `scalars_dev`
    (int32[1] / float32[2] device tensors; graph-replay mode, the call increments the counter)."""
    lib = L.load()
    for t, n in ((p, "p"), (g, "g"), (m, "m"), (v, "v")):
        _cuda(t, torch.float32, n)
    n = p.numel()
    _req(g.numel() == n and m.numel() == n and v.numel() == n, "bad shape")
    if target is not None:
        _cuda(target, torch.float32, "target"); _req(target.numel() == n, "bad shape")
    if step_dev is not None:
        _cuda(step_dev, torch.int32, "step_dev"); _cuda(scalars_dev, torch.float32, "scalars_dev")
        with _Prof("adam", (n, target is not None)):
          L.check(lib.vdqn_adam_fused_graph(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(),
                                          L.ptr(target), n, lr, betas[0], betas[1], eps, grad_scale,
                                          step_dev.data_ptr(), scalars_dev.data_ptr(), L.stream_ptr()),
                "adam_fused_graph")
    else:
        L.check(lib.vdqn_adam_fused(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), L.ptr(target),
                                    n, lr, betas[0], betas[1], eps, int(step), grad_scale, L.stream_ptr()),
                "adam_fused")
