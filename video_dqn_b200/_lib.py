"""ctypes binding of libvdqn.so (include/vdqn.h).  The product path has no fallback: if the
library is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvdqn.so")

EPI_RELU = 1
EPI_OUT_F32 = 2
EPI_SCATTER_INPUTS = 4

c_void_p, c_int, c_float, c_int64 = C.c_void_p, C.c_int32, C.c_float, C.c_int64


class ConvDesc(C.Structure):
    _fields_ = [("x", c_void_p), ("w", c_void_p), ("out", c_void_p), ("out2", c_void_p),
                ("shift", c_void_p), ("residual", c_void_p), ("mask_src", c_void_p),
                ("colsum", c_void_p)] + \
               [(n, c_int) for n in ("N", "H", "W", "Cin", "Cout", "R", "S", "stride", "dil",
                                     "pad_lo", "pad_hi", "ldc", "ldr", "ldm", "out2_ld",
                                     "out_scatter", "flags", "tile_n", "max_ctas", "algo", "pad_hi_w",
                                     "scatter_off_h", "scatter_off_w")] + \
               [("w2", c_void_p), ("shift2", c_void_p), ("split_n", c_int), ("x_alias_from", c_int),
                ("x_alias_shift", c_int), ("pool_out", c_void_p), ("pool_idx", c_void_p), ("pool_idx_images", c_int),
                ("x2", c_void_p), ("Cin2", c_int), ("H2", c_int), ("W2", c_int), ("stride2", c_int)]


class WgradDesc(C.Structure):
    _fields_ = [("x", c_void_p), ("dy", c_void_p), ("part", c_void_p)] + \
               [(n, c_int) for n in ("N", "H", "W", "Cin", "Cout", "R", "S", "stride", "dil",
                                     "pad_lo", "pad_hi", "ldy", "splits", "max_ctas", "algo")]


class WgradFinDesc(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("part", "w", "gamma", "var", "mean", "dbeta", "dw", "dgamma")] + \
               [(n, c_int) for n in ("splits", "Cout", "Cin", "R", "S", "K", "kmap")] + [("eps", c_float)]


class WgradFinItem(C.Structure):
    _fields_ = [("d", WgradFinDesc)] + [(n, c_int) for n in ("cn", "items", "TX", "TY", "nchunks", "first_block")]


class WprepDesc(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("w", "gamma", "beta", "mean", "var", "bias", "w_fwd",
                                        "w_dgrad", "shift")] + \
               [(n, c_int) for n in ("Cout", "Cin", "R", "S", "K", "kmap")] + [("eps", c_float),
                                                                                ("dgrad_parity", c_int)] + \
               [(n, c_int) for n in ("ldw_fwd", "fwd_col0", "ldw_dgrad", "dgrad_col0")] + \
               [(n, c_void_p) for n in ("gamma_b", "beta_b", "mean_b", "var_b")]


class TdDesc(C.Structure):
    _fields_ = [(n, c_void_p) for n in ("q_s", "q_next_online", "q_next_target", "act", "rew", "term",
                                        "valid", "dq", "loss_out", "best_out", "y_out")] + \
               [(n, c_int) for n in ("B", "C", "A")] + [("gamma", c_float), ("inv_count", c_float)] + \
               [(n, c_int) for n in ("double_dqn", "clip_rect", "linear", "use_valid")] + \
               [("gt", c_void_p), ("ground_truth", c_int), ("value_learning", c_int), ("labels_f32", c_int)]


class MlpOperand(C.Structure):
    _fields_ = [("ptr", c_void_p), ("rows", c_int), ("cols", c_int), ("ld", c_int)]


class MlpGemmDesc(C.Structure):
    _fields_ = [("a", MlpOperand * 3), ("b", MlpOperand * 3), ("b2", MlpOperand * 3), ("K", c_int * 3)] + \
               [(n, c_int) for n in ("nseg", "M", "N", "BN", "a_mn", "b_mn", "split_m")] + \
               [("bias", c_void_p), ("bias2", c_void_p), ("relu", c_int), ("mask_f32", c_void_p),
                ("mask_bf16", c_void_p), ("ldmask", c_int), ("out_f32", c_void_p), ("ld_f32", c_int),
                ("perm_c", c_int), ("perm_p", c_int), ("out_hi", c_void_p), ("out_lo", c_void_p), ("ld_hl", c_int),
                ("out_bf16", c_void_p), ("ld_bf16", c_int), ("colsum", c_void_p), ("colsum_mod", c_int)]


class NvlDesc(C.Structure):
    _fields_ = [("peer_bufs", C.POINTER(c_void_p)), ("peer_flags", C.POINTER(c_void_p)), ("multicast_ptr", c_void_p),
                ("epoch", c_void_p), ("counter", c_void_p), ("n", c_int64), ("rank", c_int), ("world", c_int),
                ("max_ctas", c_int), ("first", c_int64), ("channel", c_int), ("threads", c_int)]


EXPORTS = {
    "vdqn_last_error": (C.c_char_p, []),
    "vdqn_abi_version": (c_int, []),
    "vdqn_init": (c_int, [c_int]),
    "vdqn_num_sms": (c_int, []),
    "vdqn_launch_count": (C.c_longlong, []),
    "vdqn_zero": (c_int, [c_void_p, c_int64, c_void_p]),
    "vdqn_conv_gemm": (c_int, [C.POINTER(ConvDesc), c_void_p]),
    "vdqn_conv_wgrad": (c_int, [C.POINTER(WgradDesc), c_void_p]),
    "vdqn_wgrad_finalize": (c_int, [C.POINTER(WgradFinDesc), c_void_p]),
    "vdqn_wgrad_finalize_plan": (c_int, [C.POINTER(WgradFinDesc), C.POINTER(WgradFinItem)]),
    "vdqn_wgrad_finalize_multi": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "vdqn_weight_prep": (c_int, [C.POINTER(WprepDesc), c_void_p]),
    "vdqn_weight_prep_multi": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_void_p]),
    "vdqn_weight_prep_tiled": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "vdqn_stem_pack_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vdqn_stem_pack_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vdqn_maxpool_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "vdqn_maxpool_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int, c_int, c_int, c_int, c_void_p]),
    "vdqn_linear_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "vdqn_linear_bwd": (c_int, [c_void_p] * 7 + [c_int, c_int, c_int, c_int, c_void_p]),
    "vdqn_relu_mask_colsum": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vdqn_avgpool_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vdqn_head_flatten_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vdqn_head_flatten_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vdqn_mlp_gemm": (c_int, [C.POINTER(MlpGemmDesc), c_void_p]),
    "vdqn_mlp_gemm_grouped": (c_int, [C.POINTER(MlpGemmDesc), c_int, c_void_p]),
    "vdqn_split_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vdqn_nvl_allreduce": (c_int, [C.POINTER(NvlDesc), c_void_p]),
    "vdqn_td_epilogue": (c_int, [C.POINTER(TdDesc), c_void_p]),
    "vdqn_q_max": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "vdqn_bn_stats": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "vdqn_bn_finalize": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                 c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vdqn_bn_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int, c_void_p]),
    "vdqn_bn_bwd_reduce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "vdqn_bn_bwd_apply": (c_int, [c_void_p] * 9 + [c_int64, c_int, c_void_p]),
    "vdqn_avgpool_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vdqn_cross_entropy": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float,
                                   c_void_p]),
    "vdqn_dropout_mask": (c_int, [c_void_p, c_int64, c_float, C.c_uint64, C.c_uint64, c_void_p]),
    "vdqn_dropout_apply": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_int64, c_void_p]),
    "vdqn_adam_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                C.c_double, C.c_double, C.c_double, C.c_double, c_int, c_float, c_void_p]),
    "vdqn_adam_fused_graph": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                      C.c_double, C.c_double, C.c_double, C.c_double, c_float,
                                      c_void_p, c_void_p, c_void_p]),
}

_lib = None


class VdqnError(RuntimeError):
    pass


def load():
    """dlopen libvdqn.so and type its entry points.  Raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VdqnError(f"{LIB_PATH} not found: build it with `python video_dqn_b200/build.py` "
                        "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().vdqn_last_error().decode(errors="replace")
        if rc == -2 or rc == -1:
            raise ValueError(f"vdqn {what}: {msg} (code {rc})")
        raise VdqnError(f"vdqn {what}: {msg} (code {rc})")


def ptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
