"""Forward / backward schedule of the Q-network over the C-ABI kernels.

The network is the reference's `extra_capacity` HabitatDQNMultiAction
(archs/HabitatDQNMultiAction.py:27-31,44-54): ResNet-18 trunk (torchvision BasicBlocks,
BatchNorm in eval mode during training, :37-40) -> Conv2d(512,64,3)+ReLU -> Flatten ->
Linear 1600F-512-256-5A.  Layout in HBM: activations NHWC bf16, trunk/head conv weights bf16
with the BN scale folded in (both the forward [Cout][R][S][Cin] and the data-gradient
[Cin][R][S][Cout] flipped form), fp32 master parameters / gradients / Adam state in flat arenas,
the Q-head MLP entirely in fp32.

Everything here only *enqueues* kernels on the current stream (no host sync, no allocation when
a preallocated Workspace is passed), so a whole step can be captured in a CUDA graph.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import os as _os

import torch

from . import ops

bf16 = torch.bfloat16
BN_EPS = 1e-5
WGRAD_WAVES = int(_os.environ.get("VDQN_WGRAD_WAVES", "1"))   # work units per SM for the split weight gradient


@dataclass
class ConvSpec:
    name: str            # short id
    wkey: str            # state-dict key of the OIHW fp32 weight (under resnet.* / features.8)
    bn: Optional[str]    # BN prefix or None
    bias: Optional[str]  # conv bias key or None
    cin: int
    cout: int
    k: int               # filter size of the GEMM form (4 for the packed stem)
    stride: int          # stride of the GEMM form
    pad_lo: int
    pad_hi: int
    in_hw: int           # input spatial size of the GEMM form
    out_hw: int
    kmap: int = 0
    gemm_cin: int = 0    # channels of the GEMM form (16 for the packed stem)

    @property
    def K(self):
        return self.k * self.k * self.gemm_cin


@dataclass
class BlockSpec:
    conv1: ConvSpec
    conv2: ConvSpec
    ds: Optional[ConvSpec]
    in_hw: int
    out_hw: int
    cin: int
    cout: int
    stride: int


@dataclass
class NetPlan:
    stem: ConvSpec
    blocks: List[BlockSpec]
    head: ConvSpec
    action_dim: int
    num_classes: int
    num_frames: int
    convs: List[ConvSpec] = field(default_factory=list)


def make_plan(action_dim: int, num_classes: int = 5, num_frames: int = 1) -> NetPlan:
    stem = ConvSpec("stem", "resnet.conv1.weight", "resnet.bn1", None, 3, 64, 4, 1, 2, 1, 112, 112,
                    kmap=1, gemm_cin=16)
    blocks = []
    hw, cin = 56, 64
    for li, (cout, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2)), start=1):
        for b in range(2):
            s = stride if b == 0 else 1
            ohw = hw // s
            p = f"resnet.layer{li}.{b}."
            c1 = ConvSpec(f"l{li}.{b}.c1", p + "conv1.weight", p + "bn1", None, cin, cout, 3, s, 1, 1,
                          hw, ohw, gemm_cin=cin)
            c2 = ConvSpec(f"l{li}.{b}.c2", p + "conv2.weight", p + "bn2", None, cout, cout, 3, 1, 1, 1,
                          ohw, ohw, gemm_cin=cout)
            ds = None
            if b == 0 and s != 1:
                ds = ConvSpec(f"l{li}.{b}.ds", p + "downsample.0.weight", p + "downsample.1", None,
                              cin, cout, 1, s, 0, 0, hw, ohw, gemm_cin=cin)
            blocks.append(BlockSpec(c1, c2, ds, hw, ohw, cin, cout, s))
            hw, cin = ohw, cout
    head = ConvSpec("head", "features.8.weight", None, "features.8.bias", 512, 64, 3, 1, 0, 0, 7, 5,
                    gemm_cin=512)
    plan = NetPlan(stem, blocks, head, action_dim, num_classes, num_frames)
    plan.convs = [stem] + [c for b in blocks for c in (b.conv1, b.conv2, b.ds) if c is not None] + [head]
    return plan


def prep_convs(plan: NetPlan) -> List[ConvSpec]:
    """every tensor that has bf16 GEMM operands"""
    return plan.convs


def wgrad_uses_halo(spec: ConvSpec) -> bool:
    """64->64 3x3 stride-1 convolutions (layer1) and the packed stem take the halo-tile weight-gradient
    kernels."""
    if spec.kmap == 1:                       # packed stem
        return True
    return (spec.kmap == 0 and spec.cin == 64 and spec.cout == 64 and spec.k == 3 and spec.stride == 1
            and spec.pad_lo == 1)


def wgrad_splits(spec: ConvSpec, n_img: int, sms: int = 0) -> int:
    if wgrad_uses_halo(spec):
        tiles = n_img * ((spec.out_hw + 7) // 8) * ((spec.out_hw + 15) // 16)
        return min(sms or ops.num_sms(), tiles)    # = number of persistent CTAs
    M = n_img * spec.out_hw * spec.out_hw
    co_tiles = (spec.cout + 127) // 128
    per = 4 if spec.gemm_cin % 64 == 0 else 16
    slab = 64 if spec.gemm_cin % 64 == 0 else 16
    groups = -(-(spec.K // slab) // per)
    # every extra split costs a Cout x K fp32 partial written and re-read: one wave of units, no more
    s = max(1, (WGRAD_WAVES * (sms or ops.num_sms())) // (co_tiles * groups))
    return int(max(1, min(s, M // 512 if M >= 512 else 1)))


class PreparedWeights:
    """bf16 GEMM operands + fp32 shifts for every conv, regenerated from the fp32 masters."""

    def __init__(self, plan: NetPlan, device, trunk_only: bool = False, fold_bn: bool = True):
        """trunk_only: the ResNet trunk alone (the inverse-dynamics model reuses it under its own head).
        fold_bn=False: plain bf16 casts of the weights (scale 1, shift 0) -- the train-mode BatchNorm path
        of the `basic` architecture normalises with batch statistics in separate kernels."""
        self.plan = plan
        self.fold_bn = fold_bn
        self.w_fwd: Dict[str, torch.Tensor] = {}
        self.w_dgrad: Dict[str, torch.Tensor] = {}
        self.shift: Dict[str, torch.Tensor] = {}
        self.convs = [c for c in prep_convs(plan) if not (trunk_only and c.name == "head")]
        # Q-head MLP operands (mlp_gemm.cu): every fp32 weight as hi + lo bf16; top.0's columns in the NHWC
        # order of the head-conv output (p*64 + c per frame instead of the reference's Flatten order c*25 + p)
        self.mlp: Dict[str, tuple] = {}
        if not trunk_only:
            nq = plan.num_classes * plan.action_dim
            for key, (o, k) in (("top.0", (512, 1600 * plan.num_frames)), ("top.2", (256, 512)), ("top.4", (nq, 256))):
                self.mlp[key] = (torch.empty(o, k, device=device, dtype=bf16), torch.empty(o, k, device=device, dtype=bf16))
        # Strided residual blocks with folded BatchNorm: the 1x1/2 downsample is not a launch of its own.
        # Forward: its weights are appended to conv2's rows ([Cout][9 Cout | Cin]) and conv2's kernel reduces
        # over a1 AND the block input (second K segment); its shift is added to conv2's.  Backward: conv1's
        # strided data gradient is ONE dense GEMM over a [4 Cin][2x2 taps x Cout | Cout] matrix -- the four
        # output-parity classes as column groups, zeros where a class does not use a tap -- and the last Cout
        # columns are the downsample's data gradient (class (0, 0) only), reduced over the block's output gradient.
        self.fuse_ds = bool(fold_bn) and FUSE_DS
        fused_c2 = {b.conv2.name: b for b in plan.blocks if b.ds is not None} if self.fuse_ds else {}
        fused_c1 = {b.conv1.name: b for b in plan.blocks if b.ds is not None} if self.fuse_ds else {}
        fused_ds = {b.ds.name: b for b in plan.blocks if b.ds is not None} if self.fuse_ds else {}
        self._fused = (fused_c2, fused_c1, fused_ds)
        for c in self.convs:
            if c.name in fused_ds:
                pass                                                       # lives inside conv2 / conv1 buffers
            elif c.name in fused_c2:
                self.w_fwd[c.name] = torch.empty(c.cout, c.K + fused_c2[c.name].cin, device=device, dtype=bf16)
            else:
                self.w_fwd[c.name] = torch.empty(c.cout, c.k, c.k, c.gemm_cin, device=device, dtype=bf16)
            if c.kmap == 0 and c.name in fused_c1:
                self.w_dgrad[c.name] = torch.zeros(4 * c.cin, 5 * c.cout, device=device, dtype=bf16)
            elif c.kmap == 0 and c.name not in fused_ds:
                self.w_dgrad[c.name] = torch.empty(c.cin, c.k, c.k, c.cout, device=device, dtype=bf16)
            self.shift[c.name] = torch.empty(c.cout, device=device, dtype=torch.float32)

    def prepare(self, P: Dict[str, torch.Tensor]):
        """Re-derive every bf16 operand from the fp32 masters in ONE launch.  The descriptor table
        lives on the device and is rebuilt only when a parameter's storage moved."""
        from . import _lib as L
        convs = self.convs
        sig = tuple(P[c.wkey].data_ptr() for c in convs) + \
            tuple(P[c.bn + ".weight"].data_ptr() for c in convs if c.bn)
        if getattr(self, "_table_sig", None) != sig:
            # the stem (space-to-depth packing) goes through the element-wise kernel, everything else
            # through the tiled (coalesced) one
            dev = self.shift["stem"].device

            fused_c2, fused_c1, fused_ds = self._fused

            def fill(d, c):
                w = P[c.wkey]
                d.w, d.shift = w.data_ptr(), self.shift[c.name].data_ptr()
                if c.name in fused_ds:
                    b = fused_ds[c.name]          # columns [9 Cout, 9 Cout + Cin) of conv2's rows; last Cout columns of conv1's
                    d.w_fwd, d.ldw_fwd, d.fwd_col0 = self.w_fwd[b.conv2.name].data_ptr(), b.conv2.K + b.cin, b.conv2.K
                    d.w_dgrad, d.ldw_dgrad, d.dgrad_col0 = self.w_dgrad[b.conv1.name].data_ptr(), 5 * b.cout, 4 * b.cout
                else:
                    d.w_fwd = self.w_fwd[c.name].data_ptr()
                    d.w_dgrad = L.ptr(self.w_dgrad.get(c.name))
                if c.name in fused_c2:
                    b = fused_c2[c.name]
                    d.ldw_fwd = c.K + b.cin
                    bn2 = b.ds.bn
                    d.gamma_b, d.beta_b = P[bn2 + ".weight"].data_ptr(), P[bn2 + ".bias"].data_ptr()
                    d.mean_b, d.var_b = P[bn2 + ".running_mean"].data_ptr(), P[bn2 + ".running_var"].data_ptr()
                if c.name in fused_c1:
                    d.ldw_dgrad = 5 * c.cout
                if c.bn is not None and self.fold_bn:
                    d.gamma, d.beta = P[c.bn + ".weight"].data_ptr(), P[c.bn + ".bias"].data_ptr()
                    d.mean, d.var = P[c.bn + ".running_mean"].data_ptr(), P[c.bn + ".running_var"].data_ptr()
                if c.bias is not None:
                    d.bias = P[c.bias].data_ptr()
                if w.dim() == 4:
                    d.Cout, d.Cin, d.R, d.S = w.shape
                else:                                   # top.0: [512, 1600] read as OIHW [512][64][5][5]
                    d.Cout, d.Cin, d.R, d.S = c.cout, c.cin, c.k, c.k
                d.K, d.kmap, d.eps = c.K, c.kmap, BN_EPS
                d.dgrad_parity = (2 if c.name in fused_c1 else 1) if (c.kmap == 0 and c.stride == 2 and c.k == 3) else 0

            def upload(descs):
                return torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).clone().to(dev)

            # the tiled kernel holds 32 x 32 x (<= 9 taps) in shared memory; the stem (packing) and
            # top.0 (25 taps) go through the element-wise one
            elem = [c for c in convs if c.kmap != 0 or c.k * c.k > 9]
            tiled = [c for c in convs if c.kmap == 0 and c.k * c.k <= 9]
            d_elem = (L.WprepDesc * len(elem))()
            offs, total = [], 0
            for d, c in zip(d_elem, elem):
                fill(d, c)
                offs.append(total)
                total += c.cout * c.K
            d_tiled = (L.WprepDesc * len(tiled))()
            toffs, tiles = [], 0
            for d, c in zip(d_tiled, tiled):
                fill(d, c)
                toffs.append(tiles)
                tiles += (c.cout // 32) * (c.cin // 32)
            self._elem = (upload(d_elem), torch.tensor(offs, dtype=torch.int64, device=dev), len(elem), total)
            self._tiled = (upload(d_tiled), torch.tensor(toffs, dtype=torch.int32, device=dev), len(tiled), tiles)
            self._total, self._table_sig = total + sum(c.cout * c.K for c in tiled), sig
        lib = L.load()
        with ops._Prof("weight_prep", (self._total,)):
            t, o, n, tot = self._elem
            L.check(lib.vdqn_weight_prep_multi(t.data_ptr(), o.data_ptr(), n, tot, L.stream_ptr()),
                    "weight_prep_multi")
            t, o, n, tot = self._tiled
            L.check(lib.vdqn_weight_prep_tiled(t.data_ptr(), o.data_ptr(), n, tot, L.stream_ptr()),
                    "weight_prep_tiled")
        for key, (hi, lo) in self.mlp.items():
            ops.split_bf16(P[key + ".weight"], hi, lo, perm=(64, 25) if key == "top.0" else None)


# output-parity classes of a 3x3 stride-2 data gradient: (a, b, taps_h, taps_w, offset in taps)
PARITY_CLASSES = ((0, 0, 1, 1, 0), (0, 1, 1, 2, 1), (1, 0, 2, 1, 3), (1, 1, 2, 2, 5))


def parity_filters(W: "PreparedWeights", c: ConvSpec):
    """Views of the four parity-class data-gradient filters [Cin][na][nb][Cout] of a strided conv."""
    flat = W.w_dgrad[c.name].view(-1)
    unit = c.cin * c.cout
    return [(a, b, flat[off * unit:(off + na * nb) * unit].view(c.cin, na, nb, c.cout))
            for a, b, na, nb, off in PARITY_CLASSES]


class Workspace:
    """All activation / gradient buffers for one forward (and optionally its backward) at a
    fixed number of frames `n` (= B * num_frames)."""

    def __init__(self, plan: NetPlan, n: int, device, train: bool, n_bwd: Optional[int] = None,
                 keep_stem: bool = False):
        """`n` frames go through the forward; the backward (if `train`) covers the first `n_bwd`
        (default all) -- the fused step runs the online net on [s ; s'] in one 2B forward and
        back-propagates through the s half only."""
        self.n, self.train = n, train
        self.n_bwd = n if n_bwd is None else n_bwd
        self._view = None
        nf, n = n, self.n_bwd
        e = lambda *s, dt=bf16: torch.empty(*s, device=device, dtype=dt)  # noqa: E731
        z = lambda *s, dt=bf16: torch.zeros(*s, device=device, dtype=dt)  # noqa: E731
        self.xp = e(nf, 112, 112, 16)
        # the stem's ReLU output only exists in HBM when asked for (train-mode BatchNorm of the `basic`
        # architecture normalises it before pooling); otherwise the stem kernel pools in its epilogue
        self.s = e(nf, 112, 112, 64) if keep_stem else None
        # arg-max slots of the max-pool: only the frames that are back-propagated through need them
        self.idx = e(self.n_bwd, 56, 56, 64, dt=torch.uint8) if train else None
        self.p = e(nf, 56, 56, 64)
        self.a1, self.idn, self.out = [], [], []
        for b in plan.blocks:
            self.a1.append(e(nf, b.out_hw, b.out_hw, b.cout))
            self.idn.append(None)          # identity tensor of a strided block: only the unfused path makes one
            self.out.append(e(nf, b.out_hw, b.out_hw, b.cout))
        self.h = e(nf, 5, 5, 64)
        F = plan.num_frames
        Bf, B = nf // F, n // F
        f32 = torch.float32
        # Q-head MLP activations as hi + lo bf16 pairs (the operands of the next GEMM; hi > 0 is the ReLU mask)
        nq = plan.num_classes * plan.action_dim
        self.z1 = (e(Bf, 512), e(Bf, 512))
        self.z2 = (e(Bf, 256), e(Bf, 256))
        self.q = e(Bf, nq, dt=f32)
        self._fwd_names = ("xp", "s", "idx", "p", "h")        # `s` may be None
        self._mlp_names = ("q",)
        self._F = F
        if train:
            self.dq_hl = (z(B, (nq + 7) // 8 * 8), z(B, (nq + 7) // 8 * 8))
            self.dz2 = (e(B, 256), e(B, 256))
            self.dz1 = (e(B, 512), e(B, 512))
            self.dh = e(n, 5, 5, 64)
            # gradient wrt block outputs / conv1 outputs: two rotating buffers per resolution
            self.dy_out = {b.out_hw: (e(n, b.out_hw, b.out_hw, b.cout), e(n, b.out_hw, b.out_hw, b.cout))
                           for b in plan.blocks}
            self.dy_a1 = {b.out_hw: e(n, b.out_hw, b.out_hw, b.cout) for b in plan.blocks}
            # zero-dilated buffer the 1x1/2 downsample data gradient scatters into (odd positions stay
            # zero forever); it is the residual of the strided conv1 data gradient
            self.r_dil = {}                # (unfused path only, made on first use)
            self.dy_p = e(n, 56, 56, 64)
            self.dy_s = e(n, 112, 112, 64)
            max_part = max(wgrad_splits(c, n) * c.cout * c.K for c in prep_convs(plan))
            self.part = e(max_part, dt=f32)
            # deferred split reductions (set `defer_fin` to use): every weight gradient keeps its own
            # partial buffer and ONE multi-tensor launch reduces them all at the end of the backward pass
            # (or of each data-parallel stage)
            self.defer_fin = False
            self._part_of: Dict[str, torch.Tensor] = {}
            self._pending: List[tuple] = []
            self._fin_tables: Dict[tuple, tuple] = {}
            self._side = None

    def side_stream(self):
        """the stream the weight-gradient kernels run on (engine.SideStream), or None when that is switched off"""
        if not WGRAD_SIDE:
            return None
        if getattr(self, "_side", None) is None:
            self._side = SideStream(self.part.device)
        return self._side

    def bwd_view(self):
        """The workspace as the backward pass sees it: activations sliced to the first n_bwd frames."""
        if self.n_bwd == self.n:
            return self
        if self._view is None:
            import copy
            v = copy.copy(self)
            nb, Bb = self.n_bwd, self.n_bwd // self._F
            for nm in self._fwd_names:
                t = getattr(self, nm)
                setattr(v, nm, None if t is None else t[:nb])
            for nm in self._mlp_names:
                setattr(v, nm, getattr(self, nm)[:Bb])
            v.z1 = tuple(t[:Bb] for t in self.z1)
            v.z2 = tuple(t[:Bb] for t in self.z2)
            v.a1 = [t[:nb] for t in self.a1]
            v.out = [t[:nb] for t in self.out]
            v.idn = [None if t is None else t[:nb] for t in self.idn]
            v.n = nb
            self._view = v
        return self._view


import os as _os

# 1x1/2 downsample of a strided block inside conv2 (forward) / conv1's merged data gradient (backward); 0: its
# own launches and four parity-class launches for the strided data gradient (the A/B switch for measurements)
FUSE_DS = _os.environ.get("VDQN_FUSE_DS", "1") != "0"
# stem + max-pool fused into one kernel (0: the stem writes its full-resolution output and a pooling kernel
# re-reads it -- the A/B switch for measurements)
FUSE_POOL = _os.environ.get("VDQN_FUSE_POOL", "1") != "0"
# column-tile width for the Cout >= 256 layers (0 = kernel default 256); tuning knob
TILE_N_WIDE = int(_os.environ.get("VDQN_TILE_N_WIDE", "0"))
# fused step: one multi-tensor launch for all split reductions of the backward pass (0: one per layer)
DEFER_FINALIZE = _os.environ.get("VDQN_DEFER_FINALIZE", "1") != "0"
# column-tile width of the layer4 data gradients (196 tiles of 128x256 on 148 SMs = two uneven waves
# at B = 256; 128 gives 392 smaller tiles)
TILE_N_DGRAD4 = int(_os.environ.get("VDQN_TILE_N_DGRAD4", "0"))


def _scatter_tile_n(ch: int) -> int:
    """zero-dilated (stride-2) data gradients: column tiles of at most 128 keep the staged epilogue"""
    return 128 if ch >= 256 else 0


# column-tile width of the stride-1 data gradients with >= 256 channels (layers 3-4): 128 selects the
# staged epilogue (TMA-moved mask / residual / output tiles, specialised loops); with 256-wide tiles
# the direct epilogue (per-lane global loads of the mask, a warp transpose per 32 columns for the column
# sums) was the longest role -- 23.7 k cycles per tile against 18.4 k of MMA issue (role profile)
TILE_N_DGRAD = int(_os.environ.get("VDQN_TILE_N_DGRAD", "128"))


def _dgrad_tile_n(ch: int) -> int:
    if ch >= 512 and TILE_N_DGRAD4:
        return TILE_N_DGRAD4
    if ch >= 256 and TILE_N_DGRAD:
        return TILE_N_DGRAD
    return TILE_N_WIDE if ch >= 256 else 0


def _conv(W: PreparedWeights, c: ConvSpec, x, out, W2: Optional[PreparedWeights] = None, split: int = 0, **kw):
    tn = TILE_N_WIDE if c.cout >= 256 else 0
    if W2 is not None:
        kw.update(w2=W2.w_fwd[c.name], shift2=W2.shift[c.name], split_n=split)
    w = W.w_fwd[c.name]
    if w.dim() == 2:
        kw["ksize"] = (c.k, c.k)
    return ops.conv_gemm(x, w, c.stride, c.pad_lo, c.pad_hi, shift=W.shift[c.name], out=out, tile_n=tn, **kw)


def forward(plan: NetPlan, W: PreparedWeights, P: Dict[str, torch.Tensor], ws: Workspace,
            frames: torch.Tensor) -> torch.Tensor:
    """frames: [n,3,224,224] fp32 NCHW (or [n,224,224,3] uint8) -> Q [B, classes*actions] fp32."""
    ops.stem_pack(frames, ws.xp)
    return forward_packed(plan, W, P, ws)


def forward_packed(plan: NetPlan, W: PreparedWeights, P: Dict[str, torch.Tensor], ws: Workspace,
                   W2: Optional[PreparedWeights] = None, P2: Optional[Dict[str, torch.Tensor]] = None,
                   split: int = 0, trunk_only: bool = False, x_alias=None) -> torch.Tensor:
    """Forward from the packed input ws.xp.  With (W2, P2, split) the frames [split, n) go through a
    SECOND network (the target net) inside the same launches: every conv kernel partitions its CTAs
    between the two image ranges, so online and target forwards share one pass (better SM filling
    for the small late layers, a third fewer launches)."""
    dual = dict(W2=W2, split=split) if W2 is not None else {}
    # x_alias = (first, shift): packed frames >= first are read from frame - shift (the target range of
    # the 3B pass re-reads the online network's s' frames)
    nb = ws.n_bwd if ws.idx is not None else 0
    if ws.s is None and FUSE_POOL:
        # stem conv + bn1 + ReLU + max-pool in ONE kernel: only the pooled tensor (and, for the frames that are
        # back-propagated through, the arg-max slots) reaches HBM
        _conv(W, plan.stem, ws.xp, None, relu=True, x_alias=x_alias, pool_out=ws.p, pool_idx=ws.idx,
              pool_idx_images=nb, **dual)
    else:
        if ws.s is None:
            ws.s = torch.empty(ws.n, 112, 112, 64, device=ws.xp.device, dtype=bf16)
        _conv(W, plan.stem, ws.xp, ws.s, relu=True, x_alias=x_alias, **dual)
        if nb:
            ops.maxpool_fwd(ws.s[:nb], ws.p[:nb], ws.idx)
        if nb < ws.n:
            ops.maxpool_fwd(ws.s[nb:], ws.p[nb:], None)
    x = ws.p
    for i, b in enumerate(plan.blocks):
        _conv(W, b.conv1, x, ws.a1[i], relu=True, **dual)
        if b.ds is not None and W.fuse_ds:
            # downsample accumulated inside conv2's kernel: no identity tensor, no residual read
            _conv(W, b.conv2, ws.a1[i], ws.out[i], x2=x, stride2=b.stride, relu=True, **dual)
        else:
            idn = x
            if b.ds is not None:
                if ws.idn[i] is None:
                    ws.idn[i] = torch.empty_like(ws.out[i])
                _conv(W, b.ds, x, ws.idn[i], **dual)
                idn = ws.idn[i]
            _conv(W, b.conv2, ws.a1[i], ws.out[i], residual=idn, relu=True, **dual)
        x = ws.out[i]
    if trunk_only:
        return x                                      # [n, 7, 7, 512] bf16
    _conv(W, plan.head, x, ws.h, relu=True, **dual)
    mlp_forward(plan, W, P, ws, W2, P2, split)
    return ws.q


def mlp_forward(plan: NetPlan, W: PreparedWeights, P, ws: Workspace, W2=None, P2=None, split: int = 0):
    """top.0 / top.2 / top.4 (archs/HabitatDQNMultiAction.py:31,53) as three tensor-core GEMMs over split-bf16
    operands (csrc/mlp_gemm.cu): the head-conv output [frames, 5, 5, 64] is read in place as [B, 1600 F]
    (top.0's columns were permuted to that order when its operands were prepared), z1 / z2 only exist as the
    hi + lo operand pairs of the next GEMM, Q is fp32.  With a second network, rows >= split // F use its
    weights inside the same launches when that row is a multiple of 128, otherwise in a second set of launches."""
    F = plan.num_frames
    rows = ws.h.shape[0] // F
    hA = ws.h.view(rows, 1600 * F)
    nq = plan.num_classes * plan.action_dim
    bs = split // F if W2 is not None else 0

    def run(r0, r1, Wn, Pn, Wd=None, Pd=None, split_m=0):
        m = r1 - r0
        h, z1, z2, q = hA[r0:r1], tuple(t[r0:r1] for t in ws.z1), tuple(t[r0:r1] for t in ws.z2), ws.q[r0:r1]
        w0, w2, w4 = Wn.mlp["top.0"], Wn.mlp["top.2"], Wn.mlp["top.4"]
        d0, d2, d4 = (Wd.mlp[k] for k in ("top.0", "top.2", "top.4")) if Wd is not None else (None, None, None)
        kw = lambda k: dict(split_m=split_m, bias2=Pd[k + ".bias"]) if Wd is not None else {}  # noqa: E731
        ops.mlp_gemm([(h, w0[0], 1600 * F), (h, w0[1], 1600 * F)], m, 512, bias=Pn["top.0.bias"], relu=True,
                     segments2=[(h, d0[0], 0), (h, d0[1], 0)] if Wd is not None else None,
                     out_hi=z1[0], out_lo=z1[1], **kw("top.0"))
        ops.mlp_gemm([(z1[0], w2[0], 512), (z1[1], w2[0], 512), (z1[0], w2[1], 512)], m, 256, bias=Pn["top.2.bias"],
                     relu=True, out_hi=z2[0], out_lo=z2[1],
                     segments2=[(None, d2[0], 0), (None, d2[0], 0), (None, d2[1], 0)] if Wd is not None else None,
                     **kw("top.2"))
        ops.mlp_gemm([(z2[0], w4[0], 256), (z2[1], w4[0], 256), (z2[0], w4[1], 256)], m, nq, bias=Pn["top.4.bias"],
                     out_f32=q,
                     segments2=[(None, d4[0], 0), (None, d4[0], 0), (None, d4[1], 0)] if Wd is not None else None,
                     **kw("top.4"))
    if W2 is None:
        run(0, rows, W, P)
    elif bs % 128 == 0:
        run(0, rows, W, P, W2, P2, split_m=bs)
    else:
        run(0, bs, W, P)
        run(bs, rows, W2, P2)


def mlp_backward(plan: NetPlan, W: PreparedWeights, P, G, ws: Workspace, dq: torch.Tensor):
    """Backward of the Q-head MLP on the same kernel: per layer the data gradient dy W (W read MN-major, ReLU
    mask and bias-gradient column sums in the epilogue, result kept as a hi + lo pair) and the weight gradient
    dy^T x (both operands MN-major, fp32 result straight into the gradient arena; d top.0.weight is un-permuted
    on the way out).  The last data gradient is the head conv's dy: bf16 NHWC, masked by the head ReLU, with the
    head bias gradient as its column sums.  `ws` is the backward view (first n_bwd frames)."""
    F = plan.num_frames
    B = ws.h.shape[0] // F
    hA = ws.h.view(B, 1600 * F)
    nq = plan.num_classes * plan.action_dim
    w0, w2, w4 = W.mlp["top.0"], W.mlp["top.2"], W.mlp["top.4"]
    dqh, dql = ws.dq_hl
    z1, z2, dz1, dz2 = ws.z1, ws.z2, ws.dz1, ws.dz2
    ops.split_bf16(dq.view(B, nq), dqh, dql, colsum=G["top.4.bias"])
    # data-gradient chain: dq -> dz2 -> dz1 (each needs the previous one)
    ops.mlp_gemm([(dqh, w4[0], nq), (dql, w4[0], nq), (dqh, w4[1], nq)], B, 256, b_mn=True, mask_bf16=z2[0],
                 out_hi=dz2[0], out_lo=dz2[1], colsum=G["top.2.bias"])
    ops.mlp_gemm([(dz2[0], w2[0], 256), (dz2[1], w2[0], 256), (dz2[0], w2[1], 256)], B, 512, b_mn=True,
                 mask_bf16=z1[0], out_hi=dz1[0], out_lo=dz1[1], colsum=G["top.0.bias"])
    # the three weight gradients and the head conv's dy are independent of each other.  Only the head's dy is on
    # the path to everything else: the weight gradients go to the second stream (one grouped launch), or, without
    # it, into the same launch as the data gradient
    p_dh = ops.mlp_problem([(dz1[0], w0[0], 512), (dz1[1], w0[0], 512), (dz1[0], w0[1], 512)], B, 1600 * F, b_mn=True,
                           mask_bf16=hA, out_bf16=ws.dh.view(B, 1600 * F), colsum=G["features.8.bias"], colsum_mod=64)
    p_dw = [
        ops.mlp_problem([(dz1[0], hA, B), (dz1[1], hA, B)], 512, 1600 * F, a_mn=True, b_mn=True,
                        out_f32=G["top.0.weight"], perm=(64, 25)),
        ops.mlp_problem([(dz2[0], z1[0], B), (dz2[1], z1[0], B), (dz2[0], z1[1], B)], 256, 512, a_mn=True, b_mn=True,
                        out_f32=G["top.2.weight"]),
        ops.mlp_problem([(dqh, z2[0], B), (dql, z2[0], B), (dqh, z2[1], B)], nq, 256, a_mn=True, b_mn=True,
                        out_f32=G["top.4.weight"])]
    side = ws.side_stream() if hasattr(ws, "side_stream") and getattr(ws, "part", None) is not None else None
    if side is not None and WGRAD_SIDE and ops.PROFILE is None:
        side.run(lambda: ops.mlp_gemm_grouped(p_dw))
        ops.mlp_gemm_grouped([p_dh])
    else:
        ops.mlp_gemm_grouped([p_dw[0], p_dh, p_dw[1], p_dw[2]])


# Weight-gradient kernels on a second stream.  A conv's weight gradient and its data gradient both read the same
# dy and nothing else of the backward pass depends on the weight gradient, so the two chains -- data gradients
# on the main stream, weight gradients on the side stream -- only meet at the split reduction.  Every kernel is
# persistent with one CTA per SM, so two of them never share an SM; what the second stream buys is that the CTAs
# of the next runnable kernel take over each SM the moment a CTA of the running one exits, instead of all SMs
# waiting for the slowest CTA at every kernel boundary (data gradients of layers 3-4 at B = 256 are 5.3 and
# 2.65 rounds of tiles per CTA pair: 12 % of their time is such a tail) and then for the next launch.
WGRAD_SIDE = _os.environ.get("VDQN_WGRAD_SIDE", "1") != "0"


class SideStream:
    """Runs closures on a second stream after everything enqueued on the current one, and keeps, per buffer, the
    event of the last side kernel that reads it so that the main stream can wait before overwriting it."""

    def __init__(self, device, priority: int = 0):
        self.stream = torch.cuda.Stream(device=device, priority=priority)
        self._readers: Dict[int, torch.cuda.Event] = {}
        self._last = None

    def run(self, fn, reads=()):
        if not WGRAD_SIDE or ops.PROFILE is not None:       # (per-kernel timing wants the kernels one after another)
            fn()
            return
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            fn()
            done = torch.cuda.Event()
            done.record(self.stream)
        for t in reads:
            self._readers[t.data_ptr()] = done
        self._last = done

    def before_write(self, t):
        ev = self._readers.pop(t.data_ptr(), None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def join(self):
        if self._last is not None:
            torch.cuda.current_stream().wait_event(self._last)
            self._last = None
        self._readers.clear()


def _wgrad(plan, P, G, ws: Workspace, c: ConvSpec, x, dy, dbeta_key: Optional[str]):
    splits = wgrad_splits(c, ws.n)
    defer = ws.defer_fin and c.kmap == 0
    if defer:
        part = ws._part_of.get(c.name)
        if part is None:
            part = ws._part_of[c.name] = torch.empty(splits * c.cout * c.K, device=ws.part.device,
                                                     dtype=torch.float32)
    else:
        part = ws.part[: splits * c.cout * c.K]
    kw = {}
    if c.bn is not None:
        kw = dict(gamma=P[c.bn + ".weight"], var=P[c.bn + ".running_var"], mean=P[c.bn + ".running_mean"],
                  dbeta=G[c.bn + ".bias"], dgamma=G[c.bn + ".weight"])
    args = dict(splits=splits, Cout=c.cout, Cin=c.gemm_cin, R=c.k, S=c.k, K=c.K, kmap=c.kmap, eps=BN_EPS, **kw)

    def launch():
        ops.conv_wgrad(x, dy, c.k, c.k, c.stride, c.pad_lo, c.pad_hi, splits=splits, part=part,
                       algo=2 if wgrad_uses_halo(c) else 0)
        if not defer:
            # (the shared partial buffer and the reduction stay in the order of the stream the kernels run on)
            ops.wgrad_finalize(part, P[c.wkey], G[c.wkey], **args)

    side = ws.side_stream()
    if side is not None:
        side.run(launch, reads=(dy,))
    else:
        launch()
    if defer:
        ws._pending.append((c.name, part, P[c.wkey], G[c.wkey], args))


def _flush_finalize(ws: Workspace):
    """reduce the split partials of every weight gradient enqueued since the last flush: one launch"""
    side = ws.side_stream()
    if side is not None:
        side.join()                      # every weight gradient enqueued so far (and its readers) is accounted for
    if not ws._pending:
        return
    key = tuple((nm, w.data_ptr(), dw.data_ptr(), args["splits"]) for nm, _, w, dw, args in ws._pending)
    tab = ws._fin_tables.get(key)
    if tab is None:
        descs = [ops.wgrad_finalize_desc(part, w, dw, **args) for _, part, w, dw, args in ws._pending]
        tab = ws._fin_tables[key] = ops.wgrad_finalize_table(descs, ws.part.device)
    ops.wgrad_finalize_multi(tab)
    ws._pending.clear()


def backward(plan: NetPlan, W: PreparedWeights, P: Dict[str, torch.Tensor], G: Dict[str, torch.Tensor],
             ws: Workspace, dq: torch.Tensor, on_grads_ready=None):
    """dq: [B, classes*actions] fp32 (overwritten).  Writes every parameter gradient into G
    (views of the flat gradient arena; BN-bias slots must be zero on entry: they are accumulated
    with atomics).  `on_grads_ready(stage)` is called after each stage's gradients are enqueued
    (used by the data-parallel wrapper to launch bucketed all-reduces)."""
    ws = ws.bwd_view()
    if on_grads_ready is None:
        notify = lambda stage: None  # noqa: E731
    else:
        wants = getattr(getattr(on_grads_ready, "__self__", None), "wants", None)

        def notify(stage):
            if wants is not None and not wants(stage) and stage != "stem":
                return                   # this exchange does not act on the stage: no reduction needed yet
            _flush_finalize(ws)          # the stage's gradients are complete only once its reductions ran
            on_grads_ready(stage)
    side = ws.side_stream()

    def guard(t):
        """the main stream is about to overwrite `t`: wait for the weight-gradient kernels still reading it"""
        if side is not None:
            side.before_write(t)
        return t
    # ---- Q-head MLP + the head conv's dy (ws.dh)
    mlp_backward(plan, W, P, G, ws, dq)
    last = plan.blocks[-1]
    x_head = ws.out[-1]
    _wgrad(plan, P, G, ws, plan.head, x_head, ws.dh, None)
    ci = 0
    cur = ws.dy_out[last.out_hw][ci]
    ops.conv_gemm(ws.dh, W.w_dgrad["head"], 1, 2, 2, mask_src=x_head, colsum=G[last.conv2.bn + ".bias"],
                  out=guard(cur))
    notify("head")
    # ---- residual blocks, last to first.  `cur` = d loss / d (block output), already masked by
    # the block's final ReLU; its column sums are already accumulated in G[bn2.bias].
    for i in range(len(plan.blocks) - 1, -1, -1):
        b = plan.blocks[i]
        x_in = ws.out[i - 1] if i > 0 else ws.p
        prev = plan.blocks[i - 1] if i > 0 else None
        if b.ds is not None:
            # the downsample BN sees the same upstream gradient as bn2
            G[b.ds.bn + ".bias"].copy_(G[b.conv2.bn + ".bias"])
        # conv2: weight gradient, then data gradient masked by relu(bn1(conv1))
        _wgrad(plan, P, G, ws, b.conv2, ws.a1[i], cur, None)
        dy_a1 = guard(ws.dy_a1[b.out_hw])
        ops.conv_gemm(cur, W.w_dgrad[b.conv2.name], 1, 1, 1, mask_src=ws.a1[i],
                      colsum=G[b.conv1.bn + ".bias"], out=dy_a1,
                      tile_n=_dgrad_tile_n(b.cout))
        # identity / downsample branch
        if b.ds is not None:
            _wgrad(plan, P, G, ws, b.ds, x_in, cur, None)
            if W.fuse_ds:
                res = None               # the downsample's data gradient is a K segment of conv1's (below)
            else:
                if b.out_hw not in ws.r_dil:
                    ws.r_dil[b.out_hw] = torch.zeros(ws.n, b.in_hw, b.in_hw, b.cin, device=cur.device, dtype=bf16)
                res = ws.r_dil[b.out_hw]
                ops.conv_gemm(cur, W.w_dgrad[b.ds.name], 1, 0, 0, out=res, out_scatter=2, tile_n=_scatter_tile_n(b.cin))
        else:
            res = cur
        # conv1: weight gradient, then the gradient wrt the block input (+ identity branch),
        # masked by the previous block's final ReLU
        _wgrad(plan, P, G, ws, b.conv1, x_in, dy_a1, None)
        if prev is not None:
            ni = (1 - ci) if prev.out_hw == b.out_hw else 0
            dst = ws.dy_out[prev.out_hw][ni]
            colsum, mask = G[prev.conv2.bn + ".bias"], x_in
        else:
            ni, dst, colsum, mask = 0, ws.dy_p, None, None
        guard(dst)
        if b.stride == 2 and W.fuse_ds:
            # strided data gradient + downsample data gradient as ONE dense GEMM: 2x2 window over dy_a1 with the
            # four output-parity classes as column groups (taps a class does not use are zero columns: 16/9 of
            # the minimal MMA work, but one launch with 5 Cout of K per tile instead of four latency-bound ones
            # with Cout .. 4 Cout), plus the K segment over `cur` for the downsample; every row writes its 2x2
            # pixel block of dX
            ops.conv_gemm(dy_a1, W.w_dgrad[b.conv1.name], 1, 0, 1, ksize=(2, 2), x2=cur, stride2=1, mask_src=mask,
                          colsum=colsum, out=dst, out_scatter=3, scatter_inputs=True, tile_n=128)
        elif b.stride == 2:
            # strided data gradient, one small stride-1 conv per output-parity class (h%2, w%2): exactly
            # the 9 taps of work, each class scattering into its own pixels of dX
            for pa, pb, wf in parity_filters(W, b.conv1):
                ops.conv_gemm(dy_a1, wf, 1, 0, wf.shape[1] - 1, pad_hi_w=wf.shape[2] - 1, residual=res,
                              mask_src=mask, colsum=colsum, out=dst, out_scatter=2, scatter_off=(pa, pb),
                              scatter_inputs=True, tile_n=_scatter_tile_n(b.cin))
        else:
            ops.conv_gemm(dy_a1, W.w_dgrad[b.conv1.name], 1, 1, 1, residual=res, mask_src=mask,
                          colsum=colsum, out=dst, tile_n=_dgrad_tile_n(b.cin))
        notify(b.conv1.name)
        cur, ci = dst, ni
    # ---- max-pool + stem
    ops.maxpool_bwd(ws.dy_p, ws.idx, ws.p, guard(ws.dy_s), colsum=G[plan.stem.bn + ".bias"])
    _wgrad(plan, P, G, ws, plan.stem, ws.xp, ws.dy_s, None)
    _flush_finalize(ws)
    notify("stem")
