"""Forward pass of the inverse-dynamics model: the network that writes the
`inverse_actions` column of the quadruplet table, `model(be, ae)[1].argmax(dim=1)`
(`dataset/process_episodes_real.py:92-95,171-179`; architecture `archs/inverse_action2.py:45-100`).

    trunk(k), trunk(k+1)  (frozen ResNet-18 children()[:-2], eval BN)  -> cat on channels (1024)
    conv1 1x1 1024->256 + ReLU, conv2 3x3 256->256 + ReLU, conv3 3x3 256->64 + ReLU  (no padding)
    flatten (NCHW order) -> fc1 576->128 + ReLU -> [dropout: identity in eval] -> fc2 128->3
    encoding = softmax(fc2), y = fc_accuracy(fc2)

Everything runs on the kernels of the Q-learning path: both trunks as ONE 2B forward of the conv
engine, the channel concatenation folded away (conv1 = W[:, :512] * trunk(k) + W[:, 512:] *
trunk(k+1): the second GEMM takes the first as its residual), the three head convs on the im2col
tensor-core kernel, the fully connected layers on the fp32 kernels.  Weights are bf16 operands with
fp32 accumulation; the labelling script only consumes the arg-max.  Inference only (`model.eval()`,
`:95`); CUDA only, no fallback.
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch

from . import engine as E
from . import ops

_SEQ = {"0": "conv1", "1": "bn1", "4": "layer1", "5": "layer2", "6": "layer3", "7": "layer4"}


def _trunk_params(sd: Dict[str, torch.Tensor], device) -> Dict[str, torch.Tensor]:
    """`resnet18.<i>.<rest>` (positional nn.Sequential keys) -> the `resnet.<name>.<rest>` naming of
    the conv engine"""
    P = {}
    for k, v in sd.items():
        if k.startswith("resnet18."):
            idx, _, rest = k[len("resnet18."):].partition(".")
            if idx in _SEQ and not rest.endswith("num_batches_tracked"):
                P[f"resnet.{_SEQ[idx]}.{rest}"] = v.detach().to(device=device, dtype=torch.float32).contiguous()
    return P


class InverseActionModule(torch.nn.Module):
    """Parameter container with the reference module's layout and default initialisation
    (`train_inverse_model.py:30-53`, `archs/inverse_action2.py:45-70`): `resnet18` = positional
    `nn.Sequential(children()[:-2])`, frozen (`requires_grad=False`, eval), then `conv1/conv2/conv3`, `fc1`,
    `fc2` (bottleneck 3), `fc_accuracy`.  Its `state_dict()` is what `InverseActionRunner` /
    `InverseModelTrainer` take and what `model-N.pth` files hold; there is no CPU forward."""

    def __init__(self, bottleneck_size: int = 3):
        super().__init__()
        from .qnet import _make_resnet18
        nn = torch.nn
        self.resnet18 = nn.Sequential(*list(_make_resnet18(True).children())[:-2])
        self.resnet18.eval()
        for p in self.resnet18.parameters():
            p.requires_grad = False
        self.conv1 = nn.Conv2d(1024, 256, kernel_size=1)
        self.conv2 = nn.Conv2d(256, 256, kernel_size=3)
        self.conv3 = nn.Conv2d(256, 64, kernel_size=3)
        self.dropout1 = nn.Dropout2d(0.5)
        self.fc1 = nn.Linear(64 * 3 * 3, 128)
        self.fc2 = nn.Linear(128, bottleneck_size)
        self.fc_accuracy = nn.Linear(bottleneck_size, 3)

    def forward(self, k, k_plus_one):
        raise RuntimeError("InverseActionModule is a parameter container (no CPU path): run it through "
                           "InverseActionRunner / InverseModelTrainer on a B200")


class InverseActionRunner:
    """`runner(k, k_plus_one)` -> (encoding [B,3], y [B,3]); `runner.label(k, k1)` -> actions [B].
    Frames: fp32 NCHW normalised (the reference's loader output) or uint8 HWC."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], batch_size: int, device=None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("InverseActionRunner needs a CUDA device (no CPU path)")
        self.B, self.dev = batch_size, dev
        self.plan = E.make_plan(3, 5)
        self.P = _trunk_params(state_dict, dev)
        self.W = E.PreparedWeights(self.plan, dev, trunk_only=True)
        self.W.prepare(self.P)
        self.ws = E.Workspace(self.plan, 2 * batch_size, dev, train=False)
        f32 = lambda k: state_dict[k].detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        bf = torch.bfloat16

        def prep(w, bias):
            cout, cin, r, s = w.shape
            wf = torch.empty(cout, r, s, cin, device=dev, dtype=bf)
            shift = torch.empty(cout, device=dev, dtype=torch.float32)
            ops.weight_prep(w, wf, shift, bias=bias)
            return wf, shift
        w1 = f32("conv1.weight")
        zero = torch.zeros(256, device=dev)
        self.w1a, _ = prep(w1[:, :512].contiguous(), zero)
        self.w1b, self.b1 = prep(w1[:, 512:].contiguous(), f32("conv1.bias"))
        self.w2, self.b2 = prep(f32("conv2.weight"), f32("conv2.bias"))
        self.w3, self.b3 = prep(f32("conv3.weight"), f32("conv3.bias"))
        self.fc = {k: f32(k) for k in ("fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias",
                                       "fc_accuracy.weight", "fc_accuracy.bias")}
        B = batch_size
        e = lambda *s, dt=bf: torch.empty(*s, device=dev, dtype=dt)  # noqa: E731
        self.t1, self.x1 = e(B, 7, 7, 256), e(B, 7, 7, 256)
        self.x2, self.x3 = e(B, 5, 5, 256), e(B, 3, 3, 64)
        self.flat, self.h1 = e(B, 576, dt=torch.float32), e(B, 128, dt=torch.float32)
        self.z, self.y = e(B, 3, dt=torch.float32), e(B, 3, dt=torch.float32)

    def __call__(self, k: torch.Tensor, k_plus_one: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        B = self.B
        if k.shape[0] != B or k_plus_one.shape != k.shape:
            raise ValueError("bad shape")
        if not k.is_cuda:
            raise RuntimeError("frames must be CUDA tensors (no CPU path)")
        ws = self.ws
        ops.stem_pack(k.contiguous(), ws.xp[:B])
        ops.stem_pack(k_plus_one.contiguous(), ws.xp[B:])
        feat = E.forward_packed(self.plan, self.W, self.P, ws, trunk_only=True)     # [2B,7,7,512]
        ops.conv_gemm(feat[:B], self.w1a, 1, 0, 0, out=self.t1)
        ops.conv_gemm(feat[B:], self.w1b, 1, 0, 0, shift=self.b1, residual=self.t1, relu=True, out=self.x1)
        ops.conv_gemm(self.x1, self.w2, 1, 0, 0, shift=self.b2, relu=True, out=self.x2)
        ops.conv_gemm(self.x2, self.w3, 1, 0, 0, shift=self.b3, relu=True, out=self.x3)
        ops.head_flatten_fwd(self.x3, self.flat)                                     # NCHW order: c*9 + p
        fc = self.fc
        ops.linear_fwd(self.flat, fc["fc1.weight"], fc["fc1.bias"], True, self.h1)
        ops.linear_fwd(self.h1, fc["fc2.weight"], fc["fc2.bias"], False, self.z)
        ops.linear_fwd(self.z, fc["fc_accuracy.weight"], fc["fc_accuracy.bias"], False, self.y)
        return torch.softmax(self.z, dim=1), self.y

    def label(self, k: torch.Tensor, k_plus_one: torch.Tensor) -> torch.Tensor:
        """`model(be, ae)[1].argmax(dim=1)` (dataset/process_episodes_real.py:176-177)"""
        return self(k, k_plus_one)[1].argmax(dim=1)


# ------------------------------------------------------------------------------------------------
# training step (train_inverse_model.py)
# ------------------------------------------------------------------------------------------------
# arena order; `conv1.weight` is kept as its two 512-channel halves (one per trunk: the channel
# concatenation is folded into two accumulating GEMMs, so each half has its own weight gradient)
_TRAIN_SLOTS = (("conv1.weight.k", (256, 512, 1, 1)), ("conv1.weight.k1", (256, 512, 1, 1)), ("conv1.bias", (256,)),
                ("conv2.weight", (256, 256, 3, 3)), ("conv2.bias", (256,)),
                ("conv3.weight", (64, 256, 3, 3)), ("conv3.bias", (64,)),
                ("fc1.weight", (128, 576)), ("fc1.bias", (128,)), ("fc2.weight", (3, 128)), ("fc2.bias", (3,)),
                ("fc_accuracy.weight", (3, 3)), ("fc_accuracy.bias", (3,)))
DROPOUT_P = 0.5                                   # nn.Dropout2d(0.5), train_inverse_model.py:48


class InverseModelTrainer:
    """The loop body of `train()` (train_inverse_model.py:93-110) and the forward of `validate()`
    (:147-160) for the trainer's own network (:30-82 -- NOT the arch file's forward: ReLU after fc2,
    only `y = fc_accuracy(.)` returned, element dropout on the fc1 output in train mode):

        optimizer.zero_grad(); y = model(be, ae); loss = CrossEntropyLoss()(y, act)
        loss.backward(); optimizer.step()          # Adam(lr, weight_decay=0), :176

    `step(k, k_plus_one, act)` -> loss (1-element device tensor; `correct` holds the number of correct
    arg-max predictions of the step, :105-106).  The frozen ResNet-18 trunk (eval-mode BN, no gradients,
    :38-42,58) runs both frames as one 2B forward of the conv engine; the head's forward, data
    gradients and weight gradients run on the tensor-core kernels of the Q-learning path (bf16
    operands re-derived from the fp32 masters every step, fp32 accumulation), the fully connected
    layers, dropout, cross-entropy and Adam in fp32.  `lr` may be changed between steps (the
    reference's per-epoch `StepLR`, :179,185).  `state_dict()` returns the reference module's keys
    (`model-N.pth`, :131-132).  CUDA only, no fallback."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], batch_size: int, lr: float = 1e-4,
                 weight_decay: float = 0.0, seed: int = 0, device=None, use_graph: bool = True,
                 frames_uint8: bool = False):
        """frames_uint8: frames arrive as uint8 HWC [B,224,224,3] (decoder output; the ImageNet
        normalisation of the reference's loader is then fused into the first kernel, 4x less H2D)
        instead of the loader's normalised fp32 NCHW [B,3,224,224]."""
        from .optim import FlatArena
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("InverseModelTrainer needs a CUDA device (no CPU path)")
        if weight_decay != 0.0:
            raise NotImplementedError("weight_decay != 0 (the reference's default, and only documented use, is 0)")
        B = self.B = batch_size
        self.dev, self.lr, self.seed = dev, float(lr), int(seed)
        self.use_graph = use_graph
        self.plan = E.make_plan(3, 5)
        self._trunk_sd = {k: v.detach().clone() for k, v in state_dict.items() if k.startswith("resnet18.")}
        self.P = _trunk_params(state_dict, dev)
        self.W = E.PreparedWeights(self.plan, dev, trunk_only=True)
        self.W.prepare(self.P)
        self.ws = E.Workspace(self.plan, 2 * B, dev, train=False)
        shapes = [torch.Size(s) for _, s in _TRAIN_SLOTS]
        self._p, self._g = FlatArena(shapes, dev), FlatArena(shapes, dev)
        self._m, self._v = FlatArena(shapes, dev), FlatArena(shapes, dev)
        names = [n for n, _ in _TRAIN_SLOTS]
        self.p = dict(zip(names, self._p.views()))
        self.g = dict(zip(names, self._g.views()))
        self.load_state_dict(state_dict)
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda *s, dt=bf: torch.empty(*s, device=dev, dtype=dt)  # noqa: E731
        # bf16 GEMM operands of the head convs (forward [Cout][R][S][Cin], data gradient [Cin][R][S][Cout])
        self.wf = {"conv1.weight.k": e(256, 1, 1, 512), "conv1.weight.k1": e(256, 1, 1, 512),
                   "conv2.weight": e(256, 3, 3, 256), "conv3.weight": e(64, 3, 3, 256)}
        self.wd = {"conv2.weight": e(256, 3, 3, 256), "conv3.weight": e(256, 3, 3, 64)}
        self.shift = {"conv1.weight.k": e(256, dt=f32), "conv1.weight.k1": e(256, dt=f32),
                      "conv2.weight": e(256, dt=f32), "conv3.weight": e(64, dt=f32)}
        self._zero_bias = torch.zeros(256, device=dev, dtype=f32)
        # conv geometry for the split weight gradients
        mk = lambda nm, cin, cout, k, ihw, ohw: E.ConvSpec(nm, nm, None, None, cin, cout, k, 1, 0, 0, ihw, ohw,  # noqa: E731
                                                           gemm_cin=cin)
        self.spec = {"conv1.weight.k": mk("inv.c1k", 512, 256, 1, 7, 7), "conv1.weight.k1": mk("inv.c1k1", 512, 256, 1, 7, 7),
                     "conv2.weight": mk("inv.c2", 256, 256, 3, 7, 5), "conv3.weight": mk("inv.c3", 256, 64, 3, 5, 3)}
        self.part = {n: e(E.wgrad_splits(c, B) * c.cout * c.K, dt=f32) for n, c in self.spec.items()}
        # activations / gradients
        self.t1, self.x1 = e(B, 7, 7, 256), e(B, 7, 7, 256)
        self.x2, self.x3 = e(B, 5, 5, 256), e(B, 3, 3, 64)
        self.flat, self.h1, self.hd = e(B, 576, dt=f32), e(B, 128, dt=f32), e(B, 128, dt=f32)
        self.z, self.y = e(B, 3, dt=f32), e(B, 3, dt=f32)
        self.keep = torch.ones(B, 128, device=dev, dtype=torch.uint8)
        self.dy, self.dz, self.dhd = e(B, 3, dt=f32), e(B, 3, dt=f32), e(B, 128, dt=f32)
        self.dflat = e(B, 576, dt=f32)
        self.dx3, self.dx2, self.dx1 = e(B, 3, 3, 64), e(B, 5, 5, 256), e(B, 7, 7, 256)
        # static inputs (graph replay reads these)
        fshape, fdt = ((B, 224, 224, 3), torch.uint8) if frames_uint8 else ((B, 3, 224, 224), f32)
        self.k = torch.zeros(fshape, device=dev, dtype=fdt)
        self.k1 = torch.zeros(fshape, device=dev, dtype=fdt)
        self.act = torch.zeros(B, device=dev, dtype=torch.int64)
        self.loss = torch.zeros(1, device=dev, dtype=f32)
        self.correct = torch.zeros(1, device=dev, dtype=torch.int32)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self.scalars_dev = torch.zeros(2, device=dev, dtype=f32)
        self.steps_done = 0
        self._graphs: Dict[tuple, torch.cuda.CUDAGraph] = {}
        self._eager_steps = 0
        self._prepare_operands()

    # ------------------------------------------------------------------ parameters
    def load_state_dict(self, sd: Dict[str, torch.Tensor]):
        """reference-layout keys (conv1.weight [256,1024,1,1] is split into its per-trunk halves)"""
        f32 = lambda k: sd[k].detach().to(device=self.dev, dtype=torch.float32)  # noqa: E731
        w1 = f32("conv1.weight")
        if tuple(w1.shape) != (256, 1024, 1, 1):
            raise ValueError("bad shape")
        self.p["conv1.weight.k"].copy_(w1[:, :512])
        self.p["conv1.weight.k1"].copy_(w1[:, 512:])
        for n, shape in _TRAIN_SLOTS[2:]:
            t = f32(n)
            if tuple(t.shape) != shape:
                raise ValueError("bad shape")
            self.p[n].copy_(t)

    def state_dict(self) -> Dict[str, torch.Tensor]:
        """the reference module's keys: frozen trunk as loaded + the trained head (host tensors)"""
        out = dict(self._trunk_sd)
        out["conv1.weight"] = torch.cat([self.p["conv1.weight.k"], self.p["conv1.weight.k1"]], dim=1).cpu()
        for n, _ in _TRAIN_SLOTS[2:]:
            out[n] = self.p[n].detach().cpu().clone()
        return out

    def _prepare_operands(self):
        p = self.p
        ops.weight_prep(p["conv1.weight.k"], self.wf["conv1.weight.k"], self.shift["conv1.weight.k"],
                        bias=self._zero_bias)
        ops.weight_prep(p["conv1.weight.k1"], self.wf["conv1.weight.k1"], self.shift["conv1.weight.k1"],
                        bias=p["conv1.bias"])
        for n, b in (("conv2.weight", "conv2.bias"), ("conv3.weight", "conv3.bias")):
            ops.weight_prep(p[n], self.wf[n], self.shift[n], w_dgrad=self.wd[n], bias=p[b])

    # ------------------------------------------------------------------ forward / backward
    def _forward(self, train: bool):
        B, ws, p = self.B, self.ws, self.p
        ops.stem_pack(self.k, ws.xp[:B])
        ops.stem_pack(self.k1, ws.xp[B:])
        feat = E.forward_packed(self.plan, self.W, self.P, ws, trunk_only=True)      # [2B,7,7,512], no grad
        ops.conv_gemm(feat[:B], self.wf["conv1.weight.k"], 1, 0, 0, out=self.t1)
        ops.conv_gemm(feat[B:], self.wf["conv1.weight.k1"], 1, 0, 0, shift=self.shift["conv1.weight.k1"],
                      residual=self.t1, relu=True, out=self.x1)
        ops.conv_gemm(self.x1, self.wf["conv2.weight"], 1, 0, 0, shift=self.shift["conv2.weight"], relu=True,
                      out=self.x2)
        ops.conv_gemm(self.x2, self.wf["conv3.weight"], 1, 0, 0, shift=self.shift["conv3.weight"], relu=True,
                      out=self.x3)
        ops.head_flatten_fwd(self.x3, self.flat)                                      # NCHW order: c*9 + p
        ops.linear_fwd(self.flat, p["fc1.weight"], p["fc1.bias"], True, self.h1)
        h = self.h1
        if train:
            h = ops.dropout_apply(self.h1, self.keep, 1.0 / (1.0 - DROPOUT_P), self.hd)
        ops.linear_fwd(h, p["fc2.weight"], p["fc2.bias"], True, self.z)               # ReLU after fc2 (:79-80)
        ops.linear_fwd(self.z, p["fc_accuracy.weight"], p["fc_accuracy.bias"], False, self.y)
        return feat

    def _wgrad(self, name, x, dy):
        c = self.spec[name]
        splits = E.wgrad_splits(c, self.B)
        part = self.part[name][: splits * c.cout * c.K]
        ops.conv_wgrad(x, dy, c.k, c.k, 1, 0, 0, splits=splits, part=part)
        ops.wgrad_finalize(part, self.p[name], self.g[name], splits=splits, Cout=c.cout, Cin=c.cin, R=c.k, S=c.k,
                           K=c.K)

    def _enqueue_step(self):
        B, p, g = self.B, self.p, self.g
        self._g.flat.zero_()                       # bias gradients are accumulated with atomics
        self.loss.zero_()
        self.correct.zero_()
        feat = self._forward(train=True)
        ops.cross_entropy(self.y, self.act, dlogits=self.dy, loss=self.loss, correct=self.correct)
        # ---- fully connected layers (fp32)
        ops.linear_bwd(self.z, p["fc_accuracy.weight"], None, self.dy, g["fc_accuracy.weight"],
                       g["fc_accuracy.bias"], False, dx=self.dz)
        ops.linear_bwd(self.hd, p["fc2.weight"], self.z, self.dz, g["fc2.weight"], g["fc2.bias"], True, dx=self.dhd)
        ops.dropout_apply(self.dhd, self.keep, 1.0 / (1.0 - DROPOUT_P), self.dhd)
        ops.linear_bwd(self.flat, p["fc1.weight"], self.h1, self.dhd, g["fc1.weight"], g["fc1.bias"], True,
                       dx=self.dflat)
        # ---- head convs: each data gradient applies the ReLU mask of the layer below and accumulates
        # that layer's bias gradient (column sums) in its epilogue
        ops.head_flatten_bwd(self.dflat, self.x3, self.dx3, dbias=g["conv3.bias"])
        self._wgrad("conv3.weight", self.x2, self.dx3)
        tn = E._dgrad_tile_n(256)
        ops.conv_gemm(self.dx3, self.wd["conv3.weight"], 1, 2, 2, mask_src=self.x2, colsum=g["conv2.bias"],
                      out=self.dx2, tile_n=tn)
        self._wgrad("conv2.weight", self.x1, self.dx2)
        ops.conv_gemm(self.dx2, self.wd["conv2.weight"], 1, 2, 2, mask_src=self.x1, colsum=g["conv1.bias"],
                      out=self.dx1, tile_n=tn)
        self._wgrad("conv1.weight.k", feat[:B], self.dx1)
        self._wgrad("conv1.weight.k1", feat[B:], self.dx1)      # nothing flows into the frozen trunk
        # ---- Adam + operand refresh
        ops.adam_fused(self._p.flat, self._g.flat, self._m.flat, self._v.flat, lr=self.lr,
                       step_dev=self.step_dev, scalars_dev=self.scalars_dev)
        self._prepare_operands()

    # ------------------------------------------------------------------ public API
    def _load(self, k, k_plus_one, act):
        if k.shape != self.k.shape or k_plus_one.shape != k.shape or act.numel() != self.B:
            raise ValueError("bad shape")
        if k.dtype != self.k.dtype:
            raise ValueError(f"frames must be {self.k.dtype} (see frames_uint8)")
        self.k.copy_(k, non_blocking=True)
        self.k1.copy_(k_plus_one, non_blocking=True)
        self.act.copy_(act.view(-1), non_blocking=True)

    def step(self, k, k_plus_one, act, keep: torch.Tensor = None) -> torch.Tensor:
        """One training iteration.  `keep` ([B,128] uint8 in {0,1}) overrides the dropout draw (parity
        tests replay the reference's draw); by default it comes from the counter-based generator."""
        self._load(k, k_plus_one, act)
        if keep is not None:
            self.keep.copy_(keep.to(device=self.dev, dtype=torch.uint8).view(self.B, 128), non_blocking=True)
        else:
            ops.dropout_mask(self.keep, DROPOUT_P, self.seed, self.steps_done)
        key = (self.lr,)
        if self.use_graph and self._eager_steps >= 1:
            gr = self._graphs.get(key)
            if gr is None:
                gr = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(gr):
                    self._enqueue_step()
                self._graphs[key] = gr
            gr.replay()
        else:
            self._enqueue_step()
            self._eager_steps += 1
        self.steps_done += 1
        return self.loss

    @torch.no_grad()
    def evaluate(self, k, k_plus_one, act):
        """`validate()`'s per-batch work (:147-160): eval-mode forward (dropout off), cross-entropy and
        the number of correct predictions.  Returns (loss, correct) as 1-element device tensors."""
        self._load(k, k_plus_one, act)
        self._forward(train=False)
        loss = torch.zeros(1, device=self.dev, dtype=torch.float32)
        correct = torch.zeros(1, device=self.dev, dtype=torch.int32)
        ops.cross_entropy(self.y, self.act, loss=loss, correct=correct, want_grad=False)
        return loss, correct


# ------------------------------------------------------------------------------------------------
# pseudo-labelling of the quadruplet table (dataset/process_episodes_real.py:165-181)
# ------------------------------------------------------------------------------------------------
def label_frame_pairs(before_paths, after_paths, runner, *, workers: int = 4, root: str = None):
    """The reference's labelling loop: `for be, ae in DataLoader(ImageStream(ims), batch_size=8):
    acts = model(be.cuda(), ae.cuda())[1].argmax(dim=1, keepdim=True)` (:165-177).  `runner` is an
    `InverseActionRunner` (anything with `.B`, `.dev` and `.label(k, k_plus_one)`); frames are decoded to
    uint8 HWC in a thread pool (resize / centre-crop as `imageNetTransformPIL`, the normalisation is the
    first kernel's job) and fed in batches of `runner.B`; the last batch is padded by repeating its final
    pair and trimmed.  Returns int64 [N, 1] -- the shape the reference assigns to the `inverse_actions`
    column."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from . import realdata
    n = len(before_paths)
    if len(after_paths) != n:
        raise ValueError("bad shape")
    B = runner.B
    realdata._ensure_resize()
    resolve = (lambda p: p) if root is None else (lambda p: p if os.path.isabs(p) else os.path.join(root, p))
    pin = torch.cuda.is_available()
    bufs = [torch.empty(B, 224, 224, 3, dtype=torch.uint8, pin_memory=pin) for _ in range(2)]
    out = []
    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        for lo in range(0, n, B):
            rows = list(range(lo, min(lo + B, n)))
            rows += [rows[-1]] * (B - len(rows))
            jobs = []
            for which, paths in ((0, before_paths), (1, after_paths)):
                dst = bufs[which].numpy()
                jobs += [pool.submit(realdata.decode_frame, resolve(paths[r]), dst[i]) for i, r in enumerate(rows)]
            for j in jobs:
                j.result()
            acts = runner.label(bufs[0].to(runner.dev), bufs[1].to(runner.dev))
            out.append(acts.detach().cpu().view(-1, 1)[: min(lo + B, n) - lo].to(torch.int64))
    return torch.cat(out) if out else torch.zeros(0, 1, dtype=torch.int64)


def label_table(feather_path: str, runner, *, workers: int = 4, out_path: str = None):
    """Write the `inverse_actions` column of a quadruplet table the way the data-set builder does
    (dataset/process_episodes_real.py:165-181): label every (before_image, after_image) row and save the
    table.  Paths in the table are relative to its directory."""
    import pandas as pd
    t = pd.read_feather(feather_path)
    acts = label_frame_pairs(list(t["before_image"]), list(t["after_image"]), runner, workers=workers,
                             root=os.path.dirname(os.path.abspath(feather_path)))
    t["inverse_actions"] = acts.numpy()
    t.reset_index(drop=True, inplace=True)
    t.to_feather(out_path or feather_path)
    return acts
