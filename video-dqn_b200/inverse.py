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
