"""Second oracle: the reference Q-learning step with the CUDA path's ROUNDING POINTS.  TEST
INFRASTRUCTURE ONLY (imported by tests/ and smoke(); never by the product path).

`oracle/qstep.py` is the fp32 restatement pinned to the reference.  This file restates the SAME graph (it
calls qstep for everything that is not a rounding decision) but rounds where the kernels round.  What it
established (tests/test_gpu_parity_full.py, profiles/grad_bars_r02.json): even with identical rounding
points the CUDA path and this oracle differ by as much as either differs from fp32 (gradient rel-L2 0.069
vs 0.071) -- bf16 rounding makes the network chaotic at the ulp scale, so accumulation order alone
decorrelates two pipelines within a few layers.  End-to-end bars therefore stay at bf16-noise level; the
tight (1e-3) bar is per layer, teacher-forced (tests/test_gpu_teacher_forced.py, which reuses the rounding
helpers below).  Rounding points:

* input frames and every stored activation are bf16 (stem_pack, conv epilogues);
* conv operands are the BatchNorm-folded weights rounded to bf16, `w * gamma / sqrt(var + eps)`,
  accumulated in fp32, with `beta - mean * gamma / sqrt(var + eps)` added in fp32 (engine.py
  PreparedWeights; exact for the eval-mode BatchNorm of archs/HabitatDQNMultiAction.py:37-40);
* every gradient that the backward pass stores (d/d activation) is bf16; weight gradients are
  fp32 sums of products of those bf16 values; the residual that the 1x1 downsample hands back is
  rounded before it is added (it makes an HBM round trip in bf16);
* the Q-head MLP (top.0 / top.2 / top.4) and the TD loss are fp32 (train_q_network.py:126-181).

Nothing here is independent evidence about the reference: parity is pinned by qstep.py + the golden
vectors.  With the rounding functions switched off every function reduces to qstep's
(tests/test_oracle_golden.py::test_rounding_point_oracle_is_the_same_graph).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import qstep
from .qstep import BN_EPS, NUM_CLASSES, _STAGES


class _RoundBoth(torch.autograd.Function):
    """value -> bf16 in the forward pass, gradient -> bf16 in the backward pass"""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


class _RoundValue(torch.autograd.Function):
    """value -> bf16, gradient passed through (weights: the fp32 master receives the fp32 sum)"""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundGrad(torch.autograd.Function):
    """identity in the forward pass, gradient -> bf16"""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(torch.float32)


def _act(x):
    return _RoundBoth.apply(x)


def _folded_conv(x, sd, wkey, bn, stride, pad):
    """conv + eval-mode BatchNorm as the kernels compute it: bf16(w * scale) operands, fp32 shift"""
    w = sd[wkey]
    rstd = torch.rsqrt(sd[bn + ".running_var"] + BN_EPS)
    scale = sd[bn + ".weight"] * rstd
    shift = sd[bn + ".bias"] - sd[bn + ".running_mean"] * scale
    wf = _RoundValue.apply(w * scale.view(-1, 1, 1, 1))
    return F.conv2d(x, wf, None, stride, pad) + shift.view(1, -1, 1, 1)


def trunk_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """qstep.trunk_forward with the kernels' rounding points."""
    x = _RoundValue.apply(x)                                         # stem_pack: frames -> bf16
    y = _act(F.relu(_folded_conv(x, sd, "resnet.conv1.weight", "resnet.bn1", 2, 3)))
    y = F.max_pool2d(y, 3, 2, 1)
    for li, (_, _cin, _cout, stride, ds) in enumerate(_STAGES, start=1):
        for b in range(2):
            p = f"resnet.layer{li}.{b}."
            s = stride if b == 0 else 1
            o = _act(F.relu(_folded_conv(y, sd, p + "conv1.weight", p + "bn1", s, 1)))
            o = _folded_conv(o, sd, p + "conv2.weight", p + "bn2", 1, 1)
            if ds and b == 0:
                idn = _act(_folded_conv(_RoundGrad.apply(y), sd, p + "downsample.0.weight",
                                        p + "downsample.1", s, 0))
            else:
                idn = y
            y = _act(F.relu(o + idn))
    return y


def q_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, action_dim: int = 3) -> torch.Tensor:
    """qstep.q_forward with the kernels' rounding points (head conv: bf16 operands, fp32 bias; MLP fp32)."""
    if x.dim() == 4:
        x = x.unsqueeze(1)
    nf = sd["top.0.weight"].shape[1] // 1600
    if x.shape[1] != nf:
        raise Exception("bad shape")
    feats = []
    for i in range(nf):
        t = trunk_forward(sd, x[:, i])
        w8 = _RoundValue.apply(sd["features.8.weight"])
        hi = _act(F.relu(F.conv2d(t, w8, None) + sd["features.8.bias"].view(1, -1, 1, 1)))
        feats.append(hi.flatten(1))
    h = torch.cat(feats, 1)
    h = F.relu(F.linear(h, sd["top.0.weight"], sd["top.0.bias"]))
    h = F.relu(F.linear(h, sd["top.2.weight"], sd["top.2.bias"]))
    q = F.linear(h, sd["top.4.weight"], sd["top.4.bias"])
    return q.view(-1, NUM_CLASSES, action_dim)


class EmulatedTrainer(qstep.OracleTrainer):
    """qstep.OracleTrainer whose three forwards and backward use the rounding points above."""

    def loss_and_grads(self, batch) -> Tuple[torch.Tensor, Dict[str, torch.Tensor], dict]:
        before, after, act, rew, term, _gt, valid_mask = batch
        leaves = {n: self.sd[n].detach().clone().requires_grad_(True) for n in self.names}
        sd = dict(self.sd)
        sd.update(leaves)
        self._realias(sd)
        A = self.cfg.action_dim
        q_s = q_forward(sd, before, A)
        if self.cfg.TRAIN_ON_GROUND_TRUTH:
            loss = qstep.td_loss_ground_truth(q_s, act, _gt, self.cfg.VALUE_LEARNING)
            aux = {"q_s": q_s.detach()}
            grads = torch.autograd.grad(loss, [leaves[n] for n in self.names])
            return loss.detach(), dict(zip(self.names, grads)), aux
        with torch.no_grad():
            q_nt = q_forward(self.target, after, A)
            q_no = q_forward(sd, after, A)
        loss, aux = qstep.td_loss(q_s, q_no, q_nt, act, rew, term, valid_mask, self.cfg)
        grads = torch.autograd.grad(loss, [leaves[n] for n in self.names])
        aux.update(q_s=q_s.detach(), q_next_online=q_no, q_next_target=q_nt)
        return loss.detach(), dict(zip(self.names, grads)), aux


def grad_report(got: Dict[str, torch.Tensor], ref: Dict[str, torch.Tensor], names=None):
    """Per-tensor (cosine, norm ratio - 1, rel-L2) and the global rel-L2 of `got` against `ref`."""
    names = names or list(ref)
    rows, num, den = {}, 0.0, 0.0
    for n in names:
        a, b = got[n].detach().cpu().double().flatten(), ref[n].detach().cpu().double().flatten()
        na, nb = a.norm().item(), b.norm().item()
        d2 = (a - b).pow(2).sum().item()
        rows[n] = (float(a @ b / (na * nb + 1e-300)), na / (nb + 1e-300) - 1.0, (d2 ** 0.5) / (nb + 1e-300))
        num += d2
        den += nb * nb
    return rows, (num / den) ** 0.5
