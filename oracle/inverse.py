"""CPU restatement of the inverse-dynamics model's forward pass (archs/inverse_action2.py:45-100),
the network that labels the `inverse_actions` column of the quadruplet table
(dataset/process_episodes_real.py:92-95,171-179).

TEST INFRASTRUCTURE ONLY (see qstep.py).  Pinned against the reference class itself by
oracle/make_inverse_goldens.py; the training step below (train_inverse_model.py:30-110,176) by
oracle/make_inverse_train_goldens.py.  Eval mode, as the labelling script runs it (`model.eval()`, :95):
the two Dropout2d layers are identities; the ResNet-18 trunk is frozen with eval-mode BN (:56-58,75).

State-dict layout of the reference module: the trunk is `nn.Sequential(children()[:-2])`, so its keys
are positional -- `resnet18.0.weight` (conv1), `resnet18.1.*` (bn1), `resnet18.4-7.*` (layer1-4) --
followed by `conv1/conv2/conv3`, `fc1`, `fc2`, `fc_accuracy`.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from . import qstep

_SEQ = {"0": "conv1", "1": "bn1", "4": "layer1", "5": "layer2", "6": "layer3", "7": "layer4"}


def trunk_keys_to_resnet(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """`resnet18.<i>.<rest>` -> `resnet.<name>.<rest>` (the naming qstep.trunk_forward uses)"""
    out = {}
    for k, v in sd.items():
        if k.startswith("resnet18."):
            idx, _, rest = k[len("resnet18."):].partition(".")
            if idx in _SEQ:
                out[f"resnet.{_SEQ[idx]}.{rest}"] = v
    return out


def init_state(seed: int = 7, randomize_bn: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded random weights in the reference module's key layout (trunk from qstep.init_state)."""
    q = qstep.init_state(seed=seed, randomize_bn=randomize_bn)
    inv = {v: k for k, v in _SEQ.items()}
    sd = {}
    for k, v in q.items():
        if k.startswith("resnet.") and not k.startswith("resnet.fc"):
            name, _, rest = k[len("resnet."):].partition(".")
            sd[f"resnet18.{inv[name]}.{rest}"] = v.clone()
    g = torch.Generator().manual_seed(seed + 1000)

    def uni(*shape, fan_in):
        b = 1.0 / fan_in ** 0.5
        return (torch.rand(*shape, generator=g) * 2 - 1) * b
    for name, shape, fan in (("conv1", (256, 1024, 1, 1), 1024), ("conv2", (256, 256, 3, 3), 2304),
                             ("conv3", (64, 256, 3, 3), 2304), ("fc1", (128, 576), 576), ("fc2", (3, 128), 128),
                             ("fc_accuracy", (3, 3), 3)):
        sd[name + ".weight"] = uni(*shape, fan_in=fan)
        sd[name + ".bias"] = uni(shape[0], fan_in=fan)
    return sd


def forward(sd: Dict[str, torch.Tensor], k: torch.Tensor, k_plus_one: torch.Tensor):
    """(encoding [B,3] = softmax(fc2), y [B,3] = fc_accuracy(fc2)) -- archs/inverse_action2.py:73-100."""
    rs = trunk_keys_to_resnet(sd)
    a = qstep.trunk_forward(rs, k)
    b = qstep.trunk_forward(rs, k_plus_one)
    x = torch.cat([a, b], dim=1)
    x = F.relu(F.conv2d(x, sd["conv1.weight"], sd["conv1.bias"]))
    x = F.relu(F.conv2d(x, sd["conv2.weight"], sd["conv2.bias"]))
    x = F.relu(F.conv2d(x, sd["conv3.weight"], sd["conv3.bias"]))
    x = x.view(x.size(0), -1)
    x = F.relu(F.linear(x, sd["fc1.weight"], sd["fc1.bias"]))
    x = F.linear(x, sd["fc2.weight"], sd["fc2.bias"])
    return torch.softmax(x, dim=1), F.linear(x, sd["fc_accuracy.weight"], sd["fc_accuracy.bias"])


def label(sd, k, k_plus_one) -> torch.Tensor:
    """`model(be, ae)[1].argmax(dim=1)` (dataset/process_episodes_real.py:176-177)"""
    return forward(sd, k, k_plus_one)[1].argmax(dim=1)


# ----------------------------------------------------------------------------
# training step of the inverse model (train_inverse_model.py)
# ----------------------------------------------------------------------------
TRAINABLE = ("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias",
             "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc_accuracy.weight", "fc_accuracy.bias")
DROPOUT_P = 0.5


def train_forward(sd: Dict[str, torch.Tensor], k: torch.Tensor, k_plus_one: torch.Tensor,
                  keep: torch.Tensor | None) -> torch.Tensor:
    """The TRAINER's forward (train_inverse_model.py:55-82) -- not the arch file's: ReLU after fc2, only
    `y = fc_accuracy(.)` is returned, and `dropout1` (nn.Dropout2d(0.5) on the 2-D fc1 output, i.e.
    element dropout) is active in train mode.  `keep` [B,128] in {0,1} is the dropout draw (None: eval
    mode, identity); the trunk is frozen and in eval mode either way (:39-42,58)."""
    rs = trunk_keys_to_resnet(sd)
    with torch.no_grad():
        a = qstep.trunk_forward(rs, k)
        b = qstep.trunk_forward(rs, k_plus_one)
    x = torch.cat([a, b], dim=1)
    x = F.relu(F.conv2d(x, sd["conv1.weight"], sd["conv1.bias"]))
    x = F.relu(F.conv2d(x, sd["conv2.weight"], sd["conv2.bias"]))
    x = F.relu(F.conv2d(x, sd["conv3.weight"], sd["conv3.bias"]))
    x = x.view(x.size(0), -1)
    x = F.relu(F.linear(x, sd["fc1.weight"], sd["fc1.bias"]))
    if keep is not None:
        x = x * keep.to(x.dtype) * (1.0 / (1.0 - DROPOUT_P))
    x = F.relu(F.linear(x, sd["fc2.weight"], sd["fc2.bias"]))
    return F.linear(x, sd["fc_accuracy.weight"], sd["fc_accuracy.bias"])


class InverseOracleTrainer:
    """`optimizer.zero_grad(); y = model(be, ae); loss = CrossEntropyLoss()(y, act); loss.backward();
    optimizer.step()` (train_inverse_model.py:93-110) with `Adam(model.parameters(), lr, weight_decay=0)`
    (:176; the frozen trunk has no gradients and is skipped)."""

    def __init__(self, sd: Dict[str, torch.Tensor], lr: float = 1e-4):
        self.sd = {k: v.clone() for k, v in sd.items()}
        self.lr, self.t = lr, 0
        self.exp_avg = {n: torch.zeros_like(self.sd[n]) for n in TRAINABLE}
        self.exp_avg_sq = {n: torch.zeros_like(self.sd[n]) for n in TRAINABLE}

    def loss_and_grads(self, k, k1, act, keep):
        leaves = {n: self.sd[n].detach().clone().requires_grad_(True) for n in TRAINABLE}
        sd = dict(self.sd)
        sd.update(leaves)
        y = train_forward(sd, k, k1, keep)
        loss = F.cross_entropy(y, act)
        grads = torch.autograd.grad(loss, [leaves[n] for n in TRAINABLE])
        correct = int((y.argmax(dim=1) == act).sum())
        return loss.detach(), dict(zip(TRAINABLE, grads)), y.detach(), correct

    def adam(self, grads):
        import math
        self.t += 1
        b1, b2, eps = 0.9, 0.999, 1e-8
        bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        for n in TRAINABLE:
            g, m, v = grads[n], self.exp_avg[n], self.exp_avg_sq[n]
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
            self.sd[n].addcdiv_(m, denom, value=-self.lr / bc1)

    def step(self, k, k1, act, keep):
        loss, grads, y, correct = self.loss_and_grads(k, k1, act, keep)
        self.adam(grads)
        return loss, grads, y, correct
