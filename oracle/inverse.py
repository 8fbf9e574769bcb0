"""CPU restatement of the inverse-dynamics model's forward pass (archs/inverse_action2.py:45-100),
the network that labels the `inverse_actions` column of the quadruplet table
(dataset/process_episodes_real.py:92-95,171-179).

TEST INFRASTRUCTURE ONLY (see qstep.py).  Pinned against the reference class itself by
oracle/make_inverse_goldens.py.  Eval mode, as the labelling script runs it (`model.eval()`, :95):
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
