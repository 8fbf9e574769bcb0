"""CPU fp32 restatement of the reference Q-learning step.  TEST INFRASTRUCTURE ONLY.

What is restated, and where it lives in the reference (paths relative to the
reference repository root):

* the Q-network graph ``features -> top -> view(-1, 5, A)`` in the shipped
  ``extra_capacity`` / single-frame configuration
  (archs/HabitatDQNMultiAction.py:27-31, 44-54; the trunk is torchvision's
  ResNet-18 ``children()[:-2]`` = conv1, bn1, relu, maxpool, layer1..4, whose
  BasicBlock is torchvision/models/resnet.py:59-105),
* ``set_train()``: the trunk's BatchNorm layers stay in eval mode, i.e. they
  are a fixed per-channel affine of the running statistics
  (archs/HabitatDQNMultiAction.py:37-40),
* ``process_batch`` -- Double-DQN TD loss (train_q_network.py:126-181),
* the loop body ``set_train / zero_grad / backward / Adam.step``
  (train_q_network.py:221-227) and the hard target sync (:215-216).

The network is written functionally over a flat ``{state_dict key: tensor}``
mapping that uses the reference's own 250-key checkpoint layout, so the same
weights can be pushed through the reference class, this oracle, and the CUDA
engine.  ``oracle/make_goldens.py`` pins this file against the imported
reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

NUM_CLASSES = 5
BN_EPS = 1e-5

# (name in `features.*`, in_ch, out_ch, stride, has_downsample)
_STAGES = [("4", 64, 64, 1, False), ("5", 64, 128, 2, True),
           ("6", 128, 256, 2, True), ("7", 256, 512, 2, True)]


@dataclass
class StepConfig:
    """The hot-path hyper-parameters (defaults.py:4-37 overlaid with
    configs/experiments/real_data/config.yml:1-11)."""
    GAMMA: float = 0.99
    LOSS_CLIP: str = "rect"
    LINEAR: bool = False
    REMOVE_BEFORE_REWARD: bool = False
    LEARNING_RATE: float = 1e-4
    TARGET_UPDATE_INTERVAL: int = 8000
    double_dqn: bool = True          # train_q_network.py:142 (143-144 is the plain variant)
    action_dim: int = 3
    TRAIN_ON_GROUND_TRUTH: bool = False   # process_batch(compare_ground_truth=True), :224
    VALUE_LEARNING: bool = False          # NaN-masked regression (:172-176); the net then has 1 action (:38)


# ----------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------

def trunk_param_names() -> List[str]:
    """Trainable trunk tensors in ``model.parameters()`` order, named under
    ``resnet.`` (60 tensors; resnet.fc.* are parameters 60-61 and never get a
    gradient)."""
    names = ["conv1.weight", "bn1.weight", "bn1.bias"]
    for li, (_, _cin, _cout, _s, ds) in enumerate(_STAGES, start=1):
        for b in range(2):
            p = f"layer{li}.{b}."
            names += [p + "conv1.weight", p + "bn1.weight", p + "bn1.bias",
                      p + "conv2.weight", p + "bn2.weight", p + "bn2.bias"]
            if ds and b == 0:
                names += [p + "downsample.0.weight", p + "downsample.1.weight",
                          p + "downsample.1.bias"]
    return ["resnet." + n for n in names]


def grad_param_names() -> List[str]:
    """The 68 tensors that receive gradients, in ``model.parameters()`` order."""
    return trunk_param_names() + ["features.8.weight", "features.8.bias",
                                  "top.0.weight", "top.0.bias", "top.2.weight",
                                  "top.2.bias", "top.4.weight", "top.4.bias"]


def _feat_key(resnet_key: str) -> str:
    """resnet.* -> features.* alias (children()[:-2] positions 0,1,4..7)."""
    k = resnet_key[len("resnet."):]
    head, _, rest = k.partition(".")
    idx = {"conv1": "0", "bn1": "1", "layer1": "4", "layer2": "5",
           "layer3": "6", "layer4": "7"}[head]
    return f"features.{idx}.{rest}"


def init_state(seed: int = 4, action_dim: int = 3, randomize_bn: bool = False, num_frames: int = 1
               ) -> Dict[str, torch.Tensor]:
    """Random-init weights with the distributions torchvision / torch.nn use
    (kaiming-normal fan_out convs, BN gamma=1 beta=0, default Linear/Conv2d
    init for the head).  Keys follow the reference checkpoint layout.
    ``randomize_bn`` perturbs BN affine + running statistics so a mis-folded BN
    epilogue is visible in parity tests."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv_w(cout, cin, k):
        std = math.sqrt(2.0 / (cout * k * k))
        return torch.randn(cout, cin, k, k, generator=g) * std

    def bn(prefix, c):
        if randomize_bn:
            sd[prefix + ".weight"] = torch.rand(c, generator=g) + 0.5
            sd[prefix + ".bias"] = torch.randn(c, generator=g) * 0.1
            sd[prefix + ".running_mean"] = torch.randn(c, generator=g) * 0.1
            sd[prefix + ".running_var"] = torch.rand(c, generator=g) + 0.5
        else:
            sd[prefix + ".weight"] = torch.ones(c)
            sd[prefix + ".bias"] = torch.zeros(c)
            sd[prefix + ".running_mean"] = torch.zeros(c)
            sd[prefix + ".running_var"] = torch.ones(c)
        sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    def uniform(shape, bound):
        return (torch.rand(*shape, generator=g) * 2 - 1) * bound

    sd["resnet.conv1.weight"] = conv_w(64, 3, 7)
    bn("resnet.bn1", 64)
    for li, (_, cin, cout, _s, ds) in enumerate(_STAGES, start=1):
        for b in range(2):
            p = f"resnet.layer{li}.{b}."
            sd[p + "conv1.weight"] = conv_w(cout, cin if b == 0 else cout, 3)
            bn(p + "bn1", cout)
            sd[p + "conv2.weight"] = conv_w(cout, cout, 3)
            bn(p + "bn2", cout)
            if ds and b == 0:
                sd[p + "downsample.0.weight"] = conv_w(cout, cin, 1)
                bn(p + "downsample.1", cout)
    sd["resnet.fc.weight"] = uniform((1000, 512), 1 / math.sqrt(512))
    sd["resnet.fc.bias"] = uniform((1000,), 1 / math.sqrt(512))
    for k in [k for k in sd if k.startswith("resnet.") and not k.startswith("resnet.fc")]:
        sd[_feat_key(k)] = sd[k]            # aliases share storage, as in the reference
    b8 = 1 / math.sqrt(512 * 9)
    sd["features.8.weight"] = uniform((64, 512, 3, 3), b8)
    sd["features.8.bias"] = uniform((64,), b8)
    dims = [(1600 * num_frames, 512), (512, 256), (256, action_dim * NUM_CLASSES)]
    for i, (fin, fout) in zip((0, 2, 4), dims):
        sd[f"top.{i}.weight"] = uniform((fout, fin), 1 / math.sqrt(fin))
        sd[f"top.{i}.bias"] = uniform((fout,), 1 / math.sqrt(fin))
    return sd


# ----------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------

def _bn_eval(x, sd, p):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, BN_EPS)


def trunk_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """ResNet-18 children()[:-2] with eval-mode BN: [B,3,224,224] -> [B,512,7,7]."""
    y = F.conv2d(x, sd["resnet.conv1.weight"], None, 2, 3)
    y = F.relu(_bn_eval(y, sd, "resnet.bn1"))
    y = F.max_pool2d(y, 3, 2, 1)
    for li, (_, _cin, _cout, stride, ds) in enumerate(_STAGES, start=1):
        for b in range(2):
            p = f"resnet.layer{li}.{b}."
            s = stride if b == 0 else 1
            o = F.conv2d(y, sd[p + "conv1.weight"], None, s, 1)
            o = F.relu(_bn_eval(o, sd, p + "bn1"))
            o = F.conv2d(o, sd[p + "conv2.weight"], None, 1, 1)
            o = _bn_eval(o, sd, p + "bn2")
            if ds and b == 0:
                idn = F.conv2d(y, sd[p + "downsample.0.weight"], None, s, 0)
                idn = _bn_eval(idn, sd, p + "downsample.1")
            else:
                idn = y
            y = F.relu(o + idn)
    return y


def q_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, action_dim: int = 3
              ) -> torch.Tensor:
    """Q[B, 5, A].  `x` is [B,3,224,224] (one frame) or [B,F,3,224,224]; with F frames the per-frame
    features are concatenated before the MLP (archs/HabitatDQNMultiAction.py:49-52) and
    top.0 has 1600*F inputs."""
    if x.dim() == 4:
        x = x.unsqueeze(1)
    nf = sd["top.0.weight"].shape[1] // 1600
    if x.shape[1] != nf:
        raise Exception("bad shape")
    feats = []
    for i in range(nf):
        t = trunk_forward(sd, x[:, i])
        hi = F.relu(F.conv2d(t, sd["features.8.weight"], sd["features.8.bias"]))
        feats.append(hi.flatten(1))                    # NCHW flatten: c*25 + y*5 + x
    h = torch.cat(feats, 1)
    h = F.relu(F.linear(h, sd["top.0.weight"], sd["top.0.bias"]))
    h = F.relu(F.linear(h, sd["top.2.weight"], sd["top.2.bias"]))
    q = F.linear(h, sd["top.4.weight"], sd["top.4.bias"])
    return q.view(-1, NUM_CLASSES, action_dim)


# ----------------------------------------------------------------------------
# TD loss (train_q_network.py:126-181)
# ----------------------------------------------------------------------------

def init_state_basic(seed: int = 4, action_dim: int = 3, randomize_bn: bool = True, num_frames: int = 1
                     ) -> Dict[str, torch.Tensor]:
    """State dict of the `basic` architecture (extra_capacity=False,
    archs/HabitatDQNMultiAction.py:32-34): the same trunk keys (resnet.* + features.* aliases), no
    head conv, `top` = one Linear(512*F, A*5)."""
    sd = {k: v for k, v in init_state(seed, action_dim, randomize_bn).items()
          if not (k.startswith("features.8.") or k.startswith("top."))}
    g = torch.Generator().manual_seed(seed + 77)
    b = 1 / math.sqrt(512 * num_frames)
    sd["top.weight"] = (torch.rand(action_dim * NUM_CLASSES, 512 * num_frames, generator=g) * 2 - 1) * b
    sd["top.bias"] = (torch.rand(action_dim * NUM_CLASSES, generator=g) * 2 - 1) * b
    return sd


def q_forward_basic(sd: Dict[str, torch.Tensor], x: torch.Tensor, action_dim: int = 3) -> torch.Tensor:
    """Q[B,5,A] of the `basic` architecture in eval mode: trunk -> AdaptiveAvgPool2d(1) -> cat over
    frames -> Linear (archs/HabitatDQNMultiAction.py:33-34,44-54)."""
    if x.dim() == 4:
        x = x.unsqueeze(1)
    nf = sd["top.weight"].shape[1] // 512
    if x.shape[1] != nf:
        raise Exception("bad shape")
    feats = [trunk_forward(sd, x[:, i]).mean(dim=(2, 3)) for i in range(nf)]
    q = F.linear(torch.cat(feats, 1), sd["top.weight"], sd["top.bias"])
    return q.view(-1, NUM_CLASSES, action_dim)


def td_targets(q_next_online, q_next_target, rew, term, cfg: StepConfig):
    sel = q_next_online if cfg.double_dqn else q_next_target
    best = sel.argmax(-1)                                           # [B,5], first max on ties
    q_a = q_next_target.gather(2, best.unsqueeze(2)).squeeze(2).detach()
    q_a = q_a * (1 - term.float())
    if cfg.LINEAR:
        y = rew.float() + (q_a - 0.1)
    else:
        y = rew.float() + cfg.GAMMA * q_a
    if cfg.LOSS_CLIP == "rect":
        y = torch.clamp(y, max=1, min=0)
    return best, y


def td_loss(q_s, q_next_online, q_next_target, act, rew, term, valid_mask,
            cfg: StepConfig):
    """Returns (loss, aux) with aux = dict(best, y, q_b)."""
    idx = act.view(-1, 1).repeat(1, NUM_CLASSES)
    q_b = q_s.gather(2, idx.unsqueeze(2)).squeeze(2)
    best, y = td_targets(q_next_online, q_next_target, rew, term, cfg)
    losses = 0.5 * (q_b - y) ** 2
    if cfg.REMOVE_BEFORE_REWARD:
        losses = losses * valid_mask
    return losses.mean(), {"best": best, "y": y, "q_b": q_b}


def td_loss_ground_truth(q_s, act, ground_truth, value_learning: bool):
    """The ``compare_ground_truth`` branch (TRAIN_ON_GROUND_TRUTH, train_q_network.py:170-178):
    regress Q_b onto the discounted ground truth gamma^steps_to_reward
    (dataloaders/q_learning_real.py:86-89).  With VALUE_LEARNING the NaN entries (classes never
    reached) are masked: 0.5 (Q_b*mask - gt0)^2; without it the NaNs propagate, as in the reference."""
    idx = act.view(-1, 1).repeat(1, NUM_CLASSES)
    q_b = q_s.gather(2, idx.unsqueeze(2)).squeeze(2)
    if value_learning:
        mask = (1 - torch.isnan(ground_truth).int())
        gt = ground_truth.clone()
        gt[torch.isnan(ground_truth)] = 0
        losses = 0.5 * (q_b * mask - gt.float()) ** 2
    else:
        losses = 0.5 * (q_b - ground_truth.float()) ** 2
    return losses.mean()


def td_grad_closed_form(q_s, act, y, valid_mask, cfg: StepConfig):
    """dLoss/dQ(s): (Q_b - y) * mask / (5B) on the taken action, else 0."""
    B = q_s.shape[0]
    idx = act.view(-1, 1).repeat(1, NUM_CLASSES)
    q_b = q_s.gather(2, idx.unsqueeze(2)).squeeze(2)
    d = (q_b - y) / float(NUM_CLASSES * B)
    if cfg.REMOVE_BEFORE_REWARD:
        d = d * valid_mask * valid_mask.new_ones(()).float()
    g = torch.zeros_like(q_s)
    g.scatter_(2, idx.unsqueeze(2), d.unsqueeze(2))
    return g


# ----------------------------------------------------------------------------
# the step: loss, backward, Adam, target sync
# ----------------------------------------------------------------------------

class OracleTrainer:
    """Holds online + target weights and Adam state; ``step(batch)`` is one
    iteration of the reference loop body."""

    def __init__(self, sd: Dict[str, torch.Tensor], cfg: StepConfig | None = None):
        self.cfg = cfg or StepConfig()
        self.sd = {k: v.clone() for k, v in sd.items()}
        self._realias(self.sd)
        self.target = {k: v.clone() for k, v in self.sd.items()}
        self._realias(self.target)
        self.names = grad_param_names()
        self.exp_avg = {n: torch.zeros_like(self.sd[n]) for n in self.names}
        self.exp_avg_sq = {n: torch.zeros_like(self.sd[n]) for n in self.names}
        self.t = 0
        self.sample_number = 0
        self.betas, self.eps = (0.9, 0.999), 1e-8

    @staticmethod
    def _realias(sd):
        for k in list(sd):
            if k.startswith("resnet.") and not k.startswith("resnet.fc"):
                sd[_feat_key(k)] = sd[k]

    def sync_target(self):
        self.target = {k: v.clone() for k, v in self.sd.items()}
        self._realias(self.target)

    def loss_and_grads(self, batch) -> Tuple[torch.Tensor, Dict[str, torch.Tensor], dict]:
        before, after, act, rew, term, _gt, valid_mask = batch
        leaves = {n: self.sd[n].detach().clone().requires_grad_(True) for n in self.names}
        sd = dict(self.sd)
        sd.update(leaves)
        self._realias(sd)
        A = self.cfg.action_dim
        q_s = q_forward(sd, before, A)
        if self.cfg.TRAIN_ON_GROUND_TRUTH:
            loss = td_loss_ground_truth(q_s, act, _gt, self.cfg.VALUE_LEARNING)
            aux = {"q_s": q_s.detach()}
            grads = torch.autograd.grad(loss, [leaves[n] for n in self.names])
            return loss.detach(), dict(zip(self.names, grads)), aux
        with torch.no_grad():
            q_nt = q_forward(self.target, after, A)
            q_no = q_forward(sd, after, A)
        loss, aux = td_loss(q_s, q_no, q_nt, act, rew, term, valid_mask, self.cfg)
        grads = torch.autograd.grad(loss, [leaves[n] for n in self.names])
        aux.update(q_s=q_s.detach(), q_next_online=q_no, q_next_target=q_nt)
        return loss.detach(), dict(zip(self.names, grads)), aux

    def adam(self, grads: Dict[str, torch.Tensor]):
        """torch.optim.Adam defaults (betas 0.9/0.999, eps 1e-8, no weight decay)."""
        self.t += 1
        b1, b2 = self.betas
        bc1 = 1 - b1 ** self.t
        bc2 = 1 - b2 ** self.t
        step_size = self.cfg.LEARNING_RATE / bc1
        for n in self.names:
            g = grads[n]
            m, v = self.exp_avg[n], self.exp_avg_sq[n]
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
            self.sd[n].addcdiv_(m, denom, value=-step_size)

    def step(self, batch):
        self.sample_number += 1
        if self.sample_number % self.cfg.TARGET_UPDATE_INTERVAL == 0:
            self.sync_target()
        loss, grads, aux = self.loss_and_grads(batch)
        self.adam(grads)
        return loss, grads, aux


# ----------------------------------------------------------------------------
# synthetic quadruplets
# ----------------------------------------------------------------------------

def synthetic_batch(B: int, seed: int = 1, action_dim: int = 3, uint8: bool = False):
    """(before, after, act, rew, term, gt, valid_mask) with the loader's dtypes
    (dataloaders/q_learning_real.py:75-98): frames fp32 [B,3,224,224] ~ N(0,1)
    (or uint8 HWC when ``uint8``), act int64 [B], rew = term int64 [B,5] ~
    Bernoulli(0.1), gt NaN float64 [B], valid_mask int64 ones."""
    g = torch.Generator().manual_seed(seed)
    if uint8:
        before = torch.randint(0, 256, (B, 224, 224, 3), generator=g, dtype=torch.uint8)
        after = torch.randint(0, 256, (B, 224, 224, 3), generator=g, dtype=torch.uint8)
    else:
        before = torch.randn(B, 3, 224, 224, generator=g)
        after = torch.randn(B, 3, 224, 224, generator=g)
    act = torch.randint(0, action_dim, (B,), generator=g, dtype=torch.int64)
    rew = (torch.rand(B, NUM_CLASSES, generator=g) < 0.1).to(torch.int64)
    term = rew.clone()
    gt = torch.full((B,), float("nan"), dtype=torch.float64)
    valid = torch.ones(B, NUM_CLASSES, dtype=torch.int64)
    return before, after, act, rew, term, gt, valid


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def to_imgnet(im_u8_hwc: torch.Tensor) -> torch.Tensor:
    """uint8 [B,H,W,3] -> normalised fp32 [B,3,H,W] (util/torch.py:26-36)."""
    x = im_u8_hwc.float() / 255
    x = x.permute(0, 3, 1, 2)
    x = x - torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    return x / torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)


# ----------------------------------------------------------------------------
# `basic` architecture, training step
# ----------------------------------------------------------------------------
BN_MOMENTUM = 0.1


def _bn_train(x, sd, p):
    """BatchNorm2d in TRAIN mode: batch statistics, running statistics updated in place (momentum 0.1,
    unbiased variance), num_batches_tracked incremented (torch.nn.BatchNorm2d.forward)."""
    key = p + ".num_batches_tracked"
    if key in sd:
        sd[key] += 1
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        True, BN_MOMENTUM, BN_EPS)


def trunk_forward_bn_train(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """ResNet-18 children()[:-2] with TRAIN-mode BN (what `set_train()` leaves the `basic` architecture
    in: archs/HabitatDQNMultiAction.py:37-40 only freezes the trunk for extra_capacity).  Mutates the
    running statistics in `sd`."""
    y = F.conv2d(x, sd["resnet.conv1.weight"], None, 2, 3)
    y = F.relu(_bn_train(y, sd, "resnet.bn1"))
    y = F.max_pool2d(y, 3, 2, 1)
    for li, (_, _cin, _cout, stride, ds) in enumerate(_STAGES, start=1):
        for b in range(2):
            p = f"resnet.layer{li}.{b}."
            s = stride if b == 0 else 1
            o = F.conv2d(y, sd[p + "conv1.weight"], None, s, 1)
            o = F.relu(_bn_train(o, sd, p + "bn1"))
            o = F.conv2d(o, sd[p + "conv2.weight"], None, 1, 1)
            o = _bn_train(o, sd, p + "bn2")
            if ds and b == 0:
                idn = F.conv2d(y, sd[p + "downsample.0.weight"], None, s, 0)
                idn = _bn_train(idn, sd, p + "downsample.1")
            else:
                idn = y
            y = F.relu(o + idn)
    return y


def q_forward_basic_train(sd, x, action_dim: int = 3):
    """Q[B,5,A] of the `basic` architecture with train-mode BN (mutates running stats).  With F frames
    ([B,F,3,224,224], panorama / previous-images networks) every frame goes through the trunk separately, in
    order (archs/HabitatDQNMultiAction.py:49-51): F sets of batch statistics and F running-statistics
    updates per forward."""
    if x.dim() == 4:
        x = x.unsqueeze(1)
    nf = sd["top.weight"].shape[1] // 512
    if x.shape[1] != nf:
        raise Exception("bad shape")
    feats = [trunk_forward_bn_train(sd, x[:, i]).mean(dim=(2, 3)) for i in range(nf)]
    return F.linear(torch.cat(feats, 1), sd["top.weight"], sd["top.bias"]).view(-1, NUM_CLASSES, action_dim)


def grad_param_names_basic() -> List[str]:
    """The 62 tensors of the `basic` architecture that receive gradients, `model.parameters()` order."""
    return trunk_param_names() + ["top.weight", "top.bias"]


class BasicOracleTrainer:
    """One iteration of the reference loop body (train_q_network.py:211-229) for ARCHITECTURE != 
    'extra_capacity': `model(before)` and `model(after)` run the trunk BatchNorms in train mode (in
    that order: both update the running statistics, :131,142), `target_net(after)` in eval mode
    (:122,140), backward through `model(before)`, Adam."""

    def __init__(self, sd: Dict[str, torch.Tensor], cfg: StepConfig | None = None):
        self.cfg = cfg or StepConfig()
        self.sd = {k: v.clone() for k, v in sd.items()}
        self.target = {k: v.clone() for k, v in self.sd.items()}
        self.names = grad_param_names_basic()
        self.exp_avg = {n: torch.zeros_like(self.sd[n]) for n in self.names}
        self.exp_avg_sq = {n: torch.zeros_like(self.sd[n]) for n in self.names}
        self.t = 0
        self.betas, self.eps = (0.9, 0.999), 1e-8

    def sync_target(self):
        self.target = {k: v.clone() for k, v in self.sd.items()}

    def loss_and_grads(self, batch):
        before, after, act, rew, term, _gt, valid_mask = batch
        leaves = {n: self.sd[n].detach().clone().requires_grad_(True) for n in self.names}
        sd = dict(self.sd)                      # buffers are shared: running stats update self.sd
        sd.update(leaves)
        A = self.cfg.action_dim
        q_s = q_forward_basic_train(sd, before, A)
        with torch.no_grad():
            q_nt = q_forward_basic(self.target, after, A)
            q_no = q_forward_basic_train(sd, after, A)
        loss, aux = td_loss(q_s, q_no, q_nt, act, rew, term, valid_mask, self.cfg)
        grads = torch.autograd.grad(loss, [leaves[n] for n in self.names])
        aux.update(q_s=q_s.detach(), q_next_online=q_no, q_next_target=q_nt)
        return loss.detach(), dict(zip(self.names, grads)), aux

    adam = OracleTrainer.adam

    def step(self, batch):
        loss, grads, aux = self.loss_and_grads(batch)
        self.adam(grads)
        return loss, grads, aux
