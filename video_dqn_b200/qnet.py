"""Drop-in `HabitatDQNMultiAction` whose forward/backward run on the sm_100a kernels.

Mirrors the reference module's public surface (archs/HabitatDQNMultiAction.py:8-54): the same
constructor signature and attributes, `forward(inp) -> [B, num_classes, action_dim]` fp32,
`set_train()`, the 70-tensor `parameters()` order and the 250-key `state_dict()` layout
(`resnet.*`, the aliasing `features.*`, `top.*`), so it loads the reference's checkpoints
(train_q_network.py:50-57) and is usable from the reference's trainer, value-map visualiser
(visualize_value.py:96-97) and evaluator (evaluation/evaluate.py:110-114) unchanged.

The submodules below exist to own the parameters / buffers under the reference's names; their
own `forward`s are never called.  There is no CPU path and no library fallback: a CPU tensor or
an unsupported configuration raises.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn

from . import engine as E
from . import ops  # noqa: F401
from .optim import arena_epoch


def _make_resnet18(pretrained: bool):
    """torchvision ResNet-18 container.  The reference always starts from ImageNet weights
    (`models.resnet18(pretrained=True)`, archs/HabitatDQNMultiAction.py:11).  They are taken from the local
    torch hub cache (either file name torchvision has used); if they are not there -- no network on the build /
    GPU boxes -- the trunk keeps torchvision's RANDOM init and that is said loudly: a from-scratch run would
    otherwise train a random trunk with identity BatchNorm statistics without a word in the log.  Loading a
    checkpoint afterwards (what every test and the bench do) overwrites it.  Set VDQN_REQUIRE_PRETRAINED=1 to
    make the missing file an error."""
    import os
    import warnings
    import torchvision.models as tvm
    if pretrained:
        from torch.hub import get_dir
        for name in ("resnet18-f37072fd.pth", "resnet18-5c106cde.pth"):
            f = os.path.join(get_dir(), "checkpoints", name)
            if os.path.exists(f):
                m = tvm.resnet18(weights=None)
                m.load_state_dict(torch.load(f, map_location="cpu"))
                return m
        msg = ("ImageNet weights for resnet18 not found in the torch hub cache "
               f"({os.path.join(get_dir(), 'checkpoints')}): the trunk is RANDOMLY initialised, unlike the "
               "reference's models.resnet18(pretrained=True); load a checkpoint or place resnet18-f37072fd.pth there")
        if os.environ.get("VDQN_REQUIRE_PRETRAINED", "0") == "1":
            raise FileNotFoundError(msg)
        warnings.warn(msg, RuntimeWarning, stacklevel=3)
    return tvm.resnet18(weights=None)


def grad_param_names() -> List[str]:
    """The 68 tensors that receive gradients, in `model.parameters()` order
    (indices 0-59 trunk, 62-63 head conv, 64-69 MLP; 60-61 = resnet.fc never get one)."""
    names = ["conv1.weight", "bn1.weight", "bn1.bias"]
    for li in range(1, 5):
        for b in range(2):
            p = f"layer{li}.{b}."
            names += [p + "conv1.weight", p + "bn1.weight", p + "bn1.bias",
                      p + "conv2.weight", p + "bn2.weight", p + "bn2.bias"]
            if li > 1 and b == 0:
                names += [p + "downsample.0.weight", p + "downsample.1.weight", p + "downsample.1.bias"]
    return ["resnet." + n for n in names] + ["features.8.weight", "features.8.bias", "top.0.weight",
                                             "top.0.bias", "top.2.weight", "top.2.bias", "top.4.weight",
                                             "top.4.bias"]


class _QNetFn(torch.autograd.Function):
    """One autograd node for the whole network: forward saves the bf16 NHWC activations in a
    Workspace, backward runs dgrad/wgrad kernels and hands back the 68 parameter gradients."""

    @staticmethod
    def forward(ctx, frames, module, *params):
        st = module._state()
        need_grad = any(ctx.needs_input_grad[2:])
        n = frames.shape[0]
        ws = E.Workspace(st.plan, n, frames.device, train=need_grad)
        q = E.forward(st.plan, st.W, st.P, ws, frames)
        ctx.ws, ctx.module, ctx.need = ws, module, need_grad
        ctx.set_materialize_grads(False)
        return q

    @staticmethod
    def backward(ctx, dq):
        module, ws = ctx.module, ctx.ws
        names = module._grad_names
        if dq is None or not ctx.need:
            return (None, None) + (None,) * len(names)
        st = module._state()
        G, flat = make_grad_arena(st.P, names)
        E.backward(st.plan, st.W, st.P, G, ws, dq.contiguous().clone())
        ctx.ws = None
        return (None, None) + tuple(G[nm] for nm in names)


def make_grad_arena(P: Dict[str, torch.Tensor], names: List[str], flat: torch.Tensor = None):
    """Zeroed flat fp32 gradient arena with one 16-byte-aligned view per parameter."""
    offs, total = [], 0
    for nm in names:
        offs.append(total)
        total += (P[nm].numel() + 3) // 4 * 4
    if flat is None:
        dev = P[names[0]].device
        flat = torch.zeros(total, device=dev, dtype=torch.float32)
    G = {nm: flat[o:o + P[nm].numel()].view(P[nm].shape) for nm, o in zip(names, offs)}
    return G, flat


class _EngineState:
    def __init__(self, plan, device):
        self.plan = plan
        self.W = E.PreparedWeights(plan, device)
        self.P: Dict[str, torch.Tensor] = {}
        self.sig = None


class HabitatDQNMultiAction(nn.Module):
    def __init__(self, action_dim, num_classes=5, extra_capacity=False, panorama=True):
        super().__init__()
        self.resnet = _make_resnet18(pretrained=True)
        self.extra_capacity = extra_capacity
        self.num_classes = num_classes
        self.action_dim = action_dim
        self.panorama = panorama
        self.num_frames = 4 if panorama else 1
        if extra_capacity:
            trunk = list(self.resnet.children())[:-2]
            self.features = nn.Sequential(*trunk, nn.Conv2d(512, 64, (3, 3)), nn.ReLU(), nn.Flatten())
            self.top = nn.Sequential(nn.Linear(1600 * self.num_frames, 512), nn.ReLU(),
                                     nn.Linear(512, 256), nn.ReLU(),
                                     nn.Linear(256, action_dim * self.num_classes))
        else:
            self.features = nn.Sequential(*list(self.resnet.children())[:-1])
            self.top = nn.Linear(512 * self.num_frames, action_dim * self.num_classes)
        self._grad_names = grad_param_names()
        self._eng = None

    # -- reference API ---------------------------------------------------------------------
    def set_train(self):
        """train() with the ResNet trunk (its BatchNorms) kept in eval mode
        (archs/HabitatDQNMultiAction.py:37-40)."""
        self.train()
        if self.extra_capacity:
            self.resnet.eval()

    def forward(self, inp):
        if self.num_frames == 1 and inp.dim() == 4:
            inp = inp.unsqueeze(1)
        if inp.dim() != 5 or inp.shape[1] != self.num_frames:
            raise Exception("bad shape")
        if not inp.is_cuda:
            raise RuntimeError("video_dqn_b200.HabitatDQNMultiAction has no CPU path: move the module "
                               "and its input to a B200 (the CPU oracle lives in oracle/, tests only)")
        if not self.extra_capacity:
            return self._basic_forward(inp)
        if any(m.training for m in self.resnet.modules() if isinstance(m, nn.BatchNorm2d)):
            raise NotImplementedError("trunk BatchNorm in training mode is not implemented: call "
                                      "set_train() (or eval()) as the reference trainer does")
        B, F = inp.shape[0], inp.shape[1]
        if B == 0:
            return inp.new_zeros((0, self.num_classes, self.action_dim), dtype=torch.float32)
        if inp.dtype == torch.uint8:
            if inp.shape[-1] != 3:
                raise Exception("bad shape")
            frames = inp.reshape(B * F, *inp.shape[2:]).contiguous()
        else:
            if inp.shape[2] != 3 or inp.shape[3] != 224 or inp.shape[4] != 224:
                raise Exception("bad shape")
            frames = inp.reshape(B * F, 3, 224, 224).float().contiguous()
        params = [self._state().P[n] for n in self._grad_names]
        q = _QNetFn.apply(frames, self, *params)
        return q.view(-1, self.num_classes, self.action_dim)

    def _basic_forward(self, inp):
        """The `basic` architecture (extra_capacity=False, archs/HabitatDQNMultiAction.py:32-34): trunk +
        global average pool + one Linear.  The module's own forward is the eval-mode one (BatchNorm
        running statistics): what the value-map / policy callers need for checkpoints of that
        architecture.  Its training keeps the trunk BatchNorms in train mode (set_train() only freezes
        them for extra_capacity, :37-40); that step lives in learner_basic.BasicQLearner."""
        if any(m.training for m in self.resnet.modules() if isinstance(m, nn.BatchNorm2d)):
            raise NotImplementedError("the module forward of the `basic` architecture is the eval-mode one: "
                                      "call eval() first, or train it through "
                                      "video_dqn_b200.learner_basic.BasicQLearner (train-mode BatchNorm step)")
        B, F = inp.shape[0], inp.shape[1]
        if B == 0:
            return inp.new_zeros((0, self.num_classes, self.action_dim), dtype=torch.float32)
        if inp.dtype == torch.uint8:
            frames = inp.reshape(B * F, *inp.shape[2:]).contiguous()
        else:
            if inp.shape[2] != 3 or inp.shape[3] != 224 or inp.shape[4] != 224:
                raise Exception("bad shape")
            frames = inp.reshape(B * F, 3, 224, 224).float().contiguous()
        st = self._basic_state()
        dev = st["dev"]
        n = B * F
        ws = st["ws"].get(n)
        if ws is None:
            ws = st["ws"][n] = E.Workspace(st["plan"], n, dev, train=False)
        with torch.no_grad():
            ops.stem_pack(frames, ws.xp)
            feat = E.forward_packed(st["plan"], st["W"], st["P"], ws, trunk_only=True)      # [n,7,7,512]
            pooled = ops.avgpool_fwd(feat)                                                     # [n,512] fp32
            q = ops.linear_fwd(pooled.view(B, F * 512), st["P"]["top.weight"], st["P"]["top.bias"], False)
        return q.view(-1, self.num_classes, self.action_dim)

    def _basic_state(self):
        """folded bf16 operands of the `basic` architecture's forward-only path, re-derived when a parameter or
        buffer changed (torch version counters)"""
        nt = self._named_tensors()
        dev = nt["top.weight"].device
        st = getattr(self, "_basic_eng", None)
        if st is None or st["dev"] != dev:
            plan = E.make_plan(self.action_dim, self.num_classes, self.num_frames)
            st = self._basic_eng = {"dev": dev, "plan": plan, "W": E.PreparedWeights(plan, dev, trunk_only=True),
                                    "sig": None, "ws": {}}
        sig = tuple((t.data_ptr(), t._version) for t in nt.values())
        if sig != st["sig"]:
            st["P"] = {k: v.detach() for k, v in nt.items()}
            st["W"].prepare(st["P"])
            st["sig"] = sig
        return st

    # -- engine plumbing -------------------------------------------------------------------
    def _named_tensors(self) -> Dict[str, torch.Tensor]:
        d = dict(self.named_parameters())
        d.update(dict(self.named_buffers()))
        return d

    def _state(self) -> _EngineState:
        """Engine state for the device the parameters currently live on; the bf16 operands are
        re-derived whenever a parameter / buffer was modified (in-place version counters)."""
        nt = self._named_tensors()
        dev = nt["top.4.weight"].device
        if dev.type != "cuda":
            raise RuntimeError("parameters are not on a CUDA device")
        if self._eng is None or self._eng.W.shift["stem"].device != dev:
            self._eng = _EngineState(E.make_plan(self.action_dim, self.num_classes, self.num_frames), dev)
        st = self._eng
        sig = tuple((t.data_ptr(), t._version, arena_epoch(t)) for t in nt.values())
        if sig != st.sig:
            for k, t in nt.items():
                if t.dtype != torch.float32 and not k.endswith("num_batches_tracked"):
                    raise RuntimeError(f"parameter {k} must be fp32 (got {t.dtype})")
            st.P = nt
            st.W.prepare({k: v.detach() for k, v in nt.items()})
            st.sig = sig
        return st

    @torch.no_grad()
    def value(self, inp):
        """max_a Q -- the value-map call `model(images).max(2).values`
        (visualize_value.py:96-97) without recording a graph."""
        return self.forward(inp).max(2).values
