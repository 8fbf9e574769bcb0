"""Fused Adam over flat fp32 arenas, `torch.optim.Adam`-compatible.

Replaces `optim.Adam(model.parameters(), lr)` / `.step()` / `.zero_grad()`
(train_q_network.py:124,222,227).  Semantics are torch.optim.Adam's defaults (betas 0.9/0.999,
eps 1e-8, no weight decay, no amsgrad); parameters that never receive a gradient (`resnet.fc.*`)
are skipped and get no state, as in the reference.  `state_dict()` / `load_state_dict()` keep the
torch layout (state index = position in `model.parameters()`, fields `step`, `exp_avg`,
`exp_avg_sq`) so the reference's checkpoints (`optimizer_state_dict`, :198,246) round-trip.

One kernel launch updates all 12.4 M elements (28 B/element of HBM traffic) and can copy the
new parameters into the target network's arena in the same pass.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops

# storage-base pointer -> number of out-of-band (raw pointer) updates; modules include it in
# their "are my bf16 operands stale" signature because kernels do not bump tensor._version.
_ARENA_EPOCH: Dict[int, int] = {}


def arena_epoch(t: torch.Tensor) -> int:
    return _ARENA_EPOCH.get(t.untyped_storage().data_ptr(), 0)


def bump_arena_epoch(t: torch.Tensor):
    k = t.untyped_storage().data_ptr()
    _ARENA_EPOCH[k] = _ARENA_EPOCH.get(k, 0) + 1


class FlatArena:
    """Contiguous fp32 buffer with one 16-byte aligned slot per tensor."""

    def __init__(self, shapes: List[torch.Size], device, alloc=None):
        """`alloc(total, device) -> zeroed fp32 tensor of at least `total` elements`: where the buffer comes from
        (the data-parallel learner puts the gradient arena into NVLink symmetric memory)."""
        self.offsets, total = [], 0
        for s in shapes:
            self.offsets.append(total)
            total += (s.numel() + 3) // 4 * 4
        self.shapes = shapes
        if alloc is None:
            self.flat = torch.zeros(total, device=device, dtype=torch.float32)
        else:
            self.storage = alloc(total, device)          # may be longer (padded for the exchange kernel)
            self.flat = self.storage[:total]

    def view(self, i: int) -> torch.Tensor:
        s = self.shapes[i]
        return self.flat[self.offsets[i]:self.offsets[i] + s.numel()].view(s)

    def views(self) -> List[torch.Tensor]:
        return [self.view(i) for i in range(len(self.shapes))]


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_alloc=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._grad_alloc = grad_alloc
        self._members: Optional[List[torch.nn.Parameter]] = None
        self._p = self._g = self._m = self._v = None
        self._step = 0

    # ---------------------------------------------------------------- arenas
    def _build(self, members: List[torch.nn.Parameter]):
        dev = members[0].device
        shapes = [p.shape for p in members]
        self._p, self._g = FlatArena(shapes, dev), FlatArena(shapes, dev, alloc=self._grad_alloc)
        self._m, self._v = FlatArena(shapes, dev), FlatArena(shapes, dev)
        for i, p in enumerate(members):
            if p.dtype != torch.float32 or not p.is_cuda:
                raise ValueError("FusedAdam needs fp32 CUDA parameters (no CPU path)")
            v = self._p.view(i)
            v.copy_(p.data)
            p.data = v                                   # parameter now lives in the arena
            st = self.state[p]
            if "exp_avg" in st:                          # restored by load_state_dict
                self._m.view(i).copy_(st["exp_avg"])
                self._v.view(i).copy_(st["exp_avg_sq"])
                self._step = int(st["step"]) if not torch.is_tensor(st["step"]) else int(st["step"].item())
            st["step"] = torch.tensor(float(self._step))
            st["exp_avg"], st["exp_avg_sq"] = self._m.view(i), self._v.view(i)
        self._members = members

    def adopt(self, members: List[torch.nn.Parameter]):
        """Move `members` into the arenas now (the fused learner calls this up front so that
        gradients can be produced directly in `grad_views()`)."""
        if self._members is None or [id(p) for p in self._members] != [id(p) for p in members]:
            self._build(members)

    def grad_views(self) -> List[torch.Tensor]:
        return self._g.views()

    @property
    def param_arena(self) -> torch.Tensor:
        return self._p.flat

    @property
    def grad_arena(self) -> torch.Tensor:
        return self._g.flat

    # ---------------------------------------------------------------- step
    @torch.no_grad()
    def step(self, closure=None, *, target_arena: Optional[torch.Tensor] = None, grad_scale: float = 1.0,
             grads_in_arena: bool = False, step_dev=None, scalars_dev=None):
        loss = closure() if closure is not None else None
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam supports a single param group (as the reference uses)")
        grp = self.param_groups[0]
        if not grads_in_arena:
            members = [p for p in grp["params"] if p.grad is not None]
            if not members:
                return loss
            self.adopt(members)
            for i, p in enumerate(members):
                gv = self._g.view(i)
                if p.grad.data_ptr() != gv.data_ptr():
                    gv.copy_(p.grad)
        if self._members is None:
            raise RuntimeError("FusedAdam.step(grads_in_arena=True) before adopt()")
        self._step += 1
        ops.adam_fused(self._p.flat, self._g.flat, self._m.flat, self._v.flat, lr=float(grp["lr"]),
                       step=None if step_dev is not None else self._step, betas=grp["betas"],
                       eps=grp["eps"], target=target_arena, grad_scale=grad_scale,
                       step_dev=step_dev, scalars_dev=scalars_dev)
        bump_arena_epoch(self._p.flat)
        if target_arena is not None:
            bump_arena_epoch(target_arena)
        for p in self._members:
            self.state[p]["step"] = torch.tensor(float(self._step))
        return loss

    def note_graph_replay(self, target_synced: bool = False, target_arena=None):
        """Bookkeeping after a captured step (whose Adam node used the device step counter)."""
        self._step += 1
        bump_arena_epoch(self._p.flat)
        if target_synced and target_arena is not None:
            bump_arena_epoch(target_arena)

    def state_dict(self):
        if self._members is not None:
            for p in self._members:
                self.state[p]["step"] = torch.tensor(float(self._step))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        """torch layout in (e.g. the `optimizer_state_dict` of a reference checkpoint,
        train_q_network.py:198; `step` may be an int as torch 1.3.1 wrote it).  The arenas keep their
        addresses -- gradient views and captured graphs of a learner stay valid: the loaded moments are
        copied into them."""
        super().load_state_dict(state_dict)
        if self._members is None:
            return                                # folded into the arenas when they are built
        if not any("exp_avg" in self.state[p] for p in self._members):
            self._step = 0                        # a snapshot taken before the first step
        for i, p in enumerate(self._members):
            st = self.state[p]
            if "exp_avg" in st:
                self._m.view(i).copy_(st["exp_avg"])
                self._v.view(i).copy_(st["exp_avg_sq"])
                self._step = int(st["step"]) if not torch.is_tensor(st["step"]) else int(st["step"].item())
            else:
                self._m.view(i).zero_()
                self._v.view(i).zero_()
            st["step"] = torch.tensor(float(self._step))
            st["exp_avg"], st["exp_avg_sq"] = self._m.view(i), self._v.view(i)
