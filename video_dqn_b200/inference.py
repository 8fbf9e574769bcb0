"""Forward-only entry points of the Q-network: the two callers of
`HabitatDQNMultiAction.forward` outside training.

* value maps -- `build_map_gibson` pushes batches of 32 pre-rendered views through
  `model(images).max(2).values` (visualize_value.py:78-98; the reference builds and discards an
  autograd graph per batch because it has no `no_grad`);
* navigation policy -- `model_score` scores one view at a time:
  `to_imgnet(uint8 HWC frame)` -> `model(x[None])[0, class, :].max().item()`
  (evaluation/evaluate.py:110-114, util/torch.py:26-36), 12 rotations per decision (:183-215).

`QValueRunner` keeps one forward-only workspace for a fixed batch shape, takes uint8 HWC frames
(normalisation fused into the first kernel) or the loader's fp32 NCHW tensors, and replays the whole
forward as a CUDA graph -- at batch 1 the reference path is launch-latency bound (~210 launches).
"""
from __future__ import annotations

import torch

from . import engine as E
from . import ops
from .qnet import HabitatDQNMultiAction


class QValueRunner:
    def __init__(self, model: HabitatDQNMultiAction, batch: int, *, frames_uint8: bool = True,
                 use_graph: bool = True):
        if any(m.training for m in model.resnet.modules() if isinstance(m, torch.nn.BatchNorm2d)):
            raise NotImplementedError("call model.eval() or model.set_train() first")
        self.model, self.B = model, batch
        st = model._state()
        self.plan = st.plan
        F = self.plan.num_frames
        dev = st.W.shift["stem"].device
        self.ws = E.Workspace(self.plan, batch * F, dev, train=False)
        if frames_uint8:
            shp = (batch, F, 224, 224, 3) if F > 1 else (batch, 224, 224, 3)
            self.frames = torch.zeros(shp, device=dev, dtype=torch.uint8)
        else:
            shp = (batch, F, 3, 224, 224) if F > 1 else (batch, 3, 224, 224)
            self.frames = torch.zeros(shp, device=dev, dtype=torch.float32)
        C, A = self.plan.num_classes, self.plan.action_dim
        self.value = torch.empty(batch, C, device=dev, dtype=torch.float32)
        self.best = torch.empty(batch, C, device=dev, dtype=torch.int64)
        self.use_graph, self._graph, self._warm = use_graph, None, False

    def _enqueue(self):
        st = self.model._eng
        q = E.forward(self.plan, st.W, st.P, self.ws, self.frames.view(-1, *self.frames.shape[-3:]))
        B, C, A = self.B, self.plan.num_classes, self.plan.action_dim
        ops.q_max(q.view(B, C, A), self.value, self.best)

    @torch.no_grad()
    def __call__(self, frames: torch.Tensor = None):
        """frames: [B, 224, 224, 3] uint8 (or [B, 3, 224, 224] fp32), host or device; returns
        (Q [B, C, A], value [B, C] = max_a Q, best action [B, C]) as views of static buffers."""
        if frames is not None:
            if frames.shape[0] != self.B:
                raise ValueError("bad shape")
            self.frames.copy_(frames.view(self.frames.shape), non_blocking=True)
        self.model._state()                      # refresh the bf16 operands if the weights changed
        if self.use_graph and self._warm:
            if self._graph is None:
                self._graph = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(self._graph):
                    self._enqueue()
            self._graph.replay()
        else:
            self._enqueue()
            self._warm = True
        B, C, A = self.B, self.plan.num_classes, self.plan.action_dim
        return self.ws.q.view(B, C, A), self.value, self.best

    def score(self, frame_u8_hwc: torch.Tensor, class_index: int) -> float:
        """evaluation/evaluate.py:110-114 for one view (B must be 1)."""
        _q, value, _ = self(frame_u8_hwc.view(1, *frame_u8_hwc.shape[-3:]))
        return value[0, class_index].item()
