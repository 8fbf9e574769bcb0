"""Data-parallel gradient exchange for the fused step: one process per GPU, NCCL over
NVLink/NVSwitch, bucketed all-reduce launched from inside the backward pass so it overlaps the
remaining data/weight-gradient kernels (the reference itself is single-GPU,
train_q_network.py:275).

The flat gradient arena is laid out in `model.parameters()` order and the backward pass finishes
parameters in reverse order, so after every backward stage the *ready* gradients form a contiguous
suffix of the arena: a bucket is simply `arena[lo:hi]`.  The sum is turned into the mean by the
Adam kernel's `grad_scale = 1/world`.  Samples are independent (eval-mode BatchNorm: no cross-
sample statistics), so this is the only exchange in the step.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, learner, process_group=None, bucket_bytes: int = None):
        if bucket_bytes is None:
            import os
            bucket_bytes = int(os.environ.get("VDQN_DDP_BUCKET_MB", "8")) << 20
        self.pg = process_group
        self.arena = learner.opt.grad_arena
        self.bucket_elems = bucket_bytes // 4
        names: List[str] = learner.model._grad_names
        G: Dict[str, torch.Tensor] = learner.G
        base = self.arena.data_ptr()
        off = {n: (G[n].data_ptr() - base) // 4 for n in names}
        plan = learner.plan
        # stage name -> offset of the first parameter that is complete once the stage has run
        self.stage_lo = {"head": off["features.8.weight"], "stem": 0}
        for b in plan.blocks:
            self.stage_lo[b.conv1.name] = off[b.conv1.wkey]
        self.hi = self.arena.numel()
        self.pending_lo = self.hi
        self.on_cuda = self.arena.is_cuda
        self.comm = torch.cuda.Stream() if self.on_cuda else None
        self.launched: List[tuple] = []          # (lo, hi) per all-reduce, for tests / reporting

    def _reduce(self, lo: int, hi: int):
        if hi <= lo:
            return
        buf = self.arena[lo:hi]
        self.launched.append((lo, hi))
        if self.on_cuda:
            ev = torch.cuda.Event()
            ev.record()
            self.comm.wait_event(ev)
            with torch.cuda.stream(self.comm):
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.pg)

    def on_stage(self, stage: str):
        lo = self.stage_lo.get(stage)
        if lo is None:
            return
        self.pending_lo = lo
        if self.hi - lo >= self.bucket_elems or stage == "stem":
            self._reduce(lo, self.hi)
            self.hi = lo

    def finish(self):
        """Flush what is left and make the main stream wait for the exchange."""
        self._reduce(0, self.hi)
        if self.on_cuda:
            torch.cuda.current_stream().wait_stream(self.comm)
        self.hi = self.arena.numel()
        self.pending_lo = self.hi
