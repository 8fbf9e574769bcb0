"""Batch staging: reference-format batches (dataloaders/q_learning_real.py:98 7-tuples) go through
pinned host buffers and asynchronous H2D copies on a copy stream, double-buffered so the copy of
batch k+1 overlaps the step on batch k.  (The reference does 7 synchronous `.to(device)` calls from
pageable memory per step, train_q_network.py:127-129.)  Frames may be uint8 HWC -- what a JPEG
decoder yields; `to_imgnet` normalisation (util/torch.py:26-36) is then fused into the stem-pack
kernel and the copy is 4x smaller -- or the loader's fp32 NCHW tensors."""
from __future__ import annotations

import torch


class BatchStager:
    def __init__(self, learner, depth: int = 2):
        self.lr = learner
        self.copy_stream = torch.cuda.Stream()
        # ground-truth regression (TRAIN_ON_GROUND_TRUTH) consumes only (before, act, gt)
        fields = ("before", "act", "gt") if getattr(learner, "gt_mode", False) else \
            ("before", "after", "act", "rew", "term", "valid")
        self.fields = fields
        self.dev = [{f: torch.empty_like(getattr(learner, f)) for f in fields} for _ in range(depth)]
        self.pin = [{f: torch.empty(getattr(learner, f).shape, dtype=getattr(learner, f).dtype,
                                    pin_memory=True) for f in fields} for _ in range(depth)]
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.consumed = [torch.cuda.Event() for _ in range(depth)]
        for e in self.consumed:
            e.record()
        self.depth, self.head, self.tail = depth, 0, 0
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.pin[0].values())

    def push(self, batch):
        """Host side: copy into pinned memory (if not already pinned) and enqueue the H2D."""
        i = self.head % self.depth
        before, after, act, rew, term, gt, valid = batch
        src = dict(before=before, after=after, act=act.view(-1), rew=rew, term=term, valid=valid, gt=gt)
        src = {f: src[f] for f in self.fields}
        if "rew" in src:
            from .learner import check_label_dtypes
            check_label_dtypes(self.lr, src["rew"], src["term"])
        self.consumed[i].synchronize()                 # slot free (its D2D copy has run)
        with torch.cuda.stream(self.copy_stream):
            for f, t in src.items():
                p = self.pin[i][f]
                if t.is_pinned() and t.dtype == p.dtype:
                    self.dev[i][f].copy_(t.view(p.shape), non_blocking=True)
                else:
                    p.copy_(t.view(p.shape))
                    self.dev[i][f].copy_(p, non_blocking=True)
            self.ready[i].record(self.copy_stream)
        self.head += 1

    def pop_into_learner(self):
        """Device side: wait for the H2D, move the batch into the step's static buffers."""
        i = self.tail % self.depth
        cur = torch.cuda.current_stream()
        cur.wait_event(self.ready[i])
        for f, t in self.dev[i].items():
            getattr(self.lr, f).copy_(t, non_blocking=True)
        self.consumed[i].record(cur)
        self.tail += 1
