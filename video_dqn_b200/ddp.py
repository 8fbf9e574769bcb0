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


class NvlGradSync:
    """The gradient exchange as ONE kernel of this library over NVLink / NVSwitch peer memory
    (csrc/nvl_allreduce.cu): the gradient arena of every rank is a symmetric-memory allocation
    (torch.distributed._symmetric_memory: the same buffer mapped into every process, with an NVSwitch
    multicast address where the fabric offers one); after the backward pass each rank reduces its 1/W slice
    (`multimem.ld_reduce`, or W peer loads) and broadcasts the sums (`multimem.st`, or W peer stores) between
    two flag barriers in peer memory.  No NCCL kernel competes with the persistent conv kernels for SMs and no
    collective sits inside the captured step graph.

    Use: `alloc = NvlGradSync.allocator(group)`; build the learner with `grad_alloc=alloc`; then
    `learner.grad_sync = NvlGradSync(learner, alloc)`."""

    class _Alloc:
        def __init__(self, group):
            self.group, self.world, self.buf, self.handle = group, dist.get_world_size(group), None, None

        def __call__(self, total, device):
            import torch.distributed._symmetric_memory as symm
            quantum = 4 * self.world
            n = (total + quantum - 1) // quantum * quantum
            gname = self.group.group_name if self.group is not None else dist.group.WORLD.group_name
            self.buf = symm.empty(n, dtype=torch.float32, device=device)
            self.buf.zero_()
            self.flags = symm.empty(8 * 16, dtype=torch.int32, device=device)      # 4 channels x 2 x world <= 16
            self.flags.zero_()
            torch.cuda.synchronize()
            self.handle = symm.rendezvous(self.buf, gname)
            self.flag_handle = symm.rendezvous(self.flags, gname)
            return self.buf

    @staticmethod
    def allocator(group=None):
        return NvlGradSync._Alloc(group)

    # The early exchange starts once layer3's backward is done (94 % of the bytes are final) and runs NEXT TO
    # layer2's backward: those kernels (im2col conv: 320 threads x 168 registers, weight gradient: 192 x 56)
    # leave the ~5 K registers a slim exchange CTA needs on every SM; layer1's halo kernels (384 x 168 = the
    # whole register file) do not, so the exchange must be over before they start.
    EARLY_STAGE = "l3.0.c1"

    def __init__(self, learner, alloc, use_multicast=None, overlap=None):
        """use_multicast: True / False, or None = time both forms once on the real arena and keep the faster one
        (the NVSwitch reduction wins with many ranks, plain peer copies with two).
        overlap (default: VDQN_DDP_OVERLAP != 0): exchange the gradients that are final after layer3's backward
        on a side stream, in slim CTAs (128 threads) that fit on an SM next to a persistent conv CTA, while
        layer2 is being back-propagated; only the last 6 % is exchanged after the backward pass."""
        import ctypes as C
        import os
        from . import _lib as L
        self.L, self.C = L, C
        h, fh = alloc.handle, alloc.flag_handle
        if h is None:
            raise RuntimeError("NvlGradSync: the learner was not built with this allocator (grad_alloc=...)")
        self.world, self.rank = h.world_size, h.rank
        self.buf, self.flags = alloc.buf, alloc.flags            # keep the allocations alive
        if learner.opt.grad_arena.data_ptr() != self.buf.data_ptr():
            raise RuntimeError("NvlGradSync: the gradient arena is not the symmetric allocation")
        self._bufs = (C.c_void_p * self.world)(*[int(p) for p in h.buffer_ptrs])
        self._flagp = (C.c_void_p * self.world)(*[int(p) for p in fh.buffer_ptrs])
        mc = int(h.multicast_ptr) if getattr(h, "has_multicast_support", True) else 0
        self.state = torch.zeros(8, device=self.buf.device, dtype=torch.int32)    # (epoch, block counter) x 4 channels
        self.launched = []
        self.tuning = None
        self.desc = self._desc(0, self.buf.numel(), channel=0)
        if use_multicast is None and mc:
            use_multicast = self._pick(mc, alloc.group)
        self.multicast = bool(mc) and bool(use_multicast)
        self.desc.multicast_ptr = mc if self.multicast else None
        # ---- overlapped form: [split, n) early on a side stream, [0, split) after the backward pass
        if overlap is None:
            overlap = os.environ.get("VDQN_DDP_OVERLAP", "1") != "0"
        self.overlap = False
        names = getattr(learner.model, "_grad_names", None) if hasattr(learner, "model") else None
        if overlap and names is not None:
            base = self.buf.data_ptr()
            first_l3 = (learner.G["resnet.layer3.0.conv1.weight"].data_ptr() - base) // 4
            q = 4 * self.world
            split = (first_l3 + q - 1) // q * q
            if 0 < split < self.buf.numel():
                self.overlap = True
                self.split = split
                self.side = torch.cuda.Stream(priority=-1)
                self.early = self._desc(split, self.buf.numel() - split, channel=1)
                self.early.threads, self.early.max_ctas = 128, int(os.environ.get("VDQN_DDP_EARLY_CTAS", "32"))
                self.early.multicast_ptr = mc or None          # slim form: the switch does the adding
                self.late = self._desc(0, split, channel=2)
                self.late.multicast_ptr = self.desc.multicast_ptr
                self.late.max_ctas = self.desc.max_ctas
        self._early_done = None

    def _desc(self, first, n, channel):
        d = self.L.NvlDesc()
        d.peer_bufs, d.peer_flags = self._bufs, self._flagp
        d.epoch = self.state.data_ptr() + 8 * channel
        d.counter = self.state.data_ptr() + 8 * channel + 4
        d.n, d.first, d.rank, d.world, d.max_ctas, d.channel, d.threads = n, first, self.rank, self.world, 0, channel, 512
        return d

    def _launch(self, d):
        self.L.check(self.L.load().vdqn_nvl_allreduce(self.C.byref(d), self.L.stream_ptr()), "nvl_allreduce")
        self.launched.append((int(d.first), int(d.first + d.n)))

    def _pick(self, mc, group):
        """Time both forms, each at several grid sizes, once on the arena as it is (zeros stay zeros); max over
        ranks, so every rank takes the same decision.  Sets desc.max_ctas; returns whether multicast won."""
        t = {}
        for name, ptr in (("multimem", mc), ("p2p", None)):
            for ctas in (32, 0):
                self.desc.multicast_ptr, self.desc.max_ctas = ptr, ctas
                for _ in range(2):
                    self._launch(self.desc)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                dist.barrier(group)
                e0.record()
                for _ in range(5):
                    self._launch(self.desc)
                e1.record()
                torch.cuda.synchronize()
                x = torch.tensor([e0.elapsed_time(e1) / 5 * 1e3], device=self.buf.device)
                dist.all_reduce(x, op=dist.ReduceOp.MAX, group=group)
                t[f"{name}/{ctas or 'all'}"] = round(float(x.item()), 1)
        self.launched.clear()
        self.tuning = t
        best = min(t, key=t.get)
        self.desc.max_ctas = 0 if best.endswith("all") else int(best.split("/")[1])
        return best.startswith("multimem")

    # the engine reduces the split weight-gradient partials of a stage only when the exchange asks for that
    # stage (one multi-tensor reduction at EARLY_STAGE, one at the end)
    def wants(self, stage: str) -> bool:
        return self.overlap and stage == self.EARLY_STAGE

    def on_stage(self, stage: str):
        if not self.wants(stage):
            return
        ev = torch.cuda.Event()
        ev.record()
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            self._launch(self.early)
            self._early_done = torch.cuda.Event()
            self._early_done.record()

    def finish(self):
        if self.overlap and self._early_done is not None:
            self._launch(self.late)
            torch.cuda.current_stream().wait_event(self._early_done)
            self._early_done = None
        else:
            self._launch(self.desc)

    def close(self):
        torch.cuda.synchronize()
