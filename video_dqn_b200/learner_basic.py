"""Training step of the `basic` architecture (ARCHITECTURE != 'extra_capacity').

`HabitatDQNMultiAction(..., extra_capacity=False)` is trunk + AdaptiveAvgPool2d(1) + one Linear
(archs/HabitatDQNMultiAction.py:32-34).  `set_train()` only freezes the trunk for extra_capacity
(:37-40), so here the 20 trunk BatchNorms run in TRAIN mode inside the loop body of
train_q_network.py:211-229:

    model(before)       batch statistics of `before`, running statistics updated, graph kept    (:131)
    target_net(after)   eval mode (running statistics of the target copy)                        (:122,140)
    model(after)        batch statistics of `after`, running statistics updated AGAIN            (:142)
    TD loss, backward through model(before) with the batch-statistics BatchNorm backward, Adam

Batch statistics depend on the conv OUTPUT, so BatchNorm cannot be folded into the conv weights as on
the shipped (eval-mode) path: every conv runs on the tensor-core kernels with plain bf16 weights and
writes its raw output; `bn_stats -> bn_finalize -> bn_apply` (csrc/batchnorm.cu, HBM-bound) normalise
it (+ residual, + ReLU); the backward inserts `bn_bwd_reduce -> bn_bwd_apply` between the data
gradient of the layer above and the weight / data gradient of the conv below.

F frames (PANORAMA / PREVIOUS_IMAGES, train_q_network.py:36-47): the reference pushes each frame through the
trunk separately, in order (archs/HabitatDQNMultiAction.py:49-51) -- F sets of batch statistics and F
running-statistics updates per forward -- and concatenates the pooled features before the Linear.  The same
here: one train-mode trunk pass (and one workspace of saved activations) per frame, the trunk's parameter
gradients are the sum over the frames.

Data parallel (the reference has no such mode, :275): SyncBatchNorm -- the fp64 per-channel sums of every
BatchNorm are all-reduced over the ranks (ops.BnSync), forward and backward, so each rank normalises with the
statistics of the global batch -- and the gradient arena is all-reduced before Adam (`grad_scale = 1/world`).
Eager launches, no CUDA graph.  CUDA only, no fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import engine as E
from . import ops
from .learner import StepConfig
from .optim import FusedAdam
from .qnet import HabitatDQNMultiAction

bf16 = torch.bfloat16


def grad_param_names_basic() -> List[str]:
    """The 62 tensors that receive gradients, `model.parameters()` order (resnet.fc.* never get one)."""
    from .qnet import grad_param_names
    return [n for n in grad_param_names() if n.startswith("resnet.")] + ["top.weight", "top.bias"]


class BnTrainWorkspace:
    """Raw conv outputs, normalised activations and per-BatchNorm batch statistics of one train-mode
    forward over `n` frames; with `train` also the gradient buffers of its backward."""

    def __init__(self, plan: E.NetPlan, n: int, device, train: bool):
        self.n, self.train = n, train
        e = lambda *s, dt=bf16: torch.empty(*s, device=device, dtype=dt)  # noqa: E731
        z = lambda *s, dt=bf16: torch.zeros(*s, device=device, dtype=dt)  # noqa: E731
        self.xp = e(n, 112, 112, 16)
        self.s_raw, self.s = e(n, 112, 112, 64), e(n, 112, 112, 64)
        self.idx = e(n, 56, 56, 64, dt=torch.uint8) if train else None
        self.p = e(n, 56, 56, 64)
        self.c1_raw, self.a1, self.c2_raw, self.ds_raw, self.idn, self.out = [], [], [], [], [], []
        for b in plan.blocks:
            shp = (n, b.out_hw, b.out_hw, b.cout)
            self.c1_raw.append(e(*shp)); self.a1.append(e(*shp)); self.c2_raw.append(e(*shp)); self.out.append(e(*shp))
            self.ds_raw.append(e(*shp) if b.ds is not None else None)
            self.idn.append(e(*shp) if b.ds is not None else None)
        self.stats: Dict[str, ops.BnBatchStats] = {c.bn: ops.BnBatchStats(c.cout, device) for c in plan.convs
                                                   if c.bn is not None}
        f32 = torch.float32
        self.pooled = e(n, 512, dt=f32)
        self.q = e(n, plan.num_classes * plan.action_dim, dt=f32)
        if train:
            self.dpooled = e(n, 512, dt=f32)
            self.dy_out = {b.out_hw: (e(n, b.out_hw, b.out_hw, b.cout), e(n, b.out_hw, b.out_hw, b.cout))
                           for b in plan.blocks}
            self.dy_a1 = {b.out_hw: e(n, b.out_hw, b.out_hw, b.cout) for b in plan.blocks}
            self.d_c2 = {b.out_hw: e(n, b.out_hw, b.out_hw, b.cout) for b in plan.blocks}
            self.d_ds = {b.out_hw: e(n, b.out_hw, b.out_hw, b.cout) for b in plan.blocks if b.ds is not None}
            self.r_dil = {b.out_hw: z(n, b.in_hw, b.in_hw, b.cin) for b in plan.blocks if b.stride == 2}
            self.dy_p = e(n, 56, 56, 64)
            self.dy_s = e(n, 112, 112, 64)
            convs = [c for c in plan.convs if c.name != "head"]
            self.part = e(max(E.wgrad_splits(c, n) * c.cout * c.K for c in convs), dt=f32)


def _bn(ws: BnTrainWorkspace, P, c: E.ConvSpec, x_raw, y, *, residual=None, relu=False, update_running=True, sync=None):
    p = c.bn
    return ops.bn_train_fwd(x_raw, ws.stats[p], P[p + ".weight"], P[p + ".bias"], P[p + ".running_mean"],
                            P[p + ".running_var"], P.get(p + ".num_batches_tracked"), y, residual=residual,
                            relu=relu, momentum=0.1, eps=E.BN_EPS, update_running=update_running, sync=sync)


def forward_bn_train(plan: E.NetPlan, W: E.PreparedWeights, P: Dict[str, torch.Tensor], ws: BnTrainWorkspace,
                     update_running: bool = True, sync=None) -> torch.Tensor:
    """Train-mode forward of a single-frame network from the packed input ws.xp -> Q [n, classes*actions] fp32."""
    trunk_forward_bn_train(plan, W, P, ws, update_running, sync)
    ops.linear_fwd(ws.pooled, P["top.weight"], P["top.bias"], False, ws.q)
    return ws.q


def trunk_forward_bn_train(plan: E.NetPlan, W: E.PreparedWeights, P: Dict[str, torch.Tensor], ws: BnTrainWorkspace,
                           update_running: bool = True, sync=None) -> torch.Tensor:
    """Train-mode trunk + global average pool from the packed input ws.xp -> ws.pooled [n, 512] fp32.  `W` must
    hold UNFOLDED weights (PreparedWeights(fold_bn=False))."""
    assert not W.fold_bn
    # shift = 0 for unfolded weights; it is passed so the launches take the same (tested) epilogue variants
    # as the folded path
    conv = lambda c, x, out: ops.conv_gemm(x, W.w_fwd[c.name], c.stride, c.pad_lo, c.pad_hi,  # noqa: E731
                                           shift=W.shift[c.name], out=out)
    kw = dict(update_running=update_running, sync=sync)
    conv(plan.stem, ws.xp, ws.s_raw)
    _bn(ws, P, plan.stem, ws.s_raw, ws.s, relu=True, **kw)
    ops.maxpool_fwd(ws.s, ws.p, ws.idx)
    x = ws.p
    for i, b in enumerate(plan.blocks):
        conv(b.conv1, x, ws.c1_raw[i])
        _bn(ws, P, b.conv1, ws.c1_raw[i], ws.a1[i], relu=True, **kw)
        conv(b.conv2, ws.a1[i], ws.c2_raw[i])
        idn = x
        if b.ds is not None:
            conv(b.ds, x, ws.ds_raw[i])
            idn = _bn(ws, P, b.ds, ws.ds_raw[i], ws.idn[i], **kw)
        # torchvision BasicBlock: bn2 before the downsample branch's BatchNorm (resnet.py:89-105) -- the
        # order only matters for nothing observable (independent layers), kept for readability
        _bn(ws, P, b.conv2, ws.c2_raw[i], ws.out[i], residual=idn, relu=True, **kw)
        x = ws.out[i]
    ops.avgpool_fwd(x, ws.pooled)
    return ws.pooled


def _wgrad_raw(ws: BnTrainWorkspace, P, G, c: E.ConvSpec, x, dy):
    splits = E.wgrad_splits(c, ws.n)
    part = ws.part[: splits * c.cout * c.K]
    ops.conv_wgrad(x, dy, c.k, c.k, c.stride, c.pad_lo, c.pad_hi, splits=splits, part=part,
                   algo=2 if E.wgrad_uses_halo(c) else 0)
    ops.wgrad_finalize(part, P[c.wkey], G[c.wkey], splits=splits, Cout=c.cout, Cin=c.gemm_cin, R=c.k, S=c.k,
                       K=c.K, kmap=c.kmap)


def _bn_bwd(ws: BnTrainWorkspace, P, G, c: E.ConvSpec, dy, x_raw, dx, sync=None):
    p = c.bn
    return ops.bn_train_bwd(dy, x_raw, ws.stats[p], P[p + ".weight"], G[p + ".weight"], G[p + ".bias"], dx, sync=sync)


def backward_bn_train(plan: E.NetPlan, W: E.PreparedWeights, P, G, ws: BnTrainWorkspace, dq: torch.Tensor, sync=None):
    """Single-frame network: dq [n, classes*actions] fp32 (overwritten) -> every parameter gradient in G."""
    ops.linear_bwd(ws.pooled, P["top.weight"], None, dq, G["top.weight"], G["top.bias"], False, dx=ws.dpooled)
    trunk_backward_bn_train(plan, W, P, G, ws, sync)


def trunk_backward_bn_train(plan: E.NetPlan, W: E.PreparedWeights, P, G, ws: BnTrainWorkspace, sync=None):
    """ws.dpooled [n, 512] fp32 -> the trunk's parameter gradients in G (overwritten)."""
    _bnb = lambda *a: _bn_bwd(*a, sync=sync)  # noqa: E731
    last = plan.blocks[-1]
    ci = 0
    cur = ws.dy_out[last.out_hw][ci]
    ops.avgpool_bwd(ws.dpooled, ws.out[-1], cur)           # d loss / d (block sum), masked by the block's ReLU
    for i in range(len(plan.blocks) - 1, -1, -1):
        b = plan.blocks[i]
        x_in = ws.out[i - 1] if i > 0 else ws.p
        prev = plan.blocks[i - 1] if i > 0 else None
        # bn2 -> conv2
        d_c2 = _bnb(ws, P, G, b.conv2, cur, ws.c2_raw[i], ws.d_c2[b.out_hw])
        _wgrad_raw(ws, P, G, b.conv2, ws.a1[i], d_c2)
        dy_a1 = ws.dy_a1[b.out_hw]
        ops.conv_gemm(d_c2, W.w_dgrad[b.conv2.name], 1, 1, 1, mask_src=ws.a1[i], out=dy_a1,
                      tile_n=E._dgrad_tile_n(b.cout))
        # bn1 (in place: dy_a1 becomes the gradient w.r.t. conv1's raw output)
        d_c1 = _bnb(ws, P, G, b.conv1, dy_a1, ws.c1_raw[i], dy_a1)
        # identity / downsample branch
        if b.ds is not None:
            d_ds = _bnb(ws, P, G, b.ds, cur, ws.ds_raw[i], ws.d_ds[b.out_hw])
            _wgrad_raw(ws, P, G, b.ds, x_in, d_ds)
            res = ws.r_dil[b.out_hw]
            ops.conv_gemm(d_ds, W.w_dgrad[b.ds.name], 1, 0, 0, out=res, out_scatter=2,
                          tile_n=E._scatter_tile_n(b.cin))
        else:
            res = cur
        _wgrad_raw(ws, P, G, b.conv1, x_in, d_c1)
        if prev is not None:
            ni = (1 - ci) if prev.out_hw == b.out_hw else 0
            dst, mask = ws.dy_out[prev.out_hw][ni], x_in
        else:
            ni, dst, mask = 0, ws.dy_p, None
        if b.stride == 2:
            for pa, pb, wf in E.parity_filters(W, b.conv1):
                ops.conv_gemm(d_c1, wf, 1, 0, wf.shape[1] - 1, pad_hi_w=wf.shape[2] - 1, residual=res,
                              mask_src=mask, out=dst, out_scatter=2, scatter_off=(pa, pb),
                              scatter_inputs=True, tile_n=E._scatter_tile_n(b.cin))
        else:
            ops.conv_gemm(d_c1, W.w_dgrad[b.conv1.name], 1, 1, 1, residual=res, mask_src=mask, out=dst,
                          tile_n=E._dgrad_tile_n(b.cin))
        cur, ci = dst, ni
    # max-pool (routes through the arg-max and applies the stem ReLU mask), stem BatchNorm, stem conv
    ops.maxpool_bwd(ws.dy_p, ws.idx, ws.p, ws.dy_s, colsum=None)
    _bnb(ws, P, G, plan.stem, ws.dy_s, ws.s_raw, ws.dy_s)
    _wgrad_raw(ws, P, G, plan.stem, ws.xp, ws.dy_s)


class BasicQLearner:
    """`learner.step(batch)` = one iteration of train_q_network.py:211-229 for the `basic` architecture
    (see the module docstring); returns the loss as a 1-element device tensor.  `batch` is the loader's
    7-tuple (dataloaders/q_learning_real.py:98).  The hard target sync of :215-216 happens at the top of
    the step, as in the reference."""

    def __init__(self, model: HabitatDQNMultiAction, target_net: HabitatDQNMultiAction,
                 cfg: Optional[StepConfig] = None, batch_size: int = 16, *,
                 optimizer: Optional[FusedAdam] = None, world_size: int = 1, process_group=None,
                 use_graph: bool = False):
        """world_size > 1: one process per GPU over torch.distributed (`process_group`, default the world):
        SyncBatchNorm statistics + a gradient all-reduce before Adam.
        use_graph: after one eager step the ~330 launches of a step are captured in a CUDA graph and replayed
        (single process only; batches are copied into static device buffers first)."""
        if use_graph and world_size > 1:
            raise ValueError("BasicQLearner: use_graph is for a single process (the SyncBatchNorm exchanges are not captured)")
        if model.extra_capacity or target_net.extra_capacity:
            raise ValueError("BasicQLearner is for extra_capacity=False; use QLearner for the shipped architecture")
        self.cfg = cfg or StepConfig()
        if self.cfg.TRAIN_ON_GROUND_TRUTH:
            raise NotImplementedError("ground-truth regression is implemented for the extra_capacity path only")
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("BasicQLearner needs the model on a CUDA device (no CPU path)")
        self.model, self.target_net, self.B, self.device = model, target_net, batch_size, dev
        model.set_train()                                   # BatchNorms stay in train mode (:37-40)
        target_net.eval()
        self.F = F = model.num_frames
        self.world, self.group = world_size, process_group
        self.sync = ops.BnSync(world_size, process_group) if world_size > 1 else None
        self.plan = E.make_plan(model.action_dim, model.num_classes, 1)          # per-frame trunk plan
        self.names = grad_param_names_basic()
        self.opt = optimizer or FusedAdam(model.parameters(), lr=self.cfg.LEARNING_RATE)
        mp = dict(model.named_parameters())
        self.opt.adopt([mp[n] for n in self.names])
        self.G: Dict[str, torch.Tensor] = dict(zip(self.names, self.opt.grad_views()))
        self.W = E.PreparedWeights(self.plan, dev, trunk_only=True, fold_bn=False)
        # model(before): one workspace of saved activations per frame; model(after): one, reused frame by frame
        self.ws_frames = [BnTrainWorkspace(self.plan, batch_size, dev, train=True) for _ in range(F)]
        self.ws_s = self.ws_frames[0]
        self.ws_next = BnTrainWorkspace(self.plan, batch_size, dev, train=False)
        if F > 1:
            from .optim import FlatArena
            self._tmp = FlatArena([mp[n].shape for n in self.names], dev)       # gradients of frames 1 .. F-1
            self.G_tmp = dict(zip(self.names, self._tmp.views()))
        f32 = torch.float32
        nq = self.plan.num_classes * self.plan.action_dim
        self.pooled_s = torch.empty(batch_size, 512 * F, device=dev, dtype=f32)
        self.pooled_n = torch.empty(batch_size, 512 * F, device=dev, dtype=f32)
        self.dpooled = torch.empty(batch_size, 512 * F, device=dev, dtype=f32)
        self.q_s = torch.empty(batch_size, nq, device=dev, dtype=f32)
        self.q_no = torch.empty(batch_size, nq, device=dev, dtype=f32)
        C, A = self.plan.num_classes, self.plan.action_dim
        self.dq = torch.empty(batch_size, C * A, device=dev, dtype=torch.float32)
        self.loss = torch.zeros(1, device=dev, dtype=torch.float32)
        self.best = torch.empty(batch_size, C, device=dev, dtype=torch.int64)
        self.y = torch.empty(batch_size, C, device=dev, dtype=torch.float32)
        self.sample_number = 0
        self.q_next_target = None
        self.use_graph = use_graph
        self._graph = None
        self._eager_steps = 0
        self._in = None                          # static input buffers (graph mode)
        # Adam's step count lives on the device so that a captured step replays (as in learner.QLearner)
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self.step_dev.fill_(self.opt._step)
        self.scalars_dev = torch.zeros(2, device=dev, dtype=torch.float32)

    def _tensors(self) -> Dict[str, torch.Tensor]:
        d = {k: v.detach() for k, v in self.model.named_parameters()}
        d.update(dict(self.model.named_buffers()))
        return d

    @torch.no_grad()
    def step(self, batch) -> torch.Tensor:
        batch = [t.to(self.device, non_blocking=True) if torch.is_tensor(t) else t for t in batch]
        before, after = batch[0], batch[1]
        if before.shape[0] != self.B or after.shape != before.shape:
            raise ValueError("bad shape")
        self.sample_number += 1
        if self.sample_number % self.cfg.TARGET_UPDATE_INTERVAL == 0:
            self.target_net.load_state_dict(self.model.state_dict())          # :215-216
        if not self.use_graph:
            self._enqueue(batch)
        else:
            if self._in is None:
                self._in = [t.clone() if torch.is_tensor(t) else t for t in batch]
            else:
                for dst, src in zip(self._in, batch):
                    if torch.is_tensor(dst):
                        if dst.shape != src.shape or dst.dtype != src.dtype:
                            raise ValueError("BasicQLearner(use_graph=True): batches must keep their shape and dtype")
                        dst.copy_(src, non_blocking=True)
            # host-side state the captured work relies on: the target module's folded operands follow its
            # parameters (a hard sync above changes them) -- re-derived here, outside the graph
            self.target_net._basic_state()
            if self._eager_steps < 1:
                self._enqueue(self._in)
                self._eager_steps += 1
            else:
                if self._graph is None:
                    g = torch.cuda.CUDAGraph()
                    torch.cuda.synchronize()
                    with torch.cuda.graph(g):
                        self._enqueue(self._in)
                    self._graph = g
                    self.opt._step -= 1           # capture does not execute: undo the host bookkeeping
                self._graph.replay()
                self.opt.note_graph_replay()
        # parameters and running statistics changed behind torch's version counters: make the module's
        # forward-only path re-derive its folded operands next time it is used
        st = getattr(self.model, "_basic_eng", None)
        if st is not None:
            st["sig"] = None
        return self.loss

    def _enqueue(self, batch):
        """every launch of one step (the body that is captured)"""
        before, after, act, rew, term, _gt, valid = batch
        B, cfg, plan = self.B, self.cfg, self.plan
        P = self._tensors()
        self.W.prepare(P)                                    # plain bf16 operands of the current weights
        C, A = plan.num_classes, plan.action_dim
        self.opt.grad_arena.zero_()
        self.loss.zero_()
        F, sync = self.F, self.sync
        frames = lambda t, f: (t[:, f] if t.dim() == 5 else t).contiguous()  # noqa: E731
        # model(before): every frame through the trunk on its own (its own batch statistics, one
        # running-statistics update each, in frame order), activations kept
        for f, ws in enumerate(self.ws_frames):
            ops.stem_pack(frames(before, f), ws.xp)
            trunk_forward_bn_train(plan, self.W, P, ws, sync=sync)
            if F > 1:
                self.pooled_s[:, 512 * f:512 * (f + 1)].copy_(ws.pooled)
        pooled_s = self.pooled_s if F > 1 else self.ws_s.pooled
        q_s = ops.linear_fwd(pooled_s, P["top.weight"], P["top.bias"], False, self.q_s)
        # target_net(after): eval mode, the module's own forward-only path
        q_nt = self.target_net(after).reshape(B, C, A).contiguous()
        # model(after): train mode again (F more running-statistics updates), nothing kept
        for f in range(F):
            ops.stem_pack(frames(after, f), self.ws_next.xp)
            trunk_forward_bn_train(plan, self.W, P, self.ws_next, sync=sync)
            if F > 1:
                self.pooled_n[:, 512 * f:512 * (f + 1)].copy_(self.ws_next.pooled)
        q_no = ops.linear_fwd(self.pooled_n if F > 1 else self.ws_next.pooled, P["top.weight"], P["top.bias"], False,
                              self.q_no)
        ops.td_epilogue(q_s.view(B, C, A), q_no.view(B, C, A), q_nt, act.view(-1), rew, term, valid,
                        gamma=cfg.GAMMA, double_dqn=cfg.double_dqn, clip_rect=(cfg.LOSS_CLIP == "rect"),
                        linear=cfg.LINEAR, use_valid=cfg.REMOVE_BEFORE_REWARD, inv_count=1.0 / (B * C),
                        dq=self.dq.view(B, C, A), loss=self.loss, best=self.best, y=self.y)
        self.q_next_target = q_nt
        # backward: the Linear once, then the trunk per frame (the trunk's parameter gradients add up)
        ops.linear_bwd(pooled_s, P["top.weight"], None, self.dq, self.G["top.weight"], self.G["top.bias"], False,
                       dx=self.dpooled if F > 1 else self.ws_s.dpooled)
        for f, ws in enumerate(self.ws_frames):
            if F > 1:
                ws.dpooled.copy_(self.dpooled[:, 512 * f:512 * (f + 1)])
            if f == 0:
                trunk_backward_bn_train(plan, self.W, P, self.G, ws, sync)
            else:
                self._tmp.flat.zero_()
                trunk_backward_bn_train(plan, self.W, P, self.G_tmp, ws, sync)
                self.opt.grad_arena.add_(self._tmp.flat)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.opt.grad_arena, group=self.group)
        self.opt.step(grads_in_arena=True, grad_scale=1.0 / self.world, step_dev=self.step_dev,
                      scalars_dev=self.scalars_dev)

    # ------------------------------------------------------------------ snapshots (train_q_network.py:190-208,241-247)
    def checkpoint(self) -> dict:
        from .learner import _to_cpu
        torch.cuda.current_stream().synchronize()
        return {"sample_number": self.sample_number,
                "model_state_dict": {k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()},
                "optimizer_state_dict": _to_cpu(self.opt.state_dict())}

    def save_checkpoint(self, path: str):
        torch.save(self.checkpoint(), path)

    def resume(self, snapshot, resume_from: Optional[int] = None):
        if isinstance(snapshot, (str, bytes)) or hasattr(snapshot, "__fspath__"):
            snapshot = torch.load(snapshot, map_location="cpu")
        self.model.load_state_dict(snapshot["model_state_dict"])           # parameters (arena views) and BN buffers
        self.opt.load_state_dict(snapshot["optimizer_state_dict"])
        n = snapshot.get("sample_number", -1) if resume_from is None else resume_from
        self.sample_number = int(n) + 1                                      # :190
        self.target_net.load_state_dict(self.model.state_dict())            # :208
        self.model.set_train()
        self.target_net.eval()
        self.step_dev.fill_(self.opt._step)
