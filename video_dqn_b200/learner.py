"""The fused Q-learning step: one call = one iteration of the reference loop body
(train_q_network.py:211-229), all on the device, replayed as a CUDA graph.

    model.set_train(); optimizer.zero_grad()
    loss = process_batch(batch)            # 3 forwards + Double-DQN TD loss   (:126-181)
    loss.backward(); optimizer.step()      # (:226-227)
    [target_net.load_state_dict(model.state_dict()) every TARGET_UPDATE_INTERVAL]   (:215-216)

Differences from the reference that do not change results: the `target_net(after)` and
`model(after)` forwards run without saving activations (the reference records and discards their
autograd graphs), the ~24 ATen launches of the TD loss are one kernel, Adam is one kernel, and the
hard target sync is folded into the Adam pass of the step *before* the one the reference performs
it at the top of (same values reach the same forward).  `loss` is returned as a device tensor;
`.item()` stays the caller's choice (the reference syncs every step, :229).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import engine as E
from . import ops
from .optim import FlatArena, FusedAdam, bump_arena_epoch
from .qnet import HabitatDQNMultiAction


@dataclass
class StepConfig:
    """Hot-path hyper-parameters (defaults.py:4-37 overlaid with
    configs/experiments/real_data/config.yml)."""
    GAMMA: float = 0.99
    LOSS_CLIP: str = "rect"
    LINEAR: bool = False
    REMOVE_BEFORE_REWARD: bool = False
    LEARNING_RATE: float = 1e-4
    TARGET_UPDATE_INTERVAL: int = 8000
    double_dqn: bool = True
    # process_batch(compare_ground_truth=config.TRAIN_ON_GROUND_TRUTH) (train_q_network.py:224):
    # regress Q(s, a) onto the loader's discounted ground truth instead of the Bellman target;
    # VALUE_LEARNING masks its NaN entries (:172-176)
    TRAIN_ON_GROUND_TRUTH: bool = False
    VALUE_LEARNING: bool = False
    # the data set then yields the detector SCORES (floating point) as reward and terminal
    # (train_q_network.py:101, dataloaders/q_learning_real.py:76-77) and process_batch uses their
    # `.float()` (:158-160): the step's label buffers are fp32 instead of int64
    CONFIDENCE_REWARD: bool = False

    @classmethod
    def from_config(cls, config):
        """Build from the reference's flattened ExperimentConfig attribute bag."""
        kw = {f: getattr(config, f) for f in ("GAMMA", "LOSS_CLIP", "LINEAR", "REMOVE_BEFORE_REWARD",
                                              "LEARNING_RATE", "TARGET_UPDATE_INTERVAL",
                                              "TRAIN_ON_GROUND_TRUTH", "VALUE_LEARNING", "CONFIDENCE_REWARD")
              if hasattr(config, f)}
        return cls(**kw)


class QLearner:
    def __init__(self, model: HabitatDQNMultiAction, target_net: HabitatDQNMultiAction,
                 cfg: Optional[StepConfig] = None, batch_size: int = 16, *,
                 optimizer: Optional[FusedAdam] = None, frames_uint8: bool = False,
                 use_graph: bool = True, grad_sync=None, world_size: int = 1, one_pass: Optional[bool] = None,
                 grad_alloc=None):
        self.cfg = cfg or StepConfig()
        self.model, self.target_net = model, target_net
        self.B = batch_size
        self.world_size = world_size
        self.grad_sync = grad_sync
        self.use_graph = use_graph
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("QLearner needs the model on a CUDA device (no CPU path)")
        self.device = dev
        model.set_train()
        target_net.eval()
        self.plan = model._state().plan
        F = self.plan.num_frames
        names = model._grad_names
        # ---- parameters / gradients / Adam state / target copy in flat arenas
        self.opt = optimizer or FusedAdam(model.parameters(), lr=self.cfg.LEARNING_RATE, grad_alloc=grad_alloc)
        mp = dict(model.named_parameters())
        self.opt.adopt([mp[n] for n in names])
        self.G: Dict[str, torch.Tensor] = dict(zip(names, self.opt.grad_views()))
        tp = dict(target_net.named_parameters())
        self.t_arena = FlatArena([mp[n].shape for n in names], dev)
        for i, n in enumerate(names):
            v = self.t_arena.view(i)
            v.copy_(tp[n].data)
            tp[n].data = v
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        if self.opt._step:
            self.step_dev.fill_(self.opt._step)
        self.scalars_dev = torch.zeros(2, device=dev, dtype=torch.float32)
        # ---- static inputs
        n = batch_size * F
        self.gt_mode = bool(self.cfg.TRAIN_ON_GROUND_TRUTH)
        nb = 1 if self.gt_mode else 2            # ground-truth regression never looks at s'
        # `before` and `after` are the two halves of one [2B, ...] buffer: the online network runs on
        # both in a single 2B forward (bit-identical to two B forwards: eval-mode BN has no batch
        # statistics)
        if frames_uint8:
            shp = (nb * batch_size, F, 224, 224, 3) if F > 1 else (nb * batch_size, 224, 224, 3)
            self.frames2 = torch.zeros(shp, device=dev, dtype=torch.uint8)
        else:
            shp = (nb * batch_size, F, 3, 224, 224) if F > 1 else (nb * batch_size, 3, 224, 224)
            self.frames2 = torch.zeros(shp, device=dev, dtype=torch.float32)
        self.before = self.frames2[:batch_size]
        self.after = None if self.gt_mode else self.frames2[batch_size:]
        C = self.plan.num_classes
        self.gt = torch.full((batch_size, C), float("nan"), device=dev, dtype=torch.float64)
        self.act = torch.zeros(batch_size, device=dev, dtype=torch.int64)
        ldt = torch.float32 if self.cfg.CONFIDENCE_REWARD else torch.int64
        self.rew = torch.zeros(batch_size, C, device=dev, dtype=ldt)
        self.term = torch.zeros(batch_size, C, device=dev, dtype=ldt)
        self.valid = torch.ones(batch_size, C, device=dev, dtype=ldt)
        # ---- workspaces and step outputs
        # B*F % 64 == 0: online [s ; s'] and target [s'] share ONE 3B forward pass (every conv launch
        # splits its CTAs between the two networks); otherwise a 2B online pass + a B target pass
        self.one_pass = (n % 64 == 0) if one_pass is None else (bool(one_pass) and n % 64 == 0)
        if self.gt_mode:
            self.one_pass = False
            self.ws_train = E.Workspace(self.plan, n, dev, train=True, n_bwd=n)
            self.ws_eval = None
        else:
            self.ws_train = E.Workspace(self.plan, (3 if self.one_pass else 2) * n, dev, train=True, n_bwd=n)
            self.ws_eval = None if self.one_pass else E.Workspace(self.plan, n, dev, train=False)
        self.ws_train.defer_fin = E.DEFER_FINALIZE
        A = self.plan.action_dim
        self.dq = torch.empty(batch_size, C * A, device=dev, dtype=torch.float32)
        self.loss = torch.zeros(1, device=dev, dtype=torch.float32)
        self.best = torch.empty(batch_size, C, device=dev, dtype=torch.int64)
        self.y = torch.empty(batch_size, C, device=dev, dtype=torch.float32)
        # every step's loss also goes to a small pinned ring with an event per slot, so a training loop
        # can read step k's loss after it has launched step k+1 (`loss_value(k)`): the device never
        # waits for the host between steps, which `loss.item()` right after `step()` makes it do
        self._loss_ring = torch.zeros(4, dtype=torch.float32, pin_memory=True)
        self._loss_events = [torch.cuda.Event() for _ in range(4)]
        self.steps_done = 0
        self.sample_number = 0
        self._graphs: Dict[bool, torch.cuda.CUDAGraph] = {}
        self._eager_steps = 0
        self.kernels_per_step = 0

    # ------------------------------------------------------------------ the step body
    def _frames(self, t):
        return t.view(-1, *t.shape[-3:])

    def _enqueue(self, sync_target: bool):
        cfg, plan = self.cfg, self.plan
        st, tt = self.model._eng, self.target_net._eng
        C, A = plan.num_classes, plan.action_dim
        ops.zero_(self.opt.grad_arena)
        ops.zero_(self.loss)
        # online net on [s ; s'] in one 2B forward (activations of the s half feed the backward),
        # target net on s' (nothing kept)
        B = self.B
        wt = self.ws_train
        if self.gt_mode:
            E.forward(plan, st.W, st.P, wt, self._frames(self.frames2))
            ops.td_epilogue(wt.q.view(B, C, A), None, None, self.act, None, None, gt=self.gt,
                            value_learning=cfg.VALUE_LEARNING, inv_count=1.0 / (B * C),
                            dq=self.dq.view(B, C, A), loss=self.loss)
        elif self.one_pass:
            n = B * plan.num_frames
            ops.stem_pack(self._frames(self.frames2), wt.xp[:2 * n])
            # the target net sees s' too: its range [2n, 3n) of the stem re-reads frames [n, 2n)
            E.forward_packed(plan, st.W, st.P, wt, W2=tt.W, P2=tt.P, split=2 * n, x_alias=(2 * n, n))
            q_nt = wt.q[2 * B:3 * B]
        else:
            E.forward(plan, st.W, st.P, wt, self._frames(self.frames2))
            E.forward(plan, tt.W, tt.P, self.ws_eval, self._frames(self.after))
            q_nt = self.ws_eval.q
        if not self.gt_mode:
            ops.td_epilogue(wt.q[:B].view(B, C, A), wt.q[B:2 * B].view(B, C, A),
                            q_nt.view(B, C, A), self.act, self.rew, self.term, self.valid,
                            gamma=cfg.GAMMA, double_dqn=cfg.double_dqn, clip_rect=(cfg.LOSS_CLIP == "rect"),
                            linear=cfg.LINEAR, use_valid=cfg.REMOVE_BEFORE_REWARD,
                            inv_count=1.0 / (B * C), dq=self.dq.view(B, C, A), loss=self.loss,
                            best=self.best, y=self.y)
        sync = self.grad_sync
        E.backward(plan, st.W, st.P, self.G, self.ws_train, self.dq,
                   on_grads_ready=(sync.on_stage if sync is not None else None))
        if sync is not None:
            sync.finish()
        self.opt.step(grads_in_arena=True, step_dev=self.step_dev, scalars_dev=self.scalars_dev,
                      grad_scale=1.0 / self.world_size,
                      target_arena=self.t_arena.flat if sync_target else None)
        st.W.prepare(st.P)
        if sync_target:
            tb, mb = dict(self.target_net.named_buffers()), dict(self.model.named_buffers())
            for k, v in tb.items():
                v.copy_(mb[k])
            tt.W.prepare(tt.P)

    # ------------------------------------------------------------------ public API
    def load_batch(self, batch, non_blocking: bool = True):
        """Copy a reference-format batch (before, after, act, rew, term, gt, valid_mask)
        (dataloaders/q_learning_real.py:98) into the static device buffers."""
        before, after, act, rew, term, gt, valid = batch
        if before.shape[0] != self.B:
            raise ValueError("bad shape")
        self.before.copy_(before.view(self.before.shape), non_blocking=non_blocking)
        if self.gt_mode:
            self.gt.copy_(torch.as_tensor(gt, dtype=torch.float64).view(self.gt.shape), non_blocking=non_blocking)
            self.act.copy_(act.view(-1), non_blocking=non_blocking)
            return
        check_label_dtypes(self, rew, term)
        self.after.copy_(after.view(self.after.shape), non_blocking=non_blocking)
        self.act.copy_(act.view(-1), non_blocking=non_blocking)
        self.rew.copy_(rew, non_blocking=non_blocking)
        self.term.copy_(term, non_blocking=non_blocking)
        self.valid.copy_(valid, non_blocking=non_blocking)

    def step(self, batch=None) -> torch.Tensor:
        """One training iteration; returns the loss as a 1-element device tensor."""
        if batch is not None:
            self.load_batch(batch)
        self.model._state(); self.target_net._state()        # refresh bf16 operands if stale
        self.sample_number += 1
        sync_target = (self.sample_number + 1) % self.cfg.TARGET_UPDATE_INTERVAL == 0
        if self.use_graph and self._eager_steps >= 1:
            g = self._graphs.get(sync_target)
            if g is None:
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    self._enqueue(sync_target)
                self._graphs[sync_target] = g
                # capture does not execute: undo the bookkeeping _enqueue did on the host
                self.opt._step -= 1
            g.replay()
            self.opt.note_graph_replay(sync_target, self.t_arena.flat)
        else:
            self._enqueue(sync_target)
            self._eager_steps += 1
        slot = self.steps_done % 4
        self._loss_ring[slot:slot + 1].copy_(self.loss, non_blocking=True)
        self._loss_events[slot].record()
        self.steps_done += 1
        # the bf16 operands were just re-derived inside the step: mark the modules in sync
        self.model._eng.sig = None
        self.target_net._eng.sig = None
        self._mark_clean()
        return self.loss

    def loss_value(self, step_index: int = -1) -> float:
        """Loss of step `step_index` (0-based count of `step()` calls; default: the latest) as a host
        float, from the pinned ring: waits for that step only.  At most the last four steps are kept."""
        k = self.steps_done - 1 if step_index < 0 else step_index
        if not (self.steps_done - 4 <= k < self.steps_done) or k < 0:
            raise ValueError("loss_value: step not in the ring")
        self._loss_events[k % 4].synchronize()
        return float(self._loss_ring[k % 4])

    def _mark_clean(self):
        from .optim import arena_epoch
        for m in (self.model, self.target_net):
            nt = m._named_tensors()
            m._eng.P = nt
            m._eng.sig = tuple((t.data_ptr(), t._version, arena_epoch(t)) for t in nt.values())

    def sync_target_now(self):
        """target_net.load_state_dict(model.state_dict()) (train_q_network.py:121,208)."""
        self.t_arena.flat.copy_(self.opt.param_arena)
        bump_arena_epoch(self.t_arena.flat)
        tb, mb = dict(self.target_net.named_buffers()), dict(self.model.named_buffers())
        for k, v in tb.items():
            v.copy_(mb[k])

    # ------------------------------------------------------------------ checkpoint / resume
    def checkpoint(self) -> dict:
        """The dictionary the reference saves every CHECKPOINT_INTERVAL steps (train_q_network.py:241-247):
        `sample_number`, the model's 250-key `state_dict()` and the optimizer's `state_dict()` in
        torch.optim.Adam's layout -- loadable by the reference's own resume code (:192-198) and by
        `load_model_number` (:50-57)."""
        torch.cuda.current_stream().synchronize()
        return {"sample_number": self.sample_number,
                "model_state_dict": {k: v.detach().cpu().clone() for k, v in self.model.state_dict().items()},
                "optimizer_state_dict": _to_cpu(self.opt.state_dict())}

    def save_checkpoint(self, path: str):
        """`torch.save({...}, f'{config.folder}/models/sample{sample_number}.torch')` (:241-247)"""
        torch.save(self.checkpoint(), path)

    def resume(self, snapshot, resume_from: Optional[int] = None):
        """The reference's resume sequence (train_q_network.py:190-208) on a `sample{n}.torch` snapshot
        (path or loaded dict, written by the reference or by `save_checkpoint`): load the model and
        optimizer state, `sample_number = resume_from + 1` (:190; default: the snapshot's own number),
        then `target_net.load_state_dict(model.state_dict())` (:208)."""
        if isinstance(snapshot, (str, bytes)) or hasattr(snapshot, "__fspath__"):
            snapshot = torch.load(snapshot, map_location="cpu")
        self.model.load_state_dict(snapshot["model_state_dict"])           # writes into the parameter arena
        self.opt.load_state_dict(snapshot["optimizer_state_dict"])
        self.step_dev.fill_(self.opt._step)
        n = snapshot.get("sample_number", -1) if resume_from is None else resume_from
        self.sample_number = int(n) + 1
        bump_arena_epoch(self.opt.param_arena)
        self.sync_target_now()
        self.model.set_train()
        self.target_net.eval()
        self.model._state(); self.target_net._state()                      # bf16 operands of both networks


def check_label_dtypes(learner, rew, term):
    """A floating-point reward / terminal (CONFIDENCE_REWARD data) copied into the int64 label buffers of a
    learner built without that switch would silently become 0: refuse."""
    if (rew.is_floating_point() or term.is_floating_point()) and not learner.rew.is_floating_point():
        raise ValueError("floating-point rewards / terminals (CONFIDENCE_REWARD data set) need a learner built "
                         "with StepConfig(CONFIDENCE_REWARD=True): its label buffers are int64")


def _to_cpu(obj):
    if torch.is_tensor(obj):
        return obj.detach().cpu().clone()
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj
