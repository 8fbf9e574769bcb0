"""Golden vectors for every branch of the reference's TD loss (train_q_network.py:126-181).

TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference).  The closure
``process_batch`` is lifted out of ``train_q_network.py`` unmodified (see make_goldens.py) and run with
stub networks that return fixed Q tensors, so the fixture isolates the loss arithmetic:
Double-DQN / plain, LINEAR, LOSS_CLIP on/off, REMOVE_BEFORE_REWARD masks, and the ground-truth
branches (``compare_ground_truth=True`` = TRAIN_ON_GROUND_TRUTH, with and without VALUE_LEARNING and
its NaN mask).  Stored per case: inputs, loss, dLoss/dQ(s).  The oracle must reproduce them here.

usage:  python -m oracle.make_td_goldens [--out tests/golden]
"""
from __future__ import annotations

import argparse
import os
import types

import numpy as np
import torch

from . import qstep
from .make_goldens import lift_process_batch

CASES = [
    # name, A, config overrides, ground-truth mode
    ("double_rect", 3, {}, False),
    ("double_noclip", 3, {"LOSS_CLIP": "none"}, False),
    ("linear_rect", 3, {"LINEAR": True}, False),
    ("masked", 3, {"REMOVE_BEFORE_REWARD": True}, False),
    ("gt_value_learning", 1, {"VALUE_LEARNING": True}, True),
    ("gt_plain", 3, {"VALUE_LEARNING": False}, True),
    # CONFIDENCE_REWARD (train_q_network.py:101): reward = terminal = the float64 detector scores
    # (dataloaders/q_learning_real.py:76-77); process_batch takes their .float() (:158-160)
    ("confidence_reward", 3, {"REMOVE_BEFORE_REWARD": True, "_float_labels": True}, False),
]


def build_case(name, A, over, gt_mode, seed):
    g = torch.Generator().manual_seed(seed)
    B, C = 6, qstep.NUM_CLASSES
    q_s = torch.randn(B, C, A, generator=g) * 0.6
    q_no = torch.randn(B, C, A, generator=g) * 0.6
    q_nt = torch.randn(B, C, A, generator=g) * 0.6
    if A > 1:
        q_no[0, 0, 1] = q_no[0, 0, 0] = q_no[0, 0].max() + 1.0      # an exact tie: first index must win
    act = torch.randint(0, A, (B,), generator=g)
    rew = (torch.rand(B, C, generator=g) < 0.3).long()
    over = dict(over)
    if over.pop("_float_labels", False):
        rew = torch.rand(B, C, generator=g, dtype=torch.float64)
    valid = (torch.rand(B, C, generator=g) < 0.7).long()
    if gt_mode:
        steps = torch.randint(0, 30, (B, C), generator=g).double()
        gt = torch.pow(torch.full((B, C), 0.99, dtype=torch.float64), steps)      # q_learning_real.py:86-89
        if over.get("VALUE_LEARNING"):
            gt[torch.rand(B, C, generator=g) < 0.35] = float("nan")
    else:
        gt = torch.full((B,), float("nan"), dtype=torch.float64)
    cfg = dict(GAMMA=0.99, LOSS_CLIP="rect", LINEAR=False, REMOVE_BEFORE_REWARD=False, VALUE_LEARNING=False,
               device="cpu")
    cfg.update(over)
    return dict(q_s=q_s, q_no=q_no, q_nt=q_nt, act=act, rew=rew, valid=valid, gt=gt, cfg=cfg)


def run_reference(c, gt_mode):
    q_s = c["q_s"].clone().requires_grad_(True)
    marker_before, marker_after = torch.zeros(1), torch.ones(1)
    model = lambda x: q_s if x is marker_before else c["q_no"]       # noqa: E731
    target = lambda x: c["q_nt"]                                        # noqa: E731
    pb = lift_process_batch(model, target, types.SimpleNamespace(**c["cfg"]))

    class _Id:                   # process_batch calls x.to(device) on every batch element
        def __init__(self, t): self.t = t
        def to(self, _): return self.t
    batch = [_Id(marker_before), _Id(marker_after), _Id(c["act"]), _Id(c["rew"]), _Id(c["rew"]), _Id(c["gt"]),
             _Id(c["valid"])]
    loss = pb(batch, compare_ground_truth=gt_mode)
    loss.backward()
    return loss.detach(), q_s.grad.detach()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    a = ap.parse_args()
    out = {}
    for i, (name, A, over, gt_mode) in enumerate(CASES):
        c = build_case(name, A, over, gt_mode, seed=100 + i)
        loss, dq = run_reference(c, gt_mode)
        # the oracle restatement must agree before the vectors are written
        ocfg = qstep.StepConfig(GAMMA=c["cfg"]["GAMMA"], LOSS_CLIP=c["cfg"]["LOSS_CLIP"], LINEAR=c["cfg"]["LINEAR"],
                                REMOVE_BEFORE_REWARD=c["cfg"]["REMOVE_BEFORE_REWARD"], action_dim=A)
        qs = c["q_s"].clone().requires_grad_(True)
        if gt_mode:
            ol = qstep.td_loss_ground_truth(qs, c["act"], c["gt"], value_learning=c["cfg"]["VALUE_LEARNING"])
        else:
            ol, _ = qstep.td_loss(qs, c["q_no"], c["q_nt"], c["act"], c["rew"], c["rew"], c["valid"], ocfg)
        ol.backward()
        both_nan = bool(torch.isnan(loss)) and bool(torch.isnan(ol))
        assert both_nan or torch.equal(ol.detach(), loss), (name, ol, loss)
        assert torch.allclose(qs.grad, dq, rtol=0, atol=0, equal_nan=True), name
        for k in ("q_s", "q_no", "q_nt", "act", "rew", "valid", "gt"):
            out[f"{name}/{k}"] = c[k].numpy()
        out[f"{name}/loss"] = loss.numpy()
        out[f"{name}/dq"] = dq.numpy()
        out[f"{name}/cfg"] = np.array([c["cfg"]["GAMMA"], float(c["cfg"]["LOSS_CLIP"] == "rect"), float(c["cfg"]["LINEAR"]),
                                       float(c["cfg"]["REMOVE_BEFORE_REWARD"]), float(c["cfg"]["VALUE_LEARNING"]),
                                       float(gt_mode), float(A)])
        print(f"{name}: loss {loss.item():.9f}  (oracle identical)")
    path = os.path.join(a.out, "td_branches.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
