"""Fixture for the real-data staging path: a tiny `data.feather` + JPEG frames and
what the REFERENCE's own `QLearningRealDataset` (dataloaders/q_learning_real.py) returns for them.

TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference).  The reference class
is imported unmodified (shims: `np.int`, which NumPy removed, and the `util` package path).  Its
frames are normalised fp32 tensors; the fixture stores the uint8 pixels they were made from (checked
here: `to_imgnet(uint8)` reproduces the reference tensor) plus every label, for three loader modes.

usage:  python -m oracle.make_loader_goldens [--out tests/golden/realdata]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys

import numpy as np
import torch

from . import qstep

REF = os.environ.get("VDQN_REFERENCE", "/root/reference")


def make_frames(root, n=7):
    from PIL import Image
    os.makedirs(os.path.join(root, "ep0"), exist_ok=True)
    rng = np.random.default_rng(5)
    sizes = [(320, 240), (256, 300), (224, 224), (400, 260), (230, 310), (300, 300), (260, 228)]
    for i in range(n):
        w, h = sizes[i % len(sizes)]
        yy, xx = np.mgrid[0:h, 0:w]
        img = np.stack([127 + 100 * np.sin(xx / (9.0 + i) + c) * np.cos(yy / (13.0 + c)) for c in range(3)], -1)
        img += rng.normal(0, 6, img.shape)
        Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(os.path.join(root, "ep0", "%04d.jpg" % i),
                                                                    quality=80)


def make_table(root):
    import pandas as pd
    rows = 4
    t = pd.DataFrame({"before_image": ["ep0/%04d.jpg" % i for i in range(rows)],
                      "after_image": ["ep0/%04d.jpg" % (i + 3) for i in range(rows)],
                      "ep_id": [0] * rows, "im_start": [1] * rows, "im_stop": [7] * rows})
    rng = np.random.default_rng(6)
    det = rng.uniform(0.5, 1.0, (rows, 5))
    det[0, 0], det[1, 3], det[2, 4] = 0.99, 0.74, 0.70
    steps = rng.integers(0, 25, (rows, 5)).astype(np.float64)
    steps[0, 1] = steps[2, 2] = steps[3, 0] = np.inf
    for c in range(5):
        t[f"detector_score{c}"] = det[:, c]
    for c in range(5):
        t[f"sparse_reward{c}"] = (det[:, c] > 0.9).astype(int)
    for c in range(5):
        t[f"steps_to_reward{c}"] = steps[:, c]
    for c in range(5):
        t[f"steps_to_reward_neg{c}"] = steps[:, c]
    t["inverse_actions"] = np.array([2, 0, 1, 2])
    t.to_feather(os.path.join(root, "data.feather"))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "realdata"))
    a = ap.parse_args()
    root = os.path.abspath(a.out)
    os.makedirs(root, exist_ok=True)
    make_frames(root)
    rows = make_table(root)

    np.int = int                                  # removed from NumPy; q_learning_real.py:82 uses it
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(root)                                # the table stores paths relative to its directory
    try:
        from dataloaders.q_learning_real import QLearningRealDataset
        out = {}
        modes = {"inverse": dict(inverse_actions=True),
                 "value": dict(one_action=True, value_learning=True),
                 "previous": dict(inverse_actions=True, previous_images=True)}
        for name, kw in modes.items():
            ds = QLearningRealDataset(location="data.feather", **kw)
            assert len(ds) == rows
            for i in range(rows):
                bi, ai, act, rew, term, gt, valid = ds[i]
                for tag, x in (("before", bi), ("after", ai)):
                    x4 = x if x.dim() == 4 else x[None]                       # [F,3,224,224]
                    u8 = torch.round((x4 * torch.tensor(qstep.IMAGENET_STD).view(1, 3, 1, 1)
                                      + torch.tensor(qstep.IMAGENET_MEAN).view(1, 3, 1, 1)) * 255).to(torch.uint8)
                    u8 = u8.permute(0, 2, 3, 1).contiguous()                 # [F,224,224,3]
                    back = qstep.to_imgnet(u8)
                    assert (back - x4).abs().max().item() <= 1e-6, (name, i, tag)
                    if name == "inverse":
                        out[f"{name}/{tag}{i}"] = u8[0].numpy()
                    out[f"{name}/{tag}{i}_sha"] = np.frombuffer(hashlib.sha256(u8.numpy().tobytes()).digest(), np.uint8)
                out[f"{name}/act{i}"] = np.asarray(act)
                out[f"{name}/rew{i}"] = np.asarray(rew)
                out[f"{name}/term{i}"] = np.asarray(term)
                out[f"{name}/gt{i}"] = np.asarray(gt, dtype=np.float64)
                out[f"{name}/valid{i}"] = np.asarray(valid)
        # `Reward Ratio` the trainer prints (train_q_network.py:110), from the reference's own class
        with open("reward_percentage.json", "w") as f:
            json.dump({"reward_percentage": float(ds.reward_percentage())}, f)
    finally:
        os.chdir(cwd)
        sys.path.remove(REF)
    np.savez_compressed(os.path.join(root, "expected.npz"), **out)
    print("wrote", root, sorted(os.listdir(root)), sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(root) for f in fs), "bytes")


if __name__ == "__main__":
    main()
