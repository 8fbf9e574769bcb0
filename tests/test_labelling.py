"""Host logic of the pseudo-labelling loop (dataset/process_episodes_real.py:165-181) over a stand-in
labeller: batching in `runner.B`, tail padding and trimming, row order, decode identical to the
Q-learning loader's frames, column written in the reference's [N, 1] int64 shape.  CPU only."""
import os
import shutil

import numpy as np
import pandas as pd
import torch

from video_dqn_b200.inverse import label_frame_pairs, label_table

ROOT = os.path.join(os.path.dirname(__file__), "golden", "realdata")


class FakeLabeller:
    """labels a pair by a checksum of its decoded frames, so order / padding mistakes are visible"""
    dev = torch.device("cpu")

    def __init__(self, B):
        self.B, self.calls = B, 0

    def label(self, k, k1):
        assert k.shape == (self.B, 224, 224, 3) and k.dtype == torch.uint8 and k1.shape == k.shape
        self.calls += 1
        return (k.reshape(self.B, -1).long().sum(1) + 2 * k1.reshape(self.B, -1).long().sum(1)) % 3


def test_label_frame_pairs_batches_pads_and_keeps_order():
    t = pd.read_feather(os.path.join(ROOT, "data.feather"))
    before, after = list(t["before_image"]), list(t["after_image"])
    ref = label_frame_pairs(before, after, FakeLabeller(1), workers=2, root=ROOT)        # one pair per call
    assert ref.shape == (len(t), 1) and ref.dtype == torch.int64
    for B in (2, 3, 8):                                   # 3 and 8 leave a padded tail
        r = FakeLabeller(B)
        out = label_frame_pairs(before, after, r, workers=2, root=ROOT)
        assert torch.equal(out, ref) and r.calls == -(-len(t) // B)
    # the decoded frames are the loader's frames (expected.npz holds the reference dataset's `before` frames)
    z = np.load(os.path.join(ROOT, "expected.npz"))

    class Capture(FakeLabeller):
        def label(self, k, k1):
            self.k = k.clone()
            return super().label(k, k1)
    c = Capture(len(t))
    label_frame_pairs(before, after, c, workers=2, root=ROOT)
    assert np.array_equal(c.k[0].numpy(), z["inverse/before0"])
    assert label_frame_pairs([], [], FakeLabeller(4)).shape == (0, 1)


def test_label_table_writes_the_column(tmp_path):
    shutil.copytree(ROOT, tmp_path / "d")
    path = str(tmp_path / "d" / "data.feather")
    t0 = pd.read_feather(path).drop(columns=["inverse_actions"])
    t0.to_feather(path)
    acts = label_table(path, FakeLabeller(3), workers=2)
    t1 = pd.read_feather(path)
    assert "inverse_actions" in t1.columns and np.array_equal(np.asarray(t1["inverse_actions"]).reshape(-1, 1), acts.numpy())
    assert list(t1["before_image"]) == list(t0["before_image"])


import pytest  # noqa: E402

REF = os.environ.get("VDQN_REFERENCE", "/root/reference")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "dataloaders")), reason="needs the reference checkout")
def test_pairs_match_the_reference_image_stream():
    """The uint8 pairs `label_frame_pairs` feeds the labeller, normalised as the first kernel does, equal
    what the reference's own `ImageStream` yields for the same rows (dataloaders/image_streams.py:15-27;
    dataset/process_episodes_real.py:168-173)."""
    import sys
    from oracle import qstep
    t = pd.read_feather(os.path.join(ROOT, "data.feather"))
    before, after = list(t["before_image"]), list(t["after_image"])
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    try:
        from dataloaders.image_streams import ImageStream
        os.chdir(ROOT)                                   # the table's paths are relative to its directory
        ims = np.stack((t["before_image"], t["after_image"]), axis=1)         # :168-169
        ref = [ImageStream(ims)[i] for i in range(len(t))]
    finally:
        os.chdir(cwd)
        sys.path.remove(REF)

    class Capture(FakeLabeller):
        def label(self, k, k1):
            self.k, self.k1 = k.clone(), k1.clone()
            return super().label(k, k1)
    c = Capture(len(t))
    label_frame_pairs(before, after, c, workers=2, root=ROOT)
    for i, (be, ae) in enumerate(ref):
        assert (qstep.to_imgnet(c.k[i:i + 1])[0] - be).abs().max().item() <= 1e-6
        assert (qstep.to_imgnet(c.k1[i:i + 1])[0] - ae).abs().max().item() <= 1e-6
