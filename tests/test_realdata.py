"""Real-data staging: our feather -> pinned uint8 batch path against what the
reference's own QLearningRealDataset returns for the committed mini data set
(tests/golden/realdata, made by oracle/make_loader_goldens.py)."""
import hashlib
import os

import numpy as np
import torch

from video_dqn_b200.realdata import PinnedFrameRing, QuadrupletLoader, QuadrupletTable

ROOT = os.path.join(os.path.dirname(__file__), "golden", "realdata")
MODES = {"inverse": dict(inverse_actions=True),
         "value": dict(one_action=True, value_learning=True),
         "previous": dict(inverse_actions=True, previous_images=True)}


def _sha(t):
    return np.frombuffer(hashlib.sha256(t.numpy().tobytes()).digest(), np.uint8)


def test_table_and_frames_match_reference_dataset():
    z = np.load(os.path.join(ROOT, "expected.npz"))
    for name, kw in MODES.items():
        tab = QuadrupletTable(os.path.join(ROOT, "data.feather"), **kw)
        assert len(tab) == 4
        ring = PinnedFrameRing(tab, batch_size=4, depth=1, workers=2, pin=False)
        before, after, act, rew, term, gt, valid = ring.fill(0, np.arange(4))
        F = 4 if name == "previous" else 1
        assert before.dtype == torch.uint8 and tuple(before.shape[-3:]) == (224, 224, 3)
        assert before.shape[0] == 4 and (before.dim() == 5) == (F == 4)
        for i in range(4):
            for tag, x in (("before", before), ("after", after)):
                fr = x[i] if F == 4 else x[i][None]
                assert np.array_equal(_sha(fr.contiguous()), z[f"{name}/{tag}{i}_sha"]), (name, tag, i)
                if name == "inverse":
                    assert np.array_equal(x[i].numpy(), z[f"{name}/{tag}{i}"])
            assert int(act[i]) == int(z[f"{name}/act{i}"])
            assert np.array_equal(rew[i].numpy(), z[f"{name}/rew{i}"])
            assert np.array_equal(term[i].numpy(), z[f"{name}/term{i}"])
            assert np.array_equal(valid[i].numpy(), z[f"{name}/valid{i}"])
            np.testing.assert_array_equal(gt[i].numpy(), z[f"{name}/gt{i}"])       # NaNs compare equal here
        assert rew.dtype == torch.int64 and gt.dtype == torch.float64


def test_loader_epochs_are_shuffled_permutations():
    tab = QuadrupletTable(os.path.join(ROOT, "data.feather"), inverse_actions=True)
    ld = QuadrupletLoader(tab, batch_size=2, seed=3, prefetch=2, workers=2, pin=False)
    acts = []
    for _ in range(4):                                   # two epochs of two batches
        b = next(ld)
        assert b[0].shape == (2, 224, 224, 3) and b[3].shape == (2, 5)
        acts.append(b[2].clone())
    for ep in (acts[:2], acts[2:]):
        assert sorted(torch.cat(ep).tolist()) == sorted(tab.action.tolist())
    ld2 = QuadrupletLoader(tab, batch_size=2, seed=3, prefetch=1, workers=1, pin=False)
    assert torch.equal(next(ld2)[2], acts[0])            # same seed, same order


def test_missing_action_source_raises_like_the_reference():
    import pytest
    with pytest.raises(Exception, match="not implemented"):
        QuadrupletTable(os.path.join(ROOT, "data.feather"))


def test_reward_ratio_matches_the_reference_dataset_class():
    """`dataset.reward_percentage()` (dataloaders/q_learning_real.py:51-53), value written by the
    reference's own class (oracle/make_loader_goldens.py)."""
    import json
    tab = QuadrupletTable(os.path.join(ROOT, "data.feather"), inverse_actions=True)
    ref = json.load(open(os.path.join(ROOT, "reward_percentage.json")))["reward_percentage"]
    assert tab.reward_percentage() == ref
