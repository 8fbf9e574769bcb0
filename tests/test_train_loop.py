"""Host logic of `video_dqn_b200.train.TrainLoop` (the bookkeeping of train_q_network.py:189-247) with a
stand-in learner: numbering after a resume, snapshot cadence and paths, the smoothed loss and its
logging cadence, lagged loss reads.  CPU only."""
import os
import types

import torch

from video_dqn_b200.train import TrainLoop


class FakeLearner:
    def __init__(self, lagged):
        self.calls, self.saved, self.resumed = [], [], None
        self.steps_done = 0
        self._losses = []
        if lagged:
            self.loss_value = self._loss_value

    def step(self, batch):
        self.steps_done += 1
        self._losses.append(float(batch))
        self.calls.append(batch)
        if not hasattr(self, "_buf"):
            self._buf = torch.zeros(1)
        self._buf.fill_(float(batch))            # ONE loss buffer, overwritten every step (as the learners do)
        return self._buf

    def _loss_value(self, k):
        assert k <= self.steps_done - 1 and k >= self.steps_done - 4      # still inside the four-slot ring
        return self._losses[k]

    def save_checkpoint(self, path):
        self.saved.append(path)

    def resume(self, path, resume_from):
        self.resumed = (path, resume_from)


class Writer:
    def __init__(self):
        self.rows = []

    def add_scalar(self, tag, value, step):
        self.rows.append((tag, value, step))


def _cfg(tmp, **kw):
    d = dict(folder=str(tmp), NUM_STEPS=250, CHECKPOINT_INTERVAL=100, writer=Writer())
    d.update(kw)
    return types.SimpleNamespace(**d)


def test_loop_numbering_snapshots_and_smoothed_loss(tmp_path):
    for lagged in (False, True):
        cfg = _cfg(tmp_path)
        lr = FakeLearner(lagged)
        losses = iter(float(i) for i in range(1, 10_000))
        out = TrainLoop(cfg, lr, losses).run()
        assert len(lr.calls) == 250 and lr.resumed is None
        # running = first loss, then 0.99 * running + 0.01 * loss (train_q_network.py:228-231)
        r = None
        for i in range(1, 251):
            r = float(i) if r is None else r * 0.99 + i * 0.01
        assert abs(out - r) < 1e-9
        assert [os.path.basename(p) for p in lr.saved] == ["sample100.torch", "sample200.torch"]   # :241-247
        assert os.path.isdir(tmp_path / "models")
        assert [row[2] for row in cfg.writer.rows] == [100, 200] and cfg.writer.rows[0][0] == "avg_q_loss/train"


def test_loop_resume_numbering(tmp_path):
    cfg = _cfg(tmp_path, NUM_STEPS=205)
    lr = FakeLearner(lagged=True)
    loop = TrainLoop(cfg, lr, iter(float(i) for i in range(100)))
    loop.run(resume_from=199)
    # sample_number = resume_from + 1, incremented BEFORE each step (:189,213): steps 201..205
    assert lr.resumed == (str(tmp_path / "models" / "sample199.torch"), 199)
    assert len(lr.calls) == 5 and loop.sample_number == 205 and lr.saved == []
    lr2 = FakeLearner(lagged=False)
    TrainLoop(cfg, lr2, iter(float(i) for i in range(100))).run(max_steps=3)
    assert len(lr2.calls) == 3


import pytest  # noqa: E402


@pytest.mark.gpu
def test_run_train_on_the_mini_data_set_and_resume(tmp_path):
    """`run_train` end to end on the committed mini data set (data.feather + JPEGs): six steps with
    snapshots every three, then a resumed run; snapshots load through `load_model_number`."""
    import math
    from video_dqn_b200.checkpoint import load_model_number
    from video_dqn_b200.train import run_train
    root = os.path.join(os.path.dirname(__file__), "golden", "realdata")
    cfg = types.SimpleNamespace(
        device="cuda", folder=str(tmp_path), writer=Writer(), DATASET=os.path.join(root, "data.feather"), SEED=0,
        ARCHITECTURE="extra_capacity", PANORAMA=False, PREVIOUS_IMAGES=False, ONE_ACTION=False, VALUE_LEARNING=False,
        USE_INVERSE_ACTIONS=True, CONFIDENCE_REWARD=False, TRAIN_ON_GROUND_TRUTH=False, LOSS_CLIP="rect", GAMMA=0.99,
        LINEAR=False, REMOVE_BEFORE_REWARD=False, LEARNING_RATE=1e-4, NUM_STEPS=6, CHECKPOINT_INTERVAL=3,
        TARGET_UPDATE_INTERVAL=4)
    out = run_train(cfg, batch_size=4, workers=2, log=lambda s: None)
    assert out is not None and math.isfinite(out) and out >= 0
    assert sorted(os.listdir(tmp_path / "models")) == ["sample3.torch", "sample6.torch"]
    snap = torch.load(tmp_path / "models" / "sample6.torch", map_location="cpu")
    assert snap["sample_number"] == 6 and len(snap["model_state_dict"]) == 250
    assert int(snap["optimizer_state_dict"]["state"][0]["step"]) == 6
    cfg.NUM_STEPS = 9
    out2 = run_train(cfg, resume_from=6, batch_size=4, workers=2, log=lambda s: None)   # steps 8 and 9 (:189,213)
    assert math.isfinite(out2)
    assert "sample9.torch" in os.listdir(tmp_path / "models")
    snap9 = torch.load(tmp_path / "models" / "sample9.torch", map_location="cpu")
    assert int(snap9["optimizer_state_dict"]["state"][0]["step"]) == 8                  # 6 + the two resumed steps
    m = load_model_number(cfg, 9)
    assert torch.equal(m.state_dict()["top.4.bias"].cpu(), snap9["model_state_dict"]["top.4.bias"])


@pytest.mark.gpu
def test_in_step_target_sync_matches_the_reference_schedule():
    """train_q_network.py:213-216: at the top of iteration n (n % TARGET_UPDATE_INTERVAL == 0) the target
    network becomes a copy of the model.  The fused learner folds that copy into the Adam pass of
    iteration n - 1 (captured as a second CUDA graph): after iteration n - 1 the target equals the model
    exactly, during iteration n the model moves on and the target stays."""
    from oracle import qstep
    from video_dqn_b200.learner import QLearner, StepConfig
    from video_dqn_b200.qnet import HabitatDQNMultiAction
    dev = torch.device("cuda:0")
    sd = qstep.init_state(seed=4, randomize_bn=True)
    nets = []
    for _ in range(2):
        m = HabitatDQNMultiAction(3, 5, extra_capacity=True, panorama=False)
        m.load_state_dict(sd, strict=True)
        nets.append(m.to(dev))
    model, target = nets
    lr = QLearner(model, target, StepConfig(TARGET_UPDATE_INTERVAL=3), batch_size=8)
    names = model._grad_names
    mp, tp = dict(model.named_parameters()), dict(target.named_parameters())
    init = {n: mp[n].detach().clone() for n in names}
    snap = {}
    for it in range(1, 7):
        lr.step([t.to(dev) for t in qstep.synthetic_batch(8, seed=it)])
        torch.cuda.synchronize()
        same = all(torch.equal(mp[n], tp[n]) for n in names)
        if it in (2, 5):                       # iteration n - 1 for n = 3, 6: the copy has just happened
            assert same, it
            snap = {n: mp[n].detach().clone() for n in names}
        elif it == 1:                          # nothing synced yet: the target still holds the initial weights
            assert all(torch.equal(tp[n], init[n]) for n in names) and not same
        else:                                  # the model has moved on, the target keeps the synced copy
            assert not same and all(torch.equal(tp[n], snap[n]) for n in names), it
    # the synced target's forward uses the new weights (its bf16 operands were re-derived)
    x = qstep.synthetic_batch(8, seed=9)[0].to(dev)
    lr2_target_q = target.eval()(x)
    for n in names:
        mp[n].data.copy_(tp[n].data)
    from video_dqn_b200.optim import bump_arena_epoch
    bump_arena_epoch(lr.opt.param_arena)
    model.eval()
    # (the fp32 Q-head GEMMs split K across blocks and combine with atomics: equal up to summation order)
    assert torch.allclose(model(x), lr2_target_q, rtol=0, atol=1e-5)


def test_run_train_has_no_cpu_fallback(tmp_path):
    """Everything up to the device boundary runs (table, loader, models, optimizer); the learner then
    refuses a CPU device instead of falling back."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from video_dqn_b200.train import run_train
    root = os.path.join(os.path.dirname(__file__), "golden", "realdata")
    cfg = types.SimpleNamespace(
        device="cpu", folder=str(tmp_path), writer=Writer(), DATASET=os.path.join(root, "data.feather"), SEED=0,
        ARCHITECTURE="extra_capacity", PANORAMA=False, PREVIOUS_IMAGES=False, ONE_ACTION=False, VALUE_LEARNING=False,
        USE_INVERSE_ACTIONS=True, CONFIDENCE_REWARD=False, TRAIN_ON_GROUND_TRUTH=False, LOSS_CLIP="rect", GAMMA=0.99,
        LINEAR=False, REMOVE_BEFORE_REWARD=False, LEARNING_RATE=1e-4, NUM_STEPS=2, CHECKPOINT_INTERVAL=10,
        TARGET_UPDATE_INTERVAL=4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        run_train(cfg, batch_size=4, workers=2, log=lambda s: None)
