"""Snapshot format and resume sequence of train_q_network.py:36-57,190-208,241-247:
`sample{n}.torch` = {sample_number, model_state_dict (250 keys), optimizer_state_dict (torch.optim.Adam
layout)}.  CPU part: file format and naming; the cross-loading test against the reference's own module
runs only where /root/reference exists (this container).  GPU part: save -> resume -> identical next
step."""
import os
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VDQN_REFERENCE", "/root/reference")


def _config(tmp, **kw):
    d = dict(device="cpu", folder=str(tmp), ARCHITECTURE="extra_capacity", PANORAMA=False, PREVIOUS_IMAGES=False,
             VALUE_LEARNING=False, ONE_ACTION=False)
    d.update(kw)
    return types.SimpleNamespace(**d)


def test_build_model_follows_the_config_switches(tmp_path):
    from video_dqn_b200.checkpoint import build_model
    m = build_model(_config(tmp_path))
    assert (m.action_dim, m.num_frames, len(m.state_dict())) == (3, 1, 250)
    assert build_model(_config(tmp_path, ONE_ACTION=True)).action_dim == 1
    assert build_model(_config(tmp_path, VALUE_LEARNING=True)).action_dim == 1
    assert build_model(_config(tmp_path, PREVIOUS_IMAGES=True)).num_frames == 4
    b = build_model(_config(tmp_path, ARCHITECTURE="basic", PANORAMA=True))
    assert isinstance(b.top, torch.nn.Linear) and b.top.in_features == 512 * 4


def test_load_model_number_reads_the_reference_snapshot_layout(tmp_path):
    from video_dqn_b200.checkpoint import build_model, load_model_number, snapshot_path
    cfg = _config(tmp_path)
    src = build_model(cfg)
    opt = torch.optim.Adam(src.parameters(), lr=1e-4)                   # the optimizer the reference saves
    os.makedirs(tmp_path / "models")
    assert snapshot_path(cfg, 7) == str(tmp_path / "models" / "sample7.torch")
    torch.save({"sample_number": 7, "model_state_dict": src.state_dict(), "optimizer_state_dict": opt.state_dict()},
               snapshot_path(cfg, 7))
    m = load_model_number(cfg, 7)
    for k, v in src.state_dict().items():
        assert torch.equal(m.state_dict()[k], v), k
    m2 = load_model_number(cfg, None, model_loc=snapshot_path(cfg, 7))   # the `model_loc` override (:52-53)
    assert torch.equal(m2.state_dict()["top.4.bias"], src.state_dict()["top.4.bias"])
    with pytest.raises(FileNotFoundError):
        load_model_number(cfg, 8)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "archs")), reason="needs the reference checkout")
def test_snapshots_cross_load_with_the_reference_module(tmp_path):
    """A snapshot written the reference's way from the reference's own module and torch.optim.Adam
    loads into the drop-in module and FusedAdam; a drop-in snapshot loads back into the reference
    module and torch.optim.Adam (strict keys, equal tensors, same optimizer-state indices)."""
    from oracle.make_goldens import import_reference_model
    from video_dqn_b200.checkpoint import build_model
    from video_dqn_b200.optim import FusedAdam
    RefNet, restore = import_reference_model()
    try:
        ref = RefNet(3, 5, extra_capacity=True, panorama=False)
    finally:
        restore()
    ropt = torch.optim.Adam(ref.parameters(), lr=1e-4)
    g = torch.Generator().manual_seed(0)
    for n, p in ref.named_parameters():
        if not n.startswith("resnet.fc"):                               # resnet.fc never gets a gradient
            p.grad = torch.randn(p.shape, generator=g) * 1e-2
    ropt.step()
    path = str(tmp_path / "sample3.torch")
    torch.save({"sample_number": 3, "model_state_dict": ref.state_dict(), "optimizer_state_dict": ropt.state_dict()},
               path)
    snap = torch.load(path, map_location="cpu")
    mine = build_model(_config(tmp_path))
    res = mine.load_state_dict(snap["model_state_dict"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    fopt = FusedAdam(mine.parameters(), lr=1e-4)
    fopt.load_state_dict(snap["optimizer_state_dict"])
    mine_params = list(mine.parameters())
    rs = ropt.state_dict()["state"]
    assert sorted(rs) == [i for i in range(70) if i not in (60, 61)]    # fc.weight / fc.bias have no state
    for i, st in rs.items():
        assert torch.equal(fopt.state[mine_params[i]]["exp_avg_sq"], st["exp_avg_sq"])
    # and back: what the drop-in side writes, the reference side reads
    out = {"sample_number": 4, "model_state_dict": mine.state_dict(), "optimizer_state_dict": fopt.state_dict()}
    torch.save(out, path)
    snap2 = torch.load(path, map_location="cpu")
    res = ref.load_state_dict(snap2["model_state_dict"], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    ropt2 = torch.optim.Adam(ref.parameters(), lr=1e-4)
    ropt2.load_state_dict(snap2["optimizer_state_dict"])
    assert torch.equal(ropt2.state_dict()["state"][69]["exp_avg"], rs[69]["exp_avg"])


@pytest.mark.gpu
def test_learner_resume_reproduces_the_uninterrupted_run(tmp_path):
    from oracle import qstep
    from video_dqn_b200.learner import QLearner, StepConfig
    from video_dqn_b200.qnet import HabitatDQNMultiAction
    dev = torch.device("cuda:0")
    sd = qstep.init_state(seed=4, randomize_bn=True)

    def fresh(state):
        nets = []
        for _ in range(2):
            m = HabitatDQNMultiAction(3, 5, extra_capacity=True, panorama=False)
            m.load_state_dict(state, strict=True)
            nets.append(m.to(dev))
        return QLearner(nets[0], nets[1], StepConfig(), batch_size=8)

    batches = [[t.to(dev) for t in qstep.synthetic_batch(8, seed=1 + i)] for i in range(4)]
    a = fresh(sd)
    for b in batches[:2]:
        a.step(b)
    path = str(tmp_path / "sample2.torch")
    a.save_checkpoint(path)
    snap = torch.load(path, map_location="cpu")
    assert set(snap) == {"sample_number", "model_state_dict", "optimizer_state_dict"}
    assert snap["sample_number"] == 2 and len(snap["model_state_dict"]) == 250
    # the optimizer state is torch.optim.Adam's: it loads into one built over the same 70 parameters
    cpu = HabitatDQNMultiAction(3, 5, extra_capacity=True, panorama=False)
    topt = torch.optim.Adam(cpu.parameters(), lr=1e-4)
    topt.load_state_dict(snap["optimizer_state_dict"])
    assert len(topt.state_dict()["state"]) == 68 and int(topt.state_dict()["state"][0]["step"]) == 2
    # resume in a fresh learner built from DIFFERENT initial weights, run two captured-graph steps first
    # so the graphs exist before the state is replaced
    other = fresh(qstep.init_state(seed=5, randomize_bn=True))
    other.step(batches[0]); other.step(batches[1])
    other.resume(path)
    assert other.sample_number == 3                       # sample_number = resume_from + 1 (:190)
    assert int(other.step_dev.item()) == 2
    # the reference's resume also re-syncs the target network (:208); give the uninterrupted run the
    # same sync so the two trajectories are comparable
    a.sync_target_now()
    la = [a.step(b).item() for b in batches[2:]]
    lo = [other.step(b).item() for b in batches[2:]]
    torch.cuda.synchronize()
    # first step after the resume: same weights, same moments, same batch -> same loss up to the summation
    # order of atomically accumulated sums.  From the second step on Adam's sign-normalised update turns
    # last-bit gradient differences of near-zero gradients into +-lr parameter differences, which moves
    # the loss by a fraction of a percent (cf. test_fused_learner_three_steps_match_oracle)
    assert abs(la[0] - lo[0]) <= 1e-5 * abs(la[0]), (la, lo)
    assert abs(la[1] - lo[1]) <= 2e-2 * abs(la[1]), (la, lo)
    pa, po = dict(a.model.named_parameters()), dict(other.model.named_parameters())
    for n in pa:
        d = (pa[n].detach() - po[n].detach()).abs()
        # Adam normalises by sqrt(v): where a gradient is ~0, last-bit differences from the atomically
        # accumulated sums can flip an update's sign -- bounded by 2 lr per step, and rare
        assert d.max().item() <= 8e-4 and d.mean().item() <= 5e-5, (n, d.max().item(), d.mean().item())
    ta = dict(a.target_net.named_parameters())
    to = dict(other.target_net.named_parameters())
    assert (ta["top.4.weight"] - to["top.4.weight"]).abs().max().item() <= 1e-7   # target <- model at resume (:208)
    assert other.opt.state_dict()["state"][0]["step"].item() == 4


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "archs")), reason="needs the reference checkout")
def test_inverse_module_state_loads_strictly_into_both_reference_classes(tmp_path):
    """`InverseActionModule().state_dict()` (what `InverseModelTrainer.state_dict()` returns and
    `model-N.pth` / `inverse_model.torch` hold) loads with strict=True into the trainer's class
    (train_inverse_model.model) and into the labeller's class (archs/inverse_action2.model), and theirs
    load into it."""
    import sys
    from oracle.make_inverse_train_goldens import import_reference_trainer
    from video_dqn_b200.inverse import InverseActionModule
    mine = InverseActionModule()
    sd = mine.state_dict()
    _T, trainer_model = import_reference_trainer()
    res = trainer_model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    # the arch file defines the same absl flag as the trainer ('bottleneck_size'), so the two reference
    # modules cannot be imported into one process: the labeller's class is checked in a child process
    import subprocess
    path = str(tmp_path / "inverse_model.torch")
    torch.save(sd, path)
    child = (
        "import sys, torch, torchvision.models as tvm\n"
        "orig = tvm.resnet18\n"
        "tvm.resnet18 = lambda pretrained=False, **kw: orig(weights=None, **kw)\n"
        f"sys.path.insert(0, {REF!r})\n"
        "from archs import inverse_action2\n"
        "m = inverse_action2.model()\n"
        f"res = m.load_state_dict(torch.load({path!r}, map_location='cpu'), strict=True)\n"   # process_episodes_real.py:92-93
        "assert not res.missing_keys and not res.unexpected_keys\n"
        "print('arch-ok')\n")
    out = subprocess.run([sys.executable, "-c", child], capture_output=True, text=True)
    assert "arch-ok" in out.stdout, out.stderr[-2000:]
    res = mine.load_state_dict(trainer_model.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert [n for n, p in mine.named_parameters() if p.requires_grad] == \
        [n for n, p in trainer_model.named_parameters() if p.requires_grad]
