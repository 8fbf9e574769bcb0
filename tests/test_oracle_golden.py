"""The oracle against the committed golden vectors (generated from the reference
itself by oracle/make_goldens.py) and the hand-checkable TD known-answer test
of SURVEY.md 8c.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import qstep

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _sample(t, n=64):
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].double().numpy()


@pytest.mark.parametrize("rbn", [1, 0])
def test_oracle_reproduces_reference_goldens(rbn):
    torch.set_num_threads(os.cpu_count() or 1)
    g = np.load(os.path.join(GOLD, f"step_b8_bn{rbn}.npz"))
    B = int(g["meta/B"])
    tr = qstep.OracleTrainer(qstep.init_state(seed=4, randomize_bn=bool(rbn)))
    for it in range(int(g["meta/steps"]) if rbn else 1):
        loss, grads, aux = tr.step(qstep.synthetic_batch(B, seed=1 + it))
        p = f"step{it}/"
        assert abs(loss.item() - float(g[p + "loss"])) <= 2e-6 * abs(float(g[p + "loss"]))
        np.testing.assert_allclose(aux["q_s"].numpy(), g[p + "q_s"], atol=5e-6)
        np.testing.assert_allclose(aux["q_next_target"].numpy(), g[p + "q_next_target"], atol=5e-6)
        assert (aux["best"].numpy() == g[p + "best"]).all()      # integer: bit-exact
        np.testing.assert_allclose(aux["y"].numpy(), g[p + "y"], atol=5e-6)
        for n in tr.names:
            ref_l2 = float(g[p + f"grad/{n}/l2"])
            assert abs(grads[n].double().norm().item() - ref_l2) <= 1e-4 * ref_l2 + 1e-12, n
            np.testing.assert_allclose(_sample(grads[n]), g[p + f"grad/{n}/sample"],
                                       rtol=1e-3, atol=1e-5 * ref_l2 + 1e-12, err_msg=n)
            np.testing.assert_allclose(_sample(tr.sd[n]), g[p + f"param/{n}/sample"],
                                       rtol=1e-5, atol=1e-7, err_msg=n)
        if it == 0:
            np.testing.assert_allclose(grads["top.4.weight"].numpy(),
                                       g[p + "gradfull/top.4.weight"], rtol=1e-4, atol=1e-8)


def test_td_known_answer():
    cfg = qstep.StepConfig()
    act = torch.tensor([2, 0])
    rew = torch.tensor([[0, 1, 0, 0, 0], [0, 0, 0, 0, 1]])
    qs = torch.tensor([[[.1, .2, .3], [.5, .4, .3], [0, 0, 0], [1.5, -1, .2], [.9, .8, .7]],
                       [[-.2, .1, 0], [.3, .3, .1], [.6, .2, .9], [.05, .15, .25], [.4, .4, .4]]],
                      requires_grad=True)
    qo = torch.tensor([[[.1, .9, .2], [.7, .7, .1], [0, .1, .2], [.3, .2, .1], [.5, .6, .4]],
                       [[.2, .1, .3], [.9, .1, .1], [.1, .8, .8], [0, 0, 0], [-1., -2, -3]]])
    qt = torch.tensor([[[.4, .5, .6], [.2, .9, .3], [1.5, 1.6, 1.7], [-.5, .1, .2], [.3, .2, .1]],
                       [[.7, .8, .9], [.25, .5, .75], [.1, .3, .2], [.6, .1, .1], [.9, .1, .1]]])
    loss, aux = qstep.td_loss(qs, qo, qt, act, rew, rew, torch.ones_like(rew), cfg)
    assert aux["best"].tolist() == [[1, 0, 2, 0, 1], [2, 0, 1, 0, 0]]
    np.testing.assert_allclose(aux["y"].numpy(),
                               [[.495, 1, 1, 0, .198], [.891, .2475, .297, .594, 1]], atol=1e-6)
    assert abs(loss.item() - 0.188040555) < 1e-7
    loss.backward()
    exp = torch.zeros(2, 5, 3)
    exp[0, :, 2] = torch.tensor([-.0195, -.07, -.1, .02, .0502])
    exp[1, :, 0] = torch.tensor([-.1091, .00525, .0303, -.0544, -.06])
    np.testing.assert_allclose(qs.grad.numpy(), exp.numpy(), atol=1e-6)
    np.testing.assert_allclose(
        qstep.td_grad_closed_form(qs.detach(), act, aux["y"], torch.ones_like(rew), cfg).numpy(),
        exp.numpy(), atol=1e-6)


def test_batched_2b_forward_is_identical():
    """eval-mode BN => one 2B forward over cat(s, s') equals two B forwards."""
    sd = qstep.init_state(seed=4, randomize_bn=True)
    b = qstep.synthetic_batch(2, seed=3)
    with torch.no_grad():
        both = qstep.q_forward(sd, torch.cat([b[0], b[1]]))
        sep = torch.cat([qstep.q_forward(sd, b[0]), qstep.q_forward(sd, b[1])])
    assert torch.allclose(both, sep, atol=1e-6)


def test_bad_shape_raises():
    sd = qstep.init_state(seed=4)
    with pytest.raises(Exception, match="bad shape"):
        qstep.q_forward(sd, torch.zeros(1, 4, 3, 224, 224))


def test_u8_normalisation_fma_is_exact():
    """stem_pack_u8 evaluates fma(b, 1/(255 std), -mean/std) instead of ((b/255) - mean)/std
    (util/torch.py:26-36).  Exhaustive over the 3 x 256 possible inputs: both round to the same
    bf16, so the packed uint8 path is bit-identical to the reference normalisation."""
    frame = torch.arange(256, dtype=torch.uint8).view(1, 256, 1, 1).repeat(1, 1, 1, 3)
    ref = qstep.to_imgnet(frame)[0, :, :, 0].to(torch.bfloat16)              # [3, 256]
    for c in range(3):
        std, mean = np.float64(np.float32(qstep.IMAGENET_STD[c])), np.float64(np.float32(qstep.IMAGENET_MEAN[c]))
        ka = np.float64(np.float32(1.0 / (255.0 * std)))
        kb = np.float64(np.float32(-mean / std))
        # b*ka + kb is exact in float64 (8 + 24 significant bits), so one rounding = fmaf
        fma = (np.arange(256, dtype=np.float64) * ka + kb).astype(np.float32)
        got = torch.from_numpy(fma).to(torch.bfloat16)
        assert torch.equal(got, ref[c]), c


def _td_cases():
    z = np.load(os.path.join(GOLD, "td_branches.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    for nm in names:
        c = {k.split("/")[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith(nm + "/")}
        gamma, rect, linear, masked, vl, gt_mode, A = c["cfg"].tolist()
        yield nm, c, qstep.StepConfig(GAMMA=gamma, LOSS_CLIP="rect" if rect else "none", LINEAR=bool(linear),
                                      REMOVE_BEFORE_REWARD=bool(masked), action_dim=int(A)), bool(gt_mode), bool(vl)


def test_td_branches_match_reference_goldens():
    """Every branch of process_batch's loss (train_q_network.py:139-180: Double DQN, LINEAR, clip on/off,
    REMOVE_BEFORE_REWARD, and the compare_ground_truth branches with / without VALUE_LEARNING) against
    vectors produced by the reference's own closure (oracle/make_td_goldens.py): bit-identical."""
    n = 0
    for nm, c, cfg, gt_mode, vl in _td_cases():
        qs = c["q_s"].clone().requires_grad_(True)
        if gt_mode:
            loss = qstep.td_loss_ground_truth(qs, c["act"], c["gt"], value_learning=vl)
        else:
            loss, _ = qstep.td_loss(qs, c["q_no"], c["q_nt"], c["act"], c["rew"], c["rew"], c["valid"], cfg)
        loss.backward()
        assert torch.equal(loss.detach(), c["loss"]), nm
        assert torch.equal(qs.grad, c["dq"]), nm
        n += 1
    assert n == 7


def test_inverse_model_oracle_reproduces_reference_golden():
    """oracle/inverse.py against the outputs of the reference's own inverse_action2.model
    (tests/golden/inverse_b4.npz, oracle/make_inverse_goldens.py)."""
    from oracle import inverse as oinv
    z = np.load(os.path.join(GOLD, "inverse_b4.npz"))
    sd = oinv.init_state(seed=int(z["seed"]))
    g = torch.Generator().manual_seed(int(z["data_seed"]))
    k = torch.randn(4, 3, 224, 224, generator=g)
    k1 = torch.randn(4, 3, 224, 224, generator=g)
    with torch.no_grad():
        enc, y = oinv.forward(sd, k, k1)
    np.testing.assert_allclose(y.numpy(), z["y"], atol=1e-5)
    np.testing.assert_allclose(enc.numpy(), z["encoding"], atol=1e-6)
    assert torch.equal(oinv.label(sd, k, k1), torch.from_numpy(z["y"]).argmax(1))


def test_basic_architecture_oracle_reproduces_reference_golden():
    """extra_capacity=False (trunk + global average pool + one Linear), eval mode, F = 1 and F = 4:
    oracle against the outputs of the reference module (oracle/make_basic_goldens.py)."""
    z = np.load(os.path.join(GOLD, "basic_b4.npz"))
    for tag, F in (("f1", 1), ("f4", 4)):
        sd = qstep.init_state_basic(seed=4, num_frames=F)
        g = torch.Generator().manual_seed(int(z[f"{tag}/data_seed"]))
        B = 4 if F == 1 else 2
        x = torch.randn(B, F, 3, 224, 224, generator=g) if F > 1 else torch.randn(B, 3, 224, 224, generator=g)
        with torch.no_grad():
            q = qstep.q_forward_basic(sd, x)
        np.testing.assert_allclose(q.numpy(), z[f"{tag}/q"], atol=2e-6)


def inverse_train_batches(z):
    """the seeded batches oracle/make_inverse_train_goldens.py stepped the reference on"""
    g = torch.Generator().manual_seed(int(z["data_seed"]))
    B = int(z["batch"])
    for s in range(int(z["steps"])):
        k = torch.randn(B, 3, 224, 224, generator=g)
        k1 = torch.randn(B, 3, 224, 224, generator=g)
        act = torch.randint(0, 3, (B,), generator=g)
        yield s, k, k1, act, torch.from_numpy(z[f"keep{s}"])


def test_inverse_model_training_oracle_reproduces_reference_golden():
    """oracle/inverse.py's training step (trainer forward with ReLU after fc2 and element dropout,
    cross-entropy, Adam) against three steps of the reference's own train_inverse_model.model +
    nn.CrossEntropyLoss + torch.optim.Adam (tests/golden/inverse_train_b4.npz)."""
    from oracle import inverse as oinv
    torch.set_num_threads(os.cpu_count() or 1)
    z = np.load(os.path.join(GOLD, "inverse_train_b4.npz"))
    tr = oinv.InverseOracleTrainer(oinv.init_state(seed=int(z["seed"])), lr=float(z["lr"]))
    for s, k, k1, act, keep in inverse_train_batches(z):
        loss, grads, y, correct = tr.step(k, k1, act, keep)
        assert abs(loss.item() - float(z[f"loss{s}"])) <= 2e-6
        np.testing.assert_allclose(y.numpy(), z[f"y{s}"], atol=2e-6)
        assert correct == int(z[f"correct{s}"])                   # integer: exact
        for n in oinv.TRAINABLE:
            ref_l2 = float(z[f"s{s}/grad/{n}/l2"])
            assert abs(grads[n].double().norm().item() - ref_l2) <= 1e-4 * ref_l2 + 1e-12, n
            np.testing.assert_allclose(_sample(grads[n]), z[f"s{s}/grad/{n}/sample"],
                                       rtol=1e-3, atol=1e-5 * ref_l2 + 1e-12, err_msg=n)
            np.testing.assert_allclose(_sample(tr.sd[n]), z[f"s{s}/param/{n}/sample"],
                                       rtol=1e-5, atol=2e-6, err_msg=n)


@pytest.mark.parametrize("fixture", ["basic_train_b8.npz", "basic_train_f4_b4.npz"])
def test_basic_architecture_training_oracle_reproduces_reference_golden(fixture):
    """oracle/qstep.py:BasicOracleTrainer (train-mode BatchNorm: batch statistics, two running-statistics
    updates per frame and step) against two steps of the reference's own module built with
    extra_capacity=False, its own process_batch and torch.optim.Adam: single frame at B = 8 and the
    four-frame panorama layout at B = 4 (every frame through the trunk separately)."""
    from oracle.make_basic_train_goldens import frames_batch
    torch.set_num_threads(os.cpu_count() or 1)
    g = np.load(os.path.join(GOLD, fixture))
    B, nf = int(g["meta/B"]), int(g["meta/frames"])
    tr = qstep.BasicOracleTrainer(qstep.init_state_basic(seed=4, num_frames=nf))
    for it in range(int(g["meta/steps"])):
        loss, grads, aux = tr.step(frames_batch(B, nf, 1 + it))
        p = f"step{it}/"
        assert abs(loss.item() - float(g[p + "loss"])) <= 2e-6 * abs(float(g[p + "loss"]))
        np.testing.assert_allclose(aux["q_s"].numpy(), g[p + "q_s"], atol=5e-6)
        np.testing.assert_allclose(aux["q_next_online"].numpy(), g[p + "q_next_online"], atol=5e-6)
        assert (aux["best"].numpy() == g[p + "best"]).all()
        for n in tr.names:
            ref_l2 = float(g[p + f"grad/{n}/l2"])
            assert abs(grads[n].double().norm().item() - ref_l2) <= 1e-4 * ref_l2 + 1e-12, n
            np.testing.assert_allclose(_sample(tr.sd[n]), g[p + f"param/{n}/sample"], rtol=1e-5, atol=1e-7,
                                       err_msg=n)
        for k in tr.sd:
            if k.startswith("resnet.") and (k.endswith("running_mean") or k.endswith("running_var")):
                np.testing.assert_allclose(_sample(tr.sd[k]), g[p + f"buffer/{k}/sample"], rtol=1e-5, atol=1e-7,
                                           err_msg=k)


# ---------------------------------------------------------------------------------------------------------
# plain-C restatement (oracle/td_adam_ref.c -> oracle/_ref/libtdref.so): a second, torch-free oracle for the
# index / fp32 pieces, pinned to the same vectors from the reference's own process_batch closure
# ---------------------------------------------------------------------------------------------------------
def _tdref():
    import ctypes as C
    import subprocess
    odir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    so = os.path.join(odir, "_ref", "libtdref.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", odir])
    lib = C.CDLL(so)
    P = C.c_void_p
    lib.td_ref.argtypes = [P] * 8 + [C.c_int] * 3 + [C.c_float] + [C.c_int] * 6 + [P] * 4
    lib.td_ref.restype = C.c_int
    lib.td_ref_f32.argtypes = lib.td_ref.argtypes
    lib.td_ref_f32.restype = C.c_int
    lib.adam_ref.argtypes = [P, P, P, P, C.c_long, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    lib.adam_ref.restype = C.c_int
    return lib


def test_c_restatement_of_td_loss_matches_reference_goldens():
    lib = _tdref()
    z = np.load(os.path.join(GOLD, "td_branches.npz"))
    names = sorted({k.split("/")[0] for k in z.keys()})
    assert len(names) == 7
    ptr = lambda a: a.ctypes.data  # noqa: E731
    for name in names:
        gamma, rect, linear, use_valid, value_learning, gt_mode, A = z[f"{name}/cfg"]
        q_s, q_no, q_nt = (np.ascontiguousarray(z[f"{name}/{k}"], dtype=np.float32) for k in ("q_s", "q_no", "q_nt"))
        # CONFIDENCE_REWARD case: float64 scores in the fixture, `.float()` = fp32 labels for td_ref_f32
        flt = z[f"{name}/rew"].dtype.kind == "f"
        ldt = np.float32 if flt else np.int64
        act = np.ascontiguousarray(z[f"{name}/act"], dtype=np.int64)
        rew, valid = (np.ascontiguousarray(z[f"{name}/{k}"], dtype=ldt) for k in ("rew", "valid"))
        B, Cc = rew.shape
        gt = np.ascontiguousarray(np.broadcast_to(z[f"{name}/gt"].reshape(B, -1), (B, Cc)), dtype=np.float64)
        dq, y = np.empty_like(q_s), np.empty((B, Cc), np.float32)
        best, loss = np.empty((B, Cc), np.int64), np.zeros(1, np.float32)
        fn = lib.td_ref_f32 if flt else lib.td_ref
        rc = fn(ptr(q_s), ptr(q_no), ptr(q_nt), ptr(act), ptr(rew), ptr(rew), ptr(valid), ptr(gt), B, Cc,
                int(A), float(gamma), 1, int(rect), int(linear), int(use_valid), int(gt_mode),
                int(value_learning), ptr(dq), ptr(y), ptr(best), ptr(loss))
        assert rc == 0
        ref_loss = float(z[f"{name}/loss"])
        if np.isnan(ref_loss):
            assert np.isnan(loss[0]), name
        else:
            assert abs(loss[0] - ref_loss) <= 1e-6 * abs(ref_loss) + 1e-9, (name, loss[0], ref_loss)
        np.testing.assert_allclose(dq, z[f"{name}/dq"], rtol=0, atol=1e-7, equal_nan=True, err_msg=name)
        if not gt_mode:                                   # arg-max against the torch oracle: integers, exact
            cfg = qstep.StepConfig(GAMMA=float(gamma), LOSS_CLIP="rect" if rect else "none", LINEAR=bool(linear),
                                   action_dim=int(A))
            rew_t = torch.from_numpy(z[f"{name}/rew"])
            b_ref, y_ref = qstep.td_targets(torch.from_numpy(q_no), torch.from_numpy(q_nt), rew_t, rew_t, cfg)
            assert (best == b_ref.numpy()).all(), name
            np.testing.assert_array_equal(y, y_ref.numpy(), err_msg=name)      # same fp32 operations: bit-exact


def test_c_restatement_of_adam_matches_torch():
    lib = _tdref()
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(1000, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-4)
    p = p0.numpy().copy()
    m, v = np.zeros_like(p), np.zeros_like(p)
    for step in range(1, 5):
        gr = (torch.randn(1000, generator=g) * 0.01)
        ref.grad = gr.clone()
        opt.step()
        gnp = gr.numpy().copy()
        assert lib.adam_ref(p.ctypes.data, gnp.ctypes.data, m.ctypes.data, v.ctypes.data, 1000, 1e-4, 0.9, 0.999,
                            1e-8, step) == 0
    np.testing.assert_allclose(p, ref.detach().numpy(), rtol=0, atol=2e-7)


@pytest.mark.skipif(not os.path.isdir(os.path.join(os.environ.get("VDQN_REFERENCE", "/root/reference"), "archs")),
                    reason="needs the reference checkout (this container)")
def test_oracle_reproduces_the_reference_run_train_end_to_end():
    """the reference's own `train_q_network.run_train` (imported unmodified, three steps on a
    table built from the committed mini data set, recording Adam + DataLoader) against the oracle started
    from the recorded initial parameters and stepped on the recorded batches, target sync included."""
    from oracle import crosscheck_run_train
    worst = crosscheck_run_train.run(verbose=False)
    assert worst < 1e-4


def test_c_restatement_agrees_with_the_torch_oracle_on_random_cases():
    """The two oracles against each other over random shapes and switches (including B = 0, B = 1, ties in
    the arg-max, one action, all-terminal rows): arg-max and y bit-exact, dQ to 1e-7, loss to 1e-6."""
    lib = _tdref()
    g = torch.Generator().manual_seed(123)
    ptr = lambda a: a.ctypes.data  # noqa: E731
    for case in range(40):
        B = [0, 1, 2, 7, 64][case % 5]
        A = 1 if case % 7 == 0 else 3
        cfg = qstep.StepConfig(GAMMA=0.99 if case % 3 else 0.9, LOSS_CLIP="rect" if case % 2 else "none",
                               LINEAR=(case % 4 == 3), REMOVE_BEFORE_REWARD=(case % 5 == 4), action_dim=A)
        q_s, q_no, q_nt = (torch.randn(B, 5, A, generator=g) for _ in range(3))
        if B > 1:
            q_no[0, :, :] = 0.25                          # ties: the first maximum must win
        act = torch.randint(0, A, (B,), generator=g)
        rew = (torch.rand(B, 5, generator=g) < 0.3).long()
        term = rew.clone() if case % 2 else torch.ones_like(rew)
        valid = (torch.rand(B, 5, generator=g) < 0.7).long()
        dq = np.zeros((B, 5, A), np.float32)
        y, best, loss = np.zeros((B, 5), np.float32), np.zeros((B, 5), np.int64), np.zeros(1, np.float32)
        arrs = [np.ascontiguousarray(t.numpy()) for t in (q_s, q_no, q_nt, act, rew, term, valid)]
        gt = np.zeros((max(B, 1), 5), np.float64)
        rc = lib.td_ref(*[ptr(a) for a in arrs], ptr(gt), B, 5, A, float(cfg.GAMMA), 1,
                        int(cfg.LOSS_CLIP == "rect"), int(cfg.LINEAR), int(cfg.REMOVE_BEFORE_REWARD), 0, 0,
                        ptr(dq), ptr(y), ptr(best), ptr(loss))
        assert rc == 0
        if B == 0:
            continue
        qs = q_s.clone().requires_grad_(True)
        l_ref, aux = qstep.td_loss(qs, q_no, q_nt, act, rew, term, valid, cfg)
        l_ref.backward()
        assert (best == aux["best"].numpy()).all(), case
        np.testing.assert_array_equal(y, aux["y"].numpy(), err_msg=str(case))
        np.testing.assert_allclose(dq, qs.grad.numpy(), rtol=0, atol=1e-7, err_msg=str(case))
        assert abs(loss[0] - l_ref.item()) <= 1e-6 * max(1.0, abs(l_ref.item())), case


def test_rounding_point_oracle_is_the_same_graph():
    """oracle/qstep_bf16.py (the tight-bar oracle of tests/test_gpu_parity_full.py) is qstep's graph with
    rounding inserted: with the rounding functions switched off it reproduces qstep (BatchNorm folding is
    exact up to fp32 round-off), with them on it sits at bf16 distance from it."""
    from oracle import qstep_bf16 as qb
    sd = qstep.init_state(seed=4, randomize_bn=True)
    batch = qstep.synthetic_batch(2, seed=1)
    l0, g0, a0 = qstep.OracleTrainer(sd).loss_and_grads(batch)
    l1, g1, a1 = qb.EmulatedTrainer(sd).loss_and_grads(batch)
    assert abs(l1.item() - l0.item()) <= 1e-2 * abs(l0.item())
    assert (a1["q_s"] - a0["q_s"]).abs().max().item() <= 1e-2
    _, rel = qb.grad_report(g1, g0)
    assert 1e-2 < rel < 0.2                                  # it does round, and only that
    saved = (qb._RoundBoth.apply, qb._RoundValue.apply, qb._RoundGrad.apply)
    try:
        for c in (qb._RoundBoth, qb._RoundValue, qb._RoundGrad):
            c.apply = staticmethod(lambda x: x)
        l2, g2, a2 = qb.EmulatedTrainer(sd).loss_and_grads(batch)
    finally:
        qb._RoundBoth.apply, qb._RoundValue.apply, qb._RoundGrad.apply = saved
    assert abs(l2.item() - l0.item()) <= 1e-5 * abs(l0.item())
    rows, rel2 = qb.grad_report(g2, g0)
    assert rel2 <= 1e-4, rel2
    assert max(abs(v[1]) for v in rows.values()) <= 1e-4
