"""Pin the oracle against the reference itself and write tests/golden/*.npz.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference);
the GPU box never runs this -- it only reads the committed fixtures.

What is executed from the reference, unmodified and in place:
  * ``archs/HabitatDQNMultiAction.py`` is imported (with
    ``torchvision.models.resnet18(pretrained=True)`` patched to ``weights=None``
    because there is no network),
  * the body of the closure ``process_batch`` is lifted out of
    ``train_q_network.py`` with ``ast`` at run time (it cannot be imported: it
    is nested in ``run_train``) and executed against the reference model,
  * ``torch.optim.Adam`` -- the optimizer the reference constructs
    (train_q_network.py:124).
The oracle (oracle/qstep.py) must reproduce Q, loss, all 68 gradients and the
parameters after 1 and 3 steps; the vectors are then stored as the fixture.

usage:  python -m oracle.make_goldens [--out tests/golden]
"""
from __future__ import annotations

import argparse
import ast
import os
import sys
import types

import numpy as np
import torch

from . import qstep

REF = os.environ.get("VDQN_REFERENCE", "/root/reference")


def import_reference_model():
    import torchvision.models as tvm
    orig = tvm.resnet18
    tvm.resnet18 = lambda pretrained=False, **kw: orig(weights=None, **kw)
    sys.path.insert(0, REF)
    try:
        from archs.HabitatDQNMultiAction import HabitatDQNMultiAction  # noqa
    finally:
        sys.path.remove(REF)
    return HabitatDQNMultiAction, (lambda: setattr(tvm, "resnet18", orig))


def lift_process_batch(model, target_net, config):
    """Compile the reference's nested ``process_batch`` against our objects."""
    src = open(os.path.join(REF, "train_q_network.py")).read()
    tree = ast.parse(src)
    fn = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "process_batch":
            fn = node
    assert fn is not None, "process_batch not found in reference"
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"torch": torch, "model": model, "target_net": target_net, "config": config}
    exec(compile(mod, os.path.join(REF, "train_q_network.py"), "exec"), ns)
    return ns["process_batch"]


def strided_sample(t: torch.Tensor, n: int = 64) -> np.ndarray:
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step][:n].double().numpy()


def summarize(named):
    out = {}
    for k, v in named.items():
        v = v.detach()
        out[k + "/l2"] = np.float64(v.double().norm().item())
        out[k + "/sum"] = np.float64(v.double().sum().item())
        out[k + "/sample"] = strided_sample(v)
    return out


def run(out_dir: str, B: int = 8, steps: int = 3, randomize_bn: bool = True):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    RefNet, restore = import_reference_model()
    cfg = qstep.StepConfig()
    refcfg = types.SimpleNamespace(device="cpu", LINEAR=cfg.LINEAR, GAMMA=cfg.GAMMA,
                                   LOSS_CLIP=cfg.LOSS_CLIP, VALUE_LEARNING=False,
                                   REMOVE_BEFORE_REWARD=cfg.REMOVE_BEFORE_REWARD)
    sd0 = qstep.init_state(seed=4, randomize_bn=randomize_bn)

    model = RefNet(3, 5, extra_capacity=True, panorama=False)
    target = RefNet(3, 5, extra_capacity=True, panorama=False)
    restore()
    missing = model.load_state_dict(sd0, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert len(model.state_dict()) == 250, len(model.state_dict())
    target.load_state_dict(model.state_dict())
    target.eval()
    opt = torch.optim.Adam(model.parameters(), lr=cfg.LEARNING_RATE)
    process_batch = lift_process_batch(model, target, refcfg)

    oracle = qstep.OracleTrainer(sd0, cfg)
    names = qstep.grad_param_names()
    ref_named = dict(model.named_parameters())      # unique tensors, 'resnet.*' first
    ref_named = {n: ref_named[n] for n in names}

    gold = {"meta/B": np.int64(B), "meta/steps": np.int64(steps),
            "meta/randomize_bn": np.int64(randomize_bn)}
    worst = 0.0
    for it in range(steps):
        batch = qstep.synthetic_batch(B, seed=1 + it)
        # ---- reference ----
        model.set_train()
        opt.zero_grad()
        ref_loss = process_batch(batch)
        ref_loss.backward()
        ref_grads = {n: p.grad.detach().clone() for n, p in ref_named.items()}
        with torch.no_grad():
            ref_q = model(batch[0])
        opt.step()
        # ---- oracle ----
        loss, grads, aux = oracle.step(batch)

        def rel(a, b):
            return ((a - b).double().norm() / (b.double().norm() + 1e-30)).item()
        e_q = (aux["q_s"] - ref_q).abs().max().item()
        e_l = abs(loss.item() - ref_loss.item()) / abs(ref_loss.item())
        e_g = max(rel(grads[n], ref_grads[n]) for n in names)
        e_p = max(rel(oracle.sd[n], ref_named[n].detach()) for n in names)
        print(f"step {it}: ref loss {ref_loss.item():.9f} oracle {loss.item():.9f} | "
              f"max|dQ| {e_q:.2e} rel loss {e_l:.2e} worst grad rel-L2 {e_g:.2e} "
              f"worst param rel-L2 {e_p:.2e}")
        worst = max(worst, e_q, e_l, e_g, e_p)
        assert e_q < 1e-5 and e_l < 1e-5 and e_g < 1e-4 and e_p < 1e-6, "oracle != reference"

        p = f"step{it}/"
        gold[p + "loss"] = np.float64(ref_loss.item())
        gold[p + "q_s"] = ref_q.numpy().astype(np.float64)
        gold[p + "q_next_online"] = aux["q_next_online"].numpy().astype(np.float64)
        gold[p + "q_next_target"] = aux["q_next_target"].numpy().astype(np.float64)
        gold[p + "best"] = aux["best"].numpy()
        gold[p + "y"] = aux["y"].numpy().astype(np.float64)
        for k, v in summarize(ref_grads).items():
            gold[p + "grad/" + k] = v
        for k, v in summarize({n: ref_named[n] for n in names}).items():
            gold[p + "param/" + k] = v
        if it == 0:
            # the two small tensors in full: end of the backward chain and its start
            gold[p + "gradfull/top.4.weight"] = ref_grads["top.4.weight"].numpy()
            gold[p + "gradfull/resnet.bn1.weight"] = ref_grads["resnet.bn1.weight"].numpy()
            gold[p + "gradfull/resnet.bn1.bias"] = ref_grads["resnet.bn1.bias"].numpy()
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, f"step_b{B}_bn{int(randomize_bn)}.npz")
    np.savez_compressed(path, **gold)
    print(f"wrote {path} ({os.path.getsize(path)/1024:.0f} KiB); worst deviation {worst:.2e}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    ap.add_argument("--batch", type=int, default=8)
    a = ap.parse_args()
    run(a.out, a.batch, randomize_bn=True)
    run(a.out, a.batch, randomize_bn=False)
