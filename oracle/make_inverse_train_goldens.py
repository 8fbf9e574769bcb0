"""Pin oracle/inverse.py's TRAINING step against the reference's own `train_inverse_model.model`,
`nn.CrossEntropyLoss` and `torch.optim.Adam` and write tests/golden/inverse_train_b4.npz.
TEST INFRASTRUCTURE ONLY; needs /root/reference.

`train_inverse_model.py` is imported unmodified (stubs for the absent `matplotlib` that its
dataloader import pulls in; `resnet18(pretrained=True)` patched to `weights=None`: no network;
absl FLAGS parsed with their defaults: bottleneck 3, lr 1e-4, weight decay 0).  The module is loaded
with `oracle.inverse.init_state(seed)`, put in train mode as `train()` does (:87) and stepped three
times with the loop body of :93-110.  The dropout draw of each step is recorded with a forward hook
on `dropout1` (keep = output != 0 wherever the input is > 0; where the input is 0 the draw cannot
influence anything) and handed to the oracle, which must reproduce loss, logits, every gradient and
the parameters after each Adam step.  Finally the reference's own `train()` function (:85-140) is run on the same
batches and dropout seed and must end on bit-identical parameters.

usage:  python -m oracle.make_inverse_train_goldens [--out tests/golden]
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch

from . import inverse
from .make_goldens import summarize

REF = os.environ.get("VDQN_REFERENCE", "/root/reference")
STEPS, B = 3, 4


def import_reference_trainer():
    import torchvision.models as tvm
    orig = tvm.resnet18
    tvm.resnet18 = lambda pretrained=False, **kw: orig(weights=None, **kw)
    stubbed = []
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
            stubbed.append(name)
    if "matplotlib" in stubbed:
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(np, "int"):
        np.int = int
    sys.path.insert(0, REF)
    try:
        import train_inverse_model as T
        if not T.FLAGS.is_parsed():
            T.FLAGS(["make_inverse_train_goldens"])
        m = T.model()
    finally:
        sys.path.remove(REF)
        tvm.resnet18 = orig
        for name in stubbed:
            sys.modules.pop(name, None)
    return T, m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    a = ap.parse_args()
    T, m = import_reference_trainer()
    sd = inverse.init_state(seed=7)
    missing = m.load_state_dict(sd, strict=False)
    assert all(k.endswith("num_batches_tracked") for k in missing.missing_keys), missing.missing_keys
    assert not missing.unexpected_keys
    names = [n for n, p in m.named_parameters() if p.requires_grad]
    assert tuple(names) == inverse.TRAINABLE, names
    opt = torch.optim.Adam(m.parameters(), lr=T.FLAGS.lr, weight_decay=T.FLAGS.weight_decay)   # :176
    seen = {}
    m.dropout1.register_forward_hook(lambda mod, inp, out: seen.update(inp=inp[0].detach(), out=out.detach()))
    oracle = inverse.InverseOracleTrainer(sd, lr=T.FLAGS.lr)
    g = torch.Generator().manual_seed(33)
    torch.manual_seed(5)                       # the reference's dropout draws
    out = {"seed": 7, "data_seed": 33, "steps": STEPS, "batch": B, "lr": T.FLAGS.lr}
    worst = 0.0
    m.train()                                                                                    # :87
    batches = []
    for s in range(STEPS):
        k = torch.randn(B, 3, 224, 224, generator=g)
        k1 = torch.randn(B, 3, 224, 224, generator=g)
        act = torch.randint(0, 3, (B,), generator=g)
        batches.append((k, k1, act, None, None, None))
        opt.zero_grad()                                                                          # :94
        y = m(k, k1)                                                                             # :97
        loss = torch.nn.CrossEntropyLoss()(y, act)                                               # :100-101
        loss.backward()                                                                          # :109
        keep = torch.where(seen["inp"] > 0, (seen["out"] != 0), torch.ones_like(seen["inp"], dtype=torch.bool))
        grads_ref = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.requires_grad}
        opt.step()                                                                               # :110
        o_loss, o_grads, o_y, o_correct = oracle.step(k, k1, act, keep.to(torch.uint8))
        assert torch.allclose(o_y, y.detach(), atol=1e-6), (o_y - y).abs().max()
        assert abs(o_loss.item() - loss.item()) < 1e-6
        for n in names:
            d = (o_grads[n] - grads_ref[n]).abs().max().item() / max(grads_ref[n].abs().max().item(), 1e-12)
            worst = max(worst, d)
            assert d < 1e-4, (s, n, d)
        ref_p = dict(m.named_parameters())
        for n in names:
            d = (oracle.sd[n] - ref_p[n].detach()).abs().max().item()
            assert d < 2e-6, (s, n, d)
        out[f"keep{s}"] = keep.numpy().astype(np.uint8)
        out[f"y{s}"] = y.detach().numpy()
        out[f"loss{s}"] = np.float32(loss.item())
        out[f"correct{s}"] = int((y.argmax(1) == act).sum())
        # per tensor: l2 norm, sum and a 64-element strided sample (the format of step_b8_*.npz)
        out.update(summarize({f"s{s}/grad/{n}": grads_ref[n] for n in names}))
        out.update(summarize({f"s{s}/param/{n}": ref_p[n] for n in names}))
    print("oracle == reference over", STEPS, "steps; worst relative gradient deviation", worst)
    # the loop body above is a transcription of train() (:93-110); run the reference's OWN train() on the same
    # batches and dropout seed with a fresh copy of the module: it must end on the same parameters, bit for bit
    import torchvision.models as tvm
    orig = tvm.resnet18
    tvm.resnet18 = lambda pretrained=False, **kw: orig(weights=None, **kw)
    try:
        m2 = T.model()
    finally:
        tvm.resnet18 = orig
    m2.load_state_dict(sd, strict=False)
    opt2 = torch.optim.Adam(m2.parameters(), lr=T.FLAGS.lr, weight_decay=T.FLAGS.weight_decay)
    torch.manual_seed(5)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        T.train(m2, torch.device("cpu"), batches, opt2, 1, None, None, 0)
    p1, p2 = dict(m.named_parameters()), dict(m2.named_parameters())
    assert all(torch.equal(p1[n], p2[n]) for n in names), "transcribed loop != the reference's train()"
    print("the reference's own train() ends on identical parameters")
    path = os.path.join(a.out, "inverse_train_b4.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
