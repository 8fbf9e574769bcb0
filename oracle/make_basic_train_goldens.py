"""Pin oracle.qstep.BasicOracleTrainer (training step of the `basic` architecture: trunk BatchNorms in
TRAIN mode) against the reference's own module built with extra_capacity=False, its own
`process_batch` (lifted with ast) and torch.optim.Adam, and write tests/golden/basic_train_b8.npz.
TEST INFRASTRUCTURE ONLY; needs /root/reference.

usage:  python -m oracle.make_basic_train_goldens [--out tests/golden]
"""
from __future__ import annotations

import argparse
import os
import types

import numpy as np
import torch

from . import qstep
from .make_goldens import import_reference_model, lift_process_batch, summarize

STEPS = 2
CASES = ((1, 8, "basic_train_b8.npz"), (4, 4, "basic_train_f4_b4.npz"))     # (frames, batch, fixture)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    a = ap.parse_args()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for nf, B, fname in CASES:
        run(a.out, nf, B, fname)


def frames_batch(B, nf, seed):
    """synthetic quadruplets with F frames per state (panorama / previous-images layout [B,F,3,224,224])"""
    b = list(qstep.synthetic_batch(B, seed=seed))
    if nf > 1:
        g = torch.Generator().manual_seed(1000 + seed)
        b[0] = torch.randn(B, nf, 3, 224, 224, generator=g)
        b[1] = torch.randn(B, nf, 3, 224, 224, generator=g)
    return tuple(b)


def run(out_dir, nf, B, fname):
    Ref, restore = import_reference_model()
    try:
        model = Ref(3, 5, extra_capacity=False, panorama=(nf == 4))
        target = Ref(3, 5, extra_capacity=False, panorama=(nf == 4))
    finally:
        restore()
    cfg = qstep.StepConfig()
    refcfg = types.SimpleNamespace(device="cpu", LINEAR=cfg.LINEAR, GAMMA=cfg.GAMMA, LOSS_CLIP=cfg.LOSS_CLIP,
                                   VALUE_LEARNING=False, REMOVE_BEFORE_REWARD=cfg.REMOVE_BEFORE_REWARD)
    sd0 = qstep.init_state_basic(seed=4, num_frames=nf)
    res = model.load_state_dict(sd0, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    target.load_state_dict(model.state_dict())
    target.eval()
    opt = torch.optim.Adam(model.parameters(), lr=cfg.LEARNING_RATE)
    process_batch = lift_process_batch(model, target, refcfg)
    oracle = qstep.BasicOracleTrainer(sd0, cfg)
    names = qstep.grad_param_names_basic()
    ref_named = {n: p for n, p in model.named_parameters() if n in names}
    assert list(ref_named) == names
    gold = {"meta/B": np.int64(B), "meta/steps": np.int64(STEPS), "meta/frames": np.int64(nf)}
    worst = 0.0
    for it in range(STEPS):
        batch = frames_batch(B, nf, 1 + it)
        model.set_train()                                   # BN stays in train mode for this architecture
        assert model.resnet.bn1.training
        opt.zero_grad()
        ref_loss = process_batch(batch)
        ref_loss.backward()
        ref_grads = {n: p.grad.detach().clone() for n, p in ref_named.items()}
        opt.step()
        loss, grads, aux = oracle.step(batch)

        def rel(x, y):
            return ((x - y).double().norm() / (y.double().norm() + 1e-30)).item()
        e_l = abs(loss.item() - ref_loss.item()) / abs(ref_loss.item())
        e_g = max(rel(grads[n], ref_grads[n]) for n in names)
        e_p = max(rel(oracle.sd[n], ref_named[n].detach()) for n in names)
        bufs = dict(model.named_buffers())
        e_b = max(rel(oracle.sd[k].float(), v.float()) for k, v in bufs.items() if k.startswith("resnet."))
        print(f"step {it}: ref loss {ref_loss.item():.9f} oracle {loss.item():.9f} | rel loss {e_l:.2e} "
              f"worst grad {e_g:.2e} worst param {e_p:.2e} worst buffer {e_b:.2e}")
        worst = max(worst, e_l, e_g, e_p, e_b)
        assert e_l < 1e-5 and e_g < 2e-4 and e_p < 1e-6 and e_b < 1e-6, "oracle != reference"
        assert int(bufs["resnet.bn1.num_batches_tracked"]) == 2 * nf * (it + 1)     # two train-mode forwards x F frames
        p = f"step{it}/"
        gold[p + "loss"] = np.float64(ref_loss.item())
        gold[p + "q_s"] = aux["q_s"].numpy().astype(np.float64)
        gold[p + "q_next_online"] = aux["q_next_online"].numpy().astype(np.float64)
        gold[p + "q_next_target"] = aux["q_next_target"].numpy().astype(np.float64)
        gold[p + "best"] = aux["best"].numpy()
        for k, v in summarize(ref_grads).items():
            gold[p + "grad/" + k] = v
        for k, v in summarize({n: ref_named[n] for n in names}).items():
            gold[p + "param/" + k] = v
        for k, v in summarize({k: v.float() for k, v in bufs.items()
                               if k.startswith("resnet.") and not k.endswith("num_batches_tracked")}).items():
            gold[p + "buffer/" + k] = v
    path = os.path.join(out_dir, fname)
    np.savez_compressed(path, **gold)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB); worst deviation {worst:.2e}")


if __name__ == "__main__":
    main()
