"""Pin oracle/inverse.py against the reference's own `archs/inverse_action2.model` and write
tests/golden/inverse_b4.npz (inputs seed, outputs).  TEST INFRASTRUCTURE ONLY; needs /root/reference.

The reference class is imported unmodified (`resnet18(pretrained=True)` patched to `weights=None`: no
network), loaded with `oracle.inverse.init_state(seed)` (strict), put in eval mode as the labelling
script does (dataset/process_episodes_real.py:93-96) and run on random frame pairs.

usage:  python -m oracle.make_inverse_goldens [--out tests/golden]
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

from . import inverse

REF = os.environ.get("VDQN_REFERENCE", "/root/reference")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    a = ap.parse_args()
    import torchvision.models as tvm
    orig = tvm.resnet18
    tvm.resnet18 = lambda pretrained=False, **kw: orig(weights=None, **kw)
    sys.path.insert(0, REF)
    try:
        from archs import inverse_action2
        m = inverse_action2.model()
    finally:
        sys.path.remove(REF)
        tvm.resnet18 = orig
    sd = inverse.init_state(seed=7)
    missing = m.load_state_dict(sd, strict=False)
    # the only keys our state does not carry are BN `num_batches_tracked` counters
    assert all(k.endswith("num_batches_tracked") for k in missing.missing_keys), missing.missing_keys
    assert not missing.unexpected_keys, missing.unexpected_keys
    m.eval()
    g = torch.Generator().manual_seed(21)
    k = torch.randn(4, 3, 224, 224, generator=g)
    k1 = torch.randn(4, 3, 224, 224, generator=g)
    with torch.no_grad():
        enc_ref, y_ref = m(k, k1)
        enc, y = inverse.forward(sd, k, k1)
    assert torch.allclose(enc, enc_ref, atol=1e-6) and torch.allclose(y, y_ref, atol=1e-5), \
        ((enc - enc_ref).abs().max(), (y - y_ref).abs().max())
    print("oracle == reference: max |dy|", (y - y_ref).abs().max().item())
    path = os.path.join(a.out, "inverse_b4.npz")
    np.savez_compressed(path, seed=7, data_seed=21, encoding=enc_ref.numpy(), y=y_ref.numpy())
    print("wrote", path)


if __name__ == "__main__":
    main()
