"""Pin oracle.qstep.q_forward_basic against the reference's own module built with
extra_capacity=False (the `basic` architecture) and write tests/golden/basic_b4.npz.
TEST INFRASTRUCTURE ONLY; needs /root/reference.

usage:  python -m oracle.make_basic_goldens [--out tests/golden]
"""
from __future__ import annotations

import argparse
import os

import numpy as np
import torch

from . import qstep
from .make_goldens import import_reference_model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    a = ap.parse_args()
    Ref, restore = import_reference_model()
    out = {}
    try:
        for tag, panorama, F in (("f1", False, 1), ("f4", True, 4)):
            sd = qstep.init_state_basic(seed=4, num_frames=F)
            m = Ref(3, 5, extra_capacity=False, panorama=panorama)
            res = m.load_state_dict(sd, strict=False)
            assert not res.unexpected_keys and all(k.endswith("num_batches_tracked") for k in res.missing_keys), res
            m.eval()
            g = torch.Generator().manual_seed(31 + F)
            B = 4 if F == 1 else 2
            x = torch.randn(B, F, 3, 224, 224, generator=g) if F > 1 else torch.randn(B, 3, 224, 224, generator=g)
            with torch.no_grad():
                q_ref = m(x)
                q = qstep.q_forward_basic(sd, x)
            assert torch.allclose(q, q_ref, atol=2e-6), (tag, (q - q_ref).abs().max())
            print(tag, "oracle == reference: max |dQ|", (q - q_ref).abs().max().item())
            out[f"{tag}/q"] = q_ref.numpy()
            out[f"{tag}/data_seed"] = np.array(31 + F)
    finally:
        restore()
    path = os.path.join(a.out, "basic_b4.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
