"""Cross-check of the oracle against the reference's OWN `run_train`, end to end:
`train_q_network.run_train(config)` is imported from /root/reference unmodified and run for a few steps on
CPU over a 16-row table built from the committed mini data set; `torch.optim.Adam` is replaced by a
recording subclass (initial parameters, the gradients of every step, the parameters after every step) and
`torch.utils.data.DataLoader` by a recording, in-process one (the batches actually drawn).  The oracle
(`qstep.OracleTrainer`) is then started from the recorded initial parameters and stepped on the recorded
batches; it must reproduce every gradient and every parameter -- including the hard target sync
(TARGET_UPDATE_INTERVAL = 2 here, so the second step bootstraps from the synced target).
TEST INFRASTRUCTURE ONLY; needs /root/reference.

Shims (none touches the reference's source): stub modules for the absent `gibson_info`, `visualize_value`,
`matplotlib`; `np.int`; `resnet18(pretrained=True)` -> `weights=None`; DataLoader workers forced to 0.

usage:  python -m oracle.crosscheck_run_train        (prints the worst deviations, exits non-zero on mismatch)
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

from . import qstep

REF = os.environ.get("VDQN_REFERENCE", "/root/reference")
MINI = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "realdata")
STEPS = 3


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def run(verbose: bool = True) -> float:
    import pandas as pd
    import torch.utils.data as tud
    import torchvision.models as tvm
    work = tempfile.mkdtemp(prefix="vdqn_rt_")
    cwd = os.getcwd()
    orig_resnet, orig_adam, orig_loader = tvm.resnet18, torch.optim.Adam, tud.DataLoader
    stubbed = []
    rec = {"batches": [], "grads": [], "params": [], "init": None, "names": None}
    try:
        # 16-row table (the trainer's batch is 16 with drop_last, train_q_network.py:98): the four rows x 4
        shutil.copytree(os.path.join(MINI, "ep0"), os.path.join(work, "ep0"))
        t = pd.read_feather(os.path.join(MINI, "data.feather"))
        pd.concat([t] * 4, ignore_index=True).to_feather(os.path.join(work, "data.feather"))
        for name, attrs in (("gibson_info", dict(get_houses=None, class_labels=[], get_house=None)),
                            ("visualize_value", dict(build_map_gibson=None)),
                            ("matplotlib", {}), ("matplotlib.pyplot", {})):
            if name not in sys.modules:
                _stub(name, **attrs)
                stubbed.append(name)
        if "matplotlib" in stubbed:
            sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        if not hasattr(np, "int"):
            np.int = int
        tvm.resnet18 = lambda pretrained=False, **kw: orig_resnet(weights=None, **kw)

        class RecAdam(orig_adam):
            def __init__(self, params, *a, **kw):
                params = list(params)
                super().__init__(params, *a, **kw)
                rec["plist"] = params
                rec["init"] = [p.detach().clone() for p in params]

            def step(self, closure=None):
                rec["grads"].append([None if p.grad is None else p.grad.detach().clone() for p in rec["plist"]])
                out = super().step(closure)
                rec["params"].append([p.detach().clone() for p in rec["plist"]])
                return out

        class RecLoader(orig_loader):
            def __init__(self, dataset, **kw):
                kw["num_workers"] = 0
                super().__init__(dataset, **kw)

            def __iter__(self):
                for b in super().__iter__():
                    rec["batches"].append([x.clone() for x in b])
                    yield b

        sys.path.insert(0, REF)
        os.chdir(work)                                   # the table stores frame paths relative to its directory
        import train_q_network as T
        T.optim.Adam = RecAdam
        T.data.DataLoader = RecLoader
        writer = types.SimpleNamespace(add_scalar=lambda *a, **k: None, add_image=lambda *a, **k: None)
        cfg = types.SimpleNamespace(
            device="cpu", folder=work, writer=writer, SEED=4, VISUALIZATION_DATA_ROOT="", DATASET="data.feather",
            CONFIDENCE_REWARD=False, VALUE_LEARNING=False, USE_INVERSE_ACTIONS=True, PREVIOUS_IMAGES=False,
            ONE_ACTION=False, ARCHITECTURE="extra_capacity", PANORAMA=False, LOSS_CLIP="rect", LEARNING_RATE=1e-4,
            BOOTSTRAP=False, NUM_STEPS=STEPS, TARGET_UPDATE_INTERVAL=2, CHECKPOINT_INTERVAL=10 ** 9,
            TRAIN_ON_GROUND_TRUTH=False, GAMMA=0.99, LINEAR=False, REMOVE_BEFORE_REWARD=False)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            T.run_train(cfg)                             # the reference's own loop, :84-250
    finally:
        os.chdir(cwd)
        if REF in sys.path:
            sys.path.remove(REF)
        tvm.resnet18, torch.optim.Adam, tud.DataLoader = orig_resnet, orig_adam, orig_loader
        for name in stubbed:
            sys.modules.pop(name, None)
        for name in ("train_q_network",):
            sys.modules.pop(name, None)
        shutil.rmtree(work, ignore_errors=True)
    assert len(rec["grads"]) == STEPS and len(rec["batches"]) >= STEPS, (len(rec["grads"]), len(rec["batches"]))
    # ---- the oracle on the recorded initial parameters and batches
    layout_names = _parameter_names()
    assert len(layout_names) == len(rec["init"]) == 70
    sd = qstep.init_state(seed=0)                       # right keys / default BatchNorm buffers; values replaced below
    for n, v in zip(layout_names, rec["init"]):
        sd[n] = v.clone()
    for k in list(sd):                                   # fresh BatchNorm buffers of an untrained torchvision net
        if k.endswith("running_mean"):
            sd[k] = torch.zeros_like(sd[k])
        elif k.endswith("running_var"):
            sd[k] = torch.ones_like(sd[k])
    oracle = qstep.OracleTrainer(sd, qstep.StepConfig(TARGET_UPDATE_INTERVAL=2))
    idx = {n: i for i, n in enumerate(layout_names)}
    worst = 0.0
    for s in range(STEPS):
        loss, grads, _aux = oracle.step(tuple(rec["batches"][s]))
        for n in oracle.names:
            g_ref, p_ref = rec["grads"][s][idx[n]], rec["params"][s][idx[n]]
            e_g = ((grads[n] - g_ref).double().norm() / (g_ref.double().norm() + 1e-30)).item()
            e_p = ((oracle.sd[n] - p_ref).double().norm() / (p_ref.double().norm() + 1e-30)).item()
            worst = max(worst, e_g, e_p)
            assert e_g < 1e-4 and e_p < 1e-6, (s, n, e_g, e_p)
        assert rec["grads"][s][idx["resnet.fc.weight"]] is None       # resnet.fc never receives a gradient
        if verbose:
            print(f"step {s}: oracle loss {loss.item():.9f}; worst relative deviation so far {worst:.2e}")
    return worst


def _parameter_names():
    """`model.parameters()` order of the reference module (tests/golden/reference_module_layout.json, dumped
    from the reference class)"""
    import json
    lay = json.load(open(os.path.join(os.path.dirname(MINI), "reference_module_layout.json")))
    return lay["named_parameters"]


if __name__ == "__main__":
    w = run()
    print(f"oracle == the reference's own run_train over {STEPS} steps (worst relative deviation {w:.2e})")
