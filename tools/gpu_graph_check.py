"""Stress the fused learner: several trials of N steps in graph mode vs eager mode from the same
init/batches; prints per-step losses.  Large trial-to-trial loss differences = a race."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import qstep
from video_dqn_b200.qnet import HabitatDQNMultiAction
from video_dqn_b200.learner import QLearner, StepConfig

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sd = qstep.init_state(seed=4, randomize_bn=True)
batches = [[t.to(dev) for t in qstep.synthetic_batch(B, seed=1 + i)] for i in range(4)]
def mk():
    m = HabitatDQNMultiAction(3, 5, extra_capacity=True, panorama=False); m.load_state_dict(sd); return m.to(dev)
for mode in (True, False):
    rows = []
    for trial in range(T):
        lr = QLearner(mk(), mk(), StepConfig(), batch_size=B, use_graph=mode)
        losses = []
        for i in range(4):
            l = lr.step(batches[i]); torch.cuda.synchronize(); losses.append(l.item())
        rows.append(losses)
        del lr
    t = torch.tensor(rows)
    print("graph" if mode else "eager", "mean", [f"{v:.6f}" for v in t.mean(0).tolist()])
    print("      spread (max-min)/mean per step", [f"{v:.2e}" for v in ((t.max(0).values - t.min(0).values) / t.mean(0)).tolist()])
    for r in rows: print("      ", [f"{v:.6f}" for v in r])
