"""Repeat the same forward+backward many times on identical inputs and report, per tensor, the
largest run-to-run deviation.  Atomics make tiny (1e-6-ish) differences legitimate; anything
large is a race."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import qstep
from video_dqn_b200.qnet import HabitatDQNMultiAction

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
R = int(sys.argv[2]) if len(sys.argv) > 2 else 30
sd = qstep.init_state(seed=4, randomize_bn=True)
def mk():
    m = HabitatDQNMultiAction(3, 5, extra_capacity=True, panorama=False); m.load_state_dict(sd); return m.to(dev)
model, target = mk(), mk()
model.set_train(); target.eval()
batch = [t.to(dev) for t in qstep.synthetic_batch(B, seed=1)]
cfg = qstep.StepConfig()
ref = None
worst = {}
for it in range(R):
    model.zero_grad()
    q_s = model(batch[0]); q_nt = target(batch[1]); q_no = model(batch[1])
    loss, _ = qstep.td_loss(q_s, q_no, q_nt, batch[2], batch[3], batch[4], batch[6], cfg)
    loss.backward()
    torch.cuda.synchronize()
    cur = {"q_s": q_s.detach().clone(), "q_nt": q_nt.detach().clone(), "loss": loss.detach().clone()}
    cur.update({n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None})
    if ref is None:
        ref = cur; continue
    for k in cur:
        d = (cur[k] - ref[k]).abs().max().item() / (ref[k].abs().max().item() + 1e-30)
        if d > worst.get(k, (0, 0))[0]:
            worst[k] = (d, it)
for k, (d, it) in sorted(worst.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{d:10.3e} (run {it:2d}) {k}")
print("max deviation", max(v[0] for v in worst.values()))
