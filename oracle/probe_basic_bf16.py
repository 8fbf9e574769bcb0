"""What PyTorch's OWN bf16 autocast does to the `basic` architecture's training step (train-mode
BatchNorm, B = 8), against its fp32 run on the same batch and weights -- the yardstick for the CUDA path's
parity bars in tests/test_gpu_basic_train.py.  TEST INFRASTRUCTURE ONLY (CPU).

Measured here (torch 2.11, CPU autocast):  max |dQ| 1.2e-2, loss 0.60907 -> 0.60974, gradient global
rel-L2 0.296, worst per-tensor cosine 0.881 (resnet.layer1.1.bn2.bias), median 0.937.
The CUDA path measures max |dQ| 1.6e-2, rel-L2 0.29, worst cosine 0.866 on the same step: the same
regime.  Why it is so much looser than the shipped eval-mode path (rel-L2 0.08): the batch-statistics
BatchNorm backward is a projection, dx = g*rstd*(dy - mean(dy) - xhat*mean(dy*xhat)); where dy is
dominated by its per-channel mean the difference is small against the bf16 rounding of dy itself, and
the loss of relative precision compounds over the 20 BatchNorms between the head and the stem.

usage:  python -m oracle.probe_basic_bf16
"""
import torch

from . import qstep
torch.set_num_threads(8)
sd = qstep.init_state_basic(seed=4, num_frames=1)
B = 8
batch = qstep.synthetic_batch(B, seed=1)
def run(autocast):
    tr = qstep.BasicOracleTrainer(sd)
    if autocast:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            return tr.loss_and_grads(batch)
    return tr.loss_and_grads(batch)
l0, g0, a0 = run(False)
l1, g1, a1 = run(True)
def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()
print("loss", l0.item(), l1.item(), "dQ", (a0["q_s"] - a1["q_s"].float()).abs().max().item())
num = den = 0
cs = []
for n in g0:
    c = cos(g0[n], g1[n].float()); cs.append((c, n))
    num += (g0[n].double() - g1[n].double()).pow(2).sum().item(); den += g0[n].double().pow(2).sum().item()
cs.sort()
print("rel-L2", (num / den) ** 0.5)
print("worst", cs[:8])
print("median", cs[len(cs) // 2])
