"""Role profile of the fused stem + max-pool kernel (instrumented build, see tools/role_profile.py)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from video_dqn_b200 import ops, _lib
lib = _lib.load()
fn = lib.vdqn_debug_role_profile
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]; fn.restype = ctypes.c_int
ROLES = ["prod0", "prod1", "prod2", "-", "mma"] + [f"drain{i}" for i in range(8)] + ["pool0", "pool1"]
g = torch.Generator(device="cuda").manual_seed(0)
bf = torch.bfloat16
N = 768
xs = torch.randn(N, 112, 112, 16, device="cuda", generator=g).to(bf)
ws = (torch.randn(64, 4, 4, 16, device="cuda", generator=g) / 16).to(bf)
sh = torch.randn(64, device="cuda", generator=g)
p = torch.empty(N, 56, 56, 64, device="cuda", dtype=bf)
idx = torch.empty(256, 56, 56, 64, device="cuda", dtype=torch.uint8)
for nidx in (256, 0):
    for _ in range(3):
        ops.conv_gemm(xs, ws, 1, 2, 1, shift=sh, relu=True, pool_out=p, pool_idx=idx, pool_idx_images=nidx)
    torch.cuda.synchronize()
    buf = np.zeros(160 * 16 * 4, dtype=np.uint64)
    assert fn(buf.ctypes.data, buf.size) == 0
    prof = buf.reshape(160, 16, 4)[:ops.num_sms()].astype(np.float64)
    print(f"== fused stem + pool, idx for {nidx} of {N} frames")
    for r, rn in enumerate(ROLES):
        q = prof[:, r]
        tiles = q[:, 3].mean()
        if tiles == 0:
            continue
        tot = q[:, 2].mean()
        print(f"  {rn:7s} tiles/cta {tiles:6.1f}  loop {tot / tiles:8.0f} cyc/tile   wait_a {q[:, 0].mean() / tiles:8.0f}"
              f"   wait_b {q[:, 1].mean() / tiles:8.0f}   busy {(tot - q[:, 0].mean() - q[:, 1].mean()) / tiles:8.0f}")
