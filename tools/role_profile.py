"""Which warp role bounds the halo convolution kernels?  Needs the instrumented library:

    VDQN_NVCC_FLAGS=-DVDQN_ROLE_PROFILE python video-dqn_b200/build.py --force
    python tools/role_profile.py

Every role (producer warps 0-3, MMA issuer, epilogue warps 0-3) reports cycles blocked in its waits
and its total loop time per CTA; printed as cycles per tile, averaged over CTAs."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from video_dqn_b200 import ops, _lib

lib = _lib.load()
fn = lib.vdqn_debug_role_profile
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
fn.restype = ctypes.c_int
ROLES = ["prod0", "prod1", "prod2", "prod3", "mma"] + [f"epi{i}" for i in range(8)]
WAITS = {"prod": ("empty", "cp.async/none"), "mma": ("tmem_empty", "full"), "epi": ("tmem_full", "inputs+store_read")}


def read(ncta):
    buf = np.zeros(160 * 16 * 4, dtype=np.uint64)
    rc = fn(buf.ctypes.data, buf.size)
    assert rc == 0, f"library not built with -DVDQN_ROLE_PROFILE (rc={rc})"
    return buf.reshape(160, 16, 4)[:ncta].astype(np.float64)


def run(name, x, w, pads, **kw):
    out = None
    lib.vdqn_debug_role_profile  # counters of roles that do not run in this variant keep stale values
    for _ in range(3):
        out = ops.conv_gemm(x, w, 1, pads[0], pads[1], out=out, algo=2, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv_gemm(x, w, 1, pads[0], pads[1], out=out, algo=2, **kw)
    e1.record()
    torch.cuda.synchronize()
    prof = read(ops.num_sms())
    print(f"== {name}: {e0.elapsed_time(e1) * 1e3:.1f} us")
    for r, rn in enumerate(ROLES):
        p = prof[:, r]
        tiles = p[:, 3].mean()
        if tiles == 0:
            continue
        wa, wb = WAITS[rn[:4] if rn.startswith("prod") else rn[:3]]
        tot = p[:, 2].mean()
        print(f"  {rn:6s} tiles/cta {tiles:6.1f}  loop {tot / tiles:8.0f} cyc/tile   wait[{wa}] {p[:, 0].mean() / tiles:8.0f}"
              f"   wait[{wb}] {p[:, 1].mean() / tiles:8.0f}   busy {(tot - p[:, 0].mean() - p[:, 1].mean()) / tiles:8.0f}")


g = torch.Generator(device="cuda").manual_seed(0)
bf = torch.bfloat16
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
xs = torch.randn(N, 112, 112, 16, device="cuda", generator=g).to(bf)
ws = (torch.randn(64, 4, 4, 16, device="cuda", generator=g) / 16).to(bf)
sh = torch.randn(64, device="cuda", generator=g)
run("stem (packed 4x4x16 -> 64), shift + relu", xs, ws, (2, 1), shift=sh, relu=True)
x1 = torch.randn(N, 56, 56, 64, device="cuda", generator=g).to(bf)
w1 = (torch.randn(64, 3, 3, 64, device="cuda", generator=g) / 24).to(bf)
res = torch.randn(N, 56, 56, 64, device="cuda", generator=g).to(bf)
run("layer1 conv, shift + relu", x1, w1, (1, 1), shift=sh, relu=True)
run("layer1 conv, shift + residual + relu", x1, w1, (1, 1), shift=sh, residual=res, relu=True)
cs = torch.zeros(64, device="cuda")
run("layer1 dgrad, residual + mask + colsum", x1, w1, (1, 1), residual=res, mask_src=res, colsum=cs)
