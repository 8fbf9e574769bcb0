"""Which warp role bounds the halo convolution kernels?  Needs the instrumented library:

    VDQN_NVCC_FLAGS=-DVDQN_ROLE_PROFILE python video_dqn_b200/build.py --force
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
fn_igemm = lib.vdqn_debug_role_profile_igemm
fn_igemm.argtypes = [ctypes.c_void_p, ctypes.c_int]
fn_igemm.restype = ctypes.c_int
ROLES = ["prod0", "prod1", "prod2", "prod3", "mma"] + [f"epi{i}" for i in range(8)]
WAITS = {"prod": ("empty", "cp.async/none"), "mma": ("tmem_empty", "full"), "epi": ("tmem_full", "inputs+store_read")}


def read(ncta, igemm=False):
    buf = np.zeros(160 * 16 * 4, dtype=np.uint64)
    rc = (fn_igemm if igemm else fn)(buf.ctypes.data, buf.size)
    assert rc == 0, f"library not built with -DVDQN_ROLE_PROFILE (rc={rc})"
    return buf.reshape(160, 16, 4)[:ncta].astype(np.float64)


def run(name, x, w, pads, igemm=False, stride=1, algo=2, out=None, **kw):
    # counters of roles that do not run in a variant (idle producer warps, the peer CTA's MMA warp) keep
    # stale values from earlier launches: read only the rows that make sense for the kernel at hand
    for _ in range(3):
        out = ops.conv_gemm(x, w, stride, pads[0], pads[1], out=out, algo=algo, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.conv_gemm(x, w, stride, pads[0], pads[1], out=out, algo=algo, **kw)
    e1.record()
    torch.cuda.synchronize()
    prof = read(ops.num_sms(), igemm)
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

# ---- im2col kernel (igemm): the small-K launches of the backward pass and a layer-2 forward conv
print("\n#### igemm_kernel (producer = prod0, MMA issuer = mma, epilogue warps epi0-7)")
x2 = torch.randn(N, 28, 28, 128, device="cuda", generator=g).to(bf)
w2 = (torch.randn(128, 3, 3, 128, device="cuda", generator=g) / 34).to(bf)
sh2 = torch.randn(128, device="cuda", generator=g)
run("layer2 conv 3x3 128->128, shift + relu (CTA pairs)", x2, w2, (1, 1), igemm=True, algo=0, shift=sh2, relu=True)
# parity class (1,1) of the layer2.0.conv1 data gradient: 2x2 taps, dy [N,28,28,128] -> dx[:, 1::2, 1::2] of [N,56,56,64]
wp = (torch.randn(64, 2, 2, 128, device="cuda", generator=g) / 23).to(bf)
res = torch.randn(N, 56, 56, 64, device="cuda", generator=g).to(bf)
dx = torch.zeros(N, 56, 56, 64, device="cuda", dtype=bf)
cs64 = torch.zeros(64, device="cuda")
run("parity class (1,1) data gradient, 4 taps, residual + mask + colsum, scatter", x2, wp, (0, 1), igemm=True, algo=0,
    out=dx, residual=res, mask_src=res, colsum=cs64, out_scatter=2, scatter_off=(1, 1), scatter_inputs=True)
wp1 = (torch.randn(64, 1, 1, 128, device="cuda", generator=g) / 11).to(bf)
run("parity class (0,0) data gradient, 1 tap", x2, wp1, (0, 0), igemm=True, algo=0,
    out=dx, residual=res, mask_src=res, colsum=cs64, out_scatter=2, scatter_off=(0, 0), scatter_inputs=True)
x3 = torch.randn(N, 14, 14, 256, device="cuda", generator=g).to(bf)
w3 = (torch.randn(256, 3, 3, 256, device="cuda", generator=g) / 48).to(bf)
cs256 = torch.zeros(256, device="cuda")
m3 = torch.randn(N, 14, 14, 256, device="cuda", generator=g).to(bf)
run("layer3 data gradient 3x3 256->256, mask + colsum (BN = 256, direct epilogue)", x3, w3, (1, 1), igemm=True, algo=0,
    mask_src=m3, colsum=cs256)
