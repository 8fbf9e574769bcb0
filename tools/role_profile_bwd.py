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
NF = 3 * N


def rn(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(bf)


print("#### halo_conv_kernel, layer 1")
w1 = rn(64, 3, 3, 64, scale=1 / 24)
sh = torch.randn(64, device="cuda", generator=g)
xf, rf = rn(NF, 56, 56, 64), rn(NF, 56, 56, 64)
run(f"layer1 forward, {NF} images, shift + relu", xf, w1, (1, 1), shift=sh, relu=True)
run(f"layer1 forward, {NF} images, shift + residual + relu", xf, w1, (1, 1), shift=sh, residual=rf, relu=True)
del xf, rf
x1, res, msk = rn(N, 56, 56, 64), rn(N, 56, 56, 64), rn(N, 56, 56, 64)
cs = torch.zeros(64, device="cuda")
run("layer1 dgrad (conv2): mask + colsum", x1, w1, (1, 1), mask_src=msk, colsum=cs)
run("layer1 dgrad (conv1): residual + mask + colsum", x1, w1, (1, 1), residual=res, mask_src=msk, colsum=cs)
run("layer1 dgrad, plain", x1, w1, (1, 1))

print("\n#### igemm_kernel, data gradients of layers 2-4 as the engine launches them")
for ch, hw, tn in ((128, 28, 0), (256, 14, 128), (512, 7, 128)):
    x = rn(N, hw, hw, ch)
    w = rn(ch, 3, 3, ch, scale=1 / (3 * ch ** 0.5))
    m, r = rn(N, hw, hw, ch), rn(N, hw, hw, ch)
    c = torch.zeros(ch, device="cuda")
    run(f"dgrad {ch}ch {hw}x{hw} (conv2): mask + colsum, tile_n={tn}", x, w, (1, 1), igemm=True, algo=0, mask_src=m, colsum=c, tile_n=tn)
    run(f"dgrad {ch}ch {hw}x{hw} (conv1): residual + mask + colsum, tile_n={tn}", x, w, (1, 1), igemm=True, algo=0,
        residual=r, mask_src=m, colsum=c, tile_n=tn)
    run(f"dgrad {ch}ch {hw}x{hw} plain, tile_n={tn}", x, w, (1, 1), igemm=True, algo=0, tile_n=tn)
