"""Where does the epilogue of the layer-1 halo kernel spend its time?  Times each variant (plain / residual /
mask + colsum / residual + mask + colsum) normally, with the TMA store skipped (debug flag 8), with the
epilogue math skipped (flag 32) and with both, on the normal build."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from video_dqn_b200 import ops

g = torch.Generator(device="cuda").manual_seed(0)
bf = torch.bfloat16


def rn(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda", generator=g) * scale).to(bf)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


w1 = rn(64, 3, 3, 64, scale=1 / 24)
sh = torch.randn(64, device="cuda", generator=g)
cs = torch.zeros(64, device="cuda")
ONLY = sys.argv[1] if len(sys.argv) > 1 else None      # e.g. `dgrad_res_mask`: that variant alone, flags 0 (ncu captures)
for N, variants in ((768, ("fwd", "fwd_res")), (256, ("dgrad_mask", "dgrad_res_mask", "plain"))):
    if ONLY is not None:
        variants = tuple(v for v in variants if v == ONLY)
        if not variants:
            continue
    # three rotating input sets: the working set stays larger than L2
    xs = [rn(N, 56, 56, 64) for _ in range(3)]
    rs = [rn(N, 56, 56, 64) for _ in range(3)]
    ms = [rn(N, 56, 56, 64) for _ in range(3)]
    out = torch.empty(N, 56, 56, 64, device="cuda", dtype=bf)
    for v in variants:
        kw = {"fwd": dict(shift=sh, relu=True), "fwd_res": dict(shift=sh, relu=True, residual=True),
              "dgrad_mask": dict(mask_src=True, colsum=cs), "dgrad_res_mask": dict(residual=True, mask_src=True, colsum=cs),
              "plain": {}}[v]
        line = []
        for flags in ((0,) if ONLY is not None else (0, 8, 32, 40)):
            i = [0]

            def fn():
                k = dict(kw)
                j = i[0] % 3
                i[0] += 1
                if k.get("residual"): k["residual"] = rs[j]
                if k.get("mask_src"): k["mask_src"] = ms[j]
                ops.conv_gemm(xs[j], w1, 1, 1, 1, out=out, algo=2, debug_flags=flags, **k)
            line.append(f"flags {flags:2d}: {timed(fn):6.1f} us")
        print(f"N={N} {v:15s} " + "   ".join(line), flush=True)
    del xs, rs, ms, out
