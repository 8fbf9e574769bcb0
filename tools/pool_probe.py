"""Fused stem + max-pool kernel against the unfused pair (same stem kernel, then maxpool_fwd): pooled tensor and
arg-max slots must be identical; timing of both at 768 frames."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_dqn_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf16 = torch.bfloat16


def run(N, split, alias, nidx, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    nx = N - (alias[1] if alias else 0) if alias else N
    xp = torch.randn(N, 112, 112, 16, device=dev, generator=g).to(bf16)
    w = (torch.randn(64, 4, 4, 16, device=dev, generator=g) / 12).to(bf16)
    w2 = (torch.randn(64, 4, 4, 16, device=dev, generator=g) / 12).to(bf16)
    sh, sh2 = torch.randn(64, device=dev, generator=g) * 0.3, torch.randn(64, device=dev, generator=g) * 0.3
    kw = dict(w2=w2, shift2=sh2, split_n=split) if split else {}
    if alias:
        kw["x_alias"] = alias
    s = torch.empty(N, 112, 112, 64, device=dev, dtype=bf16)
    ops.conv_gemm(xp, w, 1, 2, 1, shift=sh, relu=True, out=s, **kw)
    p0 = torch.empty(N, 56, 56, 64, device=dev, dtype=bf16)
    i0 = torch.empty(N, 56, 56, 64, device=dev, dtype=torch.uint8)
    ops.maxpool_fwd(s, p0, i0)
    p1 = torch.full((N, 56, 56, 64), -7.0, device=dev, dtype=bf16)
    i1 = torch.full((max(nidx, 1), 56, 56, 64), 99, device=dev, dtype=torch.uint8)
    ops.conv_gemm(xp, w, 1, 2, 1, shift=sh, relu=True, pool_out=p1, pool_idx=i1, pool_idx_images=nidx, **kw)
    torch.cuda.synchronize()
    okp = torch.equal(p0, p1)
    oki = torch.equal(i0[:nidx], i1[:nidx]) if nidx else True
    untouched = bool((i1[nidx:] == 99).all()) if nidx < i1.shape[0] else True
    print(f"N={N} split={split} alias={alias} idx_images={nidx}: pooled identical {okp}, slots identical {oki}"
          + ("" if okp else f" (max diff {(p0.float() - p1.float()).abs().max().item():.3e}, "
                            f"{(p0 != p1).float().mean().item():.4f} differ)"))
    return okp and oki and untouched, (xp, w, sh, kw, s, p0, i0, p1, i1)


def main():
    ok = True
    ok &= run(1, 0, None, 1)[0]
    ok &= run(3, 0, None, 2)[0]
    ok &= run(5, 2, None, 0)[0]
    ok &= run(6, 4, (4, 2), 2)[0]
    good, (xp, w, sh, kw, s, p0, i0, p1, i1) = run(768, 512, (512, 256), 256)
    ok &= good

    def t(fn, n=10):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n * 1e3

    def unfused():
        ops.conv_gemm(xp, w, 1, 2, 1, shift=sh, relu=True, out=s, **kw)
        ops.maxpool_fwd(s[:256], p0[:256], i0[:256])
        ops.maxpool_fwd(s[256:], p0[256:], None)
    fused = lambda: ops.conv_gemm(xp, w, 1, 2, 1, shift=sh, relu=True, pool_out=p1, pool_idx=i1, pool_idx_images=256, **kw)  # noqa: E731
    print(f"768 frames: stem + pooling kernels {t(unfused):.1f} us, fused {t(fused):.1f} us")
    for name, fl in (("no pooling work", 8), ("pooling without its global stores", 64), ("no MMAs", 16), ("no drain stores", 32), ("no pooling, no drain stores", 40),
                     ("nothing but the producer", 56)):
        f = lambda: ops.conv_gemm(xp, w, 1, 2, 1, shift=sh, relu=True, pool_out=p1, pool_idx=i1, pool_idx_images=256,  # noqa: E731
                                  debug_flags=fl, **kw)
        print(f"   fused, {name}: {t(f):.1f} us")
    f = lambda: ops.conv_gemm(xp, w, 1, 2, 1, shift=sh, relu=True, out=s, debug_flags=16, **kw)  # noqa: E731
    print(f"   unfused stem, no MMAs: {t(f):.1f} us")
    f = lambda: ops.conv_gemm(xp, w, 1, 2, 1, shift=sh, relu=True, out=s, debug_flags=8, **kw)  # noqa: E731
    print(f"   unfused stem, no stores: {t(f):.1f} us")
    f = lambda: ops.conv_gemm(xp, w, 1, 2, 1, shift=sh, relu=True, out=s, **kw)  # noqa: E731
    print(f"   unfused stem alone: {t(f):.1f} us")
    print("ALL OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
