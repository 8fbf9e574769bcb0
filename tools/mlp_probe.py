"""mlp_gemm.cu flavour by flavour against torch fp64 (run on the GPU box)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_dqn_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
bf16 = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)


def split(x, ld=None, perm=None):
    ld = ld or x.shape[1]
    hi = torch.zeros(x.shape[0], ld, device=dev, dtype=bf16); lo = torch.zeros_like(hi)
    ops.split_bf16(x.contiguous(), hi, lo, perm=perm)
    return hi, lo


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-300)).item()


def main():
    ok = True
    for M in (24, 256, 768):
        x = torch.randn(M, 1600, device=dev, generator=g).to(bf16)          # head output: bf16 exact
        W = torch.randn(512, 1600, device=dev, generator=g) / 40
        b = torch.randn(512, device=dev, generator=g)
        wh, wl = split(W)
        zh, zl = torch.empty(M, 512, device=dev, dtype=bf16), torch.empty(M, 512, device=dev, dtype=bf16)
        zf = torch.empty(M, 512, device=dev)
        ops.mlp_gemm([(x, wh, 1600), (x, wl, 1600)], M, 512, bias=b, relu=True, out_hi=zh, out_lo=zl, out_f32=zf)
        ref = torch.relu(x.double() @ W.double().t() + b.double())
        e1, e2 = rel(zf, ref), rel(zh.double() + zl.double(), ref)
        print(f"fwd KK M={M}: fp32 out rel {e1:.2e}, hi+lo rel {e2:.2e}")
        ok &= e1 < 2e-5 and e2 < 2e-5
    # dual network, split at 512
    M = 768
    x = torch.randn(M, 512, device=dev, generator=g)
    xh, xl = split(x)
    W1, W2 = (torch.randn(256, 512, device=dev, generator=g) / 20 for _ in range(2))
    b1, b2 = (torch.randn(256, device=dev, generator=g) for _ in range(2))
    (w1h, w1l), (w2h, w2l) = split(W1), split(W2)
    out = torch.empty(M, 256, device=dev)
    ops.mlp_gemm([(xh, w1h, 512), (xl, w1h, 512), (xh, w1l, 512)], M, 256, bias=b1, out_f32=out,
                 segments2=[(None, w2h, 0), (None, w2h, 0), (None, w2l, 0)], bias2=b2, split_m=512)
    ref = torch.cat([x[:512].double() @ W1.double().t() + b1.double(), x[512:].double() @ W2.double().t() + b2.double()])
    e = rel(out, ref); print(f"fwd dual: rel {e:.2e}"); ok &= e < 2e-5
    # N = 15 output (top.4), fp32 store with ld 15
    W4 = torch.randn(15, 256, device=dev, generator=g) / 16
    b4 = torch.randn(15, device=dev, generator=g)
    z = torch.randn(M, 256, device=dev, generator=g)
    zh, zl = split(z); w4h, w4l = split(W4)
    q = torch.empty(M, 15, device=dev)
    ops.mlp_gemm([(zh, w4h, 256), (zl, w4h, 256), (zh, w4l, 256)], M, 15, bias=b4, out_f32=q)
    e = rel(q, z.double() @ W4.double().t() + b4.double()); print(f"fwd N=15: rel {e:.2e}"); ok &= e < 2e-5
    for B in (8, 256):
        # data gradient: dz = (dq W4) * (z2 > 0), K = 15, B operand MN-major; bias-gradient column sums
        dq = torch.randn(B, 15, device=dev, generator=g) * 1e-3
        dqh, dql = split(dq, ld=16)
        z2 = torch.randn(B, 256, device=dev, generator=g)
        z2h, z2l = split(z2)
        dzh, dzl = torch.empty(B, 256, device=dev, dtype=bf16), torch.empty(B, 256, device=dev, dtype=bf16)
        cs = torch.zeros(256, device=dev)
        ops.mlp_gemm([(dqh, w4h, 15), (dql, w4h, 15), (dqh, w4l, 15)], B, 256, b_mn=True, mask_bf16=z2h, out_hi=dzh,
                     out_lo=dzl, colsum=cs)
        ref = (dq.double() @ W4.double()) * (z2 > 0)
        e1, e2 = rel(dzh.double() + dzl.double(), ref), rel(cs, ref.sum(0))
        print(f"dgrad K-MN B={B}: rel {e1:.2e}, colsum rel {e2:.2e}"); ok &= e1 < 2e-5 and e2 < 1e-4
        # weight gradient: dW4 = dq^T z2 (both MN-major), M = 15
        dw = torch.zeros(15, 256, device=dev)
        ops.mlp_gemm([(dqh, z2h, B), (dql, z2h, B), (dqh, z2l, B)], 15, 256, a_mn=True, b_mn=True, out_f32=dw)
        e = rel(dw, dq.double().t() @ z2.double()); print(f"wgrad MN-MN M=15 B={B}: rel {e:.2e}"); ok &= e < 2e-5
        # d top.0.weight with the un-permuting store, and dh (bf16 out, mask, colsum mod 64)
        dz1 = torch.randn(B, 512, device=dev, generator=g) * 1e-3
        d1h, d1l = split(dz1)
        h = torch.randn(B, 25, 64, device=dev, generator=g).relu().to(bf16)       # NHWC head output
        hA = h.view(B, 1600)
        dW0 = torch.zeros(512, 1600, device=dev)
        ops.mlp_gemm([(d1h, hA, B), (d1l, hA, B)], 512, 1600, a_mn=True, b_mn=True, out_f32=dW0, perm=(64, 25))
        flat = h.permute(0, 2, 1).reshape(B, 1600).double()                          # reference Flatten order c*25 + p
        e = rel(dW0, dz1.double().t() @ flat); print(f"wgrad top.0 perm B={B}: rel {e:.2e}"); ok &= e < 2e-5
        W0 = torch.randn(512, 1600, device=dev, generator=g) / 40                     # reference column order
        w0h, w0l = split(W0, perm=(64, 25))
        dh = torch.empty(B, 1600, device=dev, dtype=bf16)
        cs = torch.zeros(64, device=dev)
        ops.mlp_gemm([(d1h, w0h, 512), (d1l, w0h, 512), (d1h, w0l, 512)], B, 1600, b_mn=True, mask_bf16=hA, out_bf16=dh,
                     colsum=cs, colsum_mod=64)
        dflat = dz1.double() @ W0.double()                                           # [B, c*25+p]
        ref = dflat.view(B, 64, 25).permute(0, 2, 1) * (h > 0)                        # NHWC
        e1 = rel(dh.view(B, 25, 64), ref)
        e2 = rel(cs, dh.view(B, 25, 64).double().sum((0, 1)))
        print(f"dh B={B}: rel {e1:.2e} (bf16 out), colsum rel {e2:.2e}"); ok &= e1 < 4e-3 and e2 < 1e-4
    print("ALL OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
