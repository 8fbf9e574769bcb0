"""Achieved HBM bandwidth of the train-mode BatchNorm kernels (csrc/batchnorm.cu) at the sizes of the
B = 256 step, timed with CUDA events on the launching stream (10 warm-up + 50 timed launches each, the
tensors of consecutive launches rotate over > 126 MB so L2 does not serve them).  Run under gpurun:

    python tools/bn_bandwidth.py > gpurun_out/bn_bandwidth.json

Algorithmic bytes per element: bn_stats 2 (read x), bn_apply 4 (+2 with a residual), bn_bwd_reduce 4,
bn_bwd_apply 6.  Peak = MEASURED_PEAKS.json hbm_gbs when present, else 6556.5 GB/s.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_dqn_b200 import _lib as L  # noqa: E402
from video_dqn_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    peak = 6556.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    lib = L.load()
    out = {"peak_gbs": peak, "kernels": []}
    for (N, HW, C) in ((256, 112, 64), (256, 56, 64), (256, 28, 128), (256, 7, 512)):
        M = N * HW * HW
        nbuf = max(2, int(300e6 // (M * C * 2)) + 1)
        xs = [torch.randn(M, C, device=dev).to(torch.bfloat16) for _ in range(nbuf)]
        ys = [torch.empty_like(x) for x in xs[:2]]
        st = ops.BnBatchStats(C, dev)
        g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        dg, db = torch.empty(C, device=dev), torch.empty(C, device=dev)
        ops.bn_train_fwd(xs[0], st, g, b, rm, rv, None, ys[0])
        sp = L.stream_ptr()
        cases = {
            "bn_stats": (2, lambda i: lib.vdqn_bn_stats(xs[i % nbuf].data_ptr(), st.sums.data_ptr(), M, C, sp)),
            "bn_apply": (4, lambda i: lib.vdqn_bn_apply(xs[i % nbuf].data_ptr(), st.scale.data_ptr(), st.shift.data_ptr(),
                                                        None, 1, ys[i % 2].data_ptr(), M, C, sp)),
            "bn_apply+residual": (6, lambda i: lib.vdqn_bn_apply(xs[i % nbuf].data_ptr(), st.scale.data_ptr(),
                                                                 st.shift.data_ptr(), xs[(i + 1) % nbuf].data_ptr(), 1,
                                                                 ys[i % 2].data_ptr(), M, C, sp)),
            "bn_bwd_reduce": (4, lambda i: lib.vdqn_bn_bwd_reduce(xs[i % nbuf].data_ptr(), xs[(i + 1) % nbuf].data_ptr(),
                                                                  st.mean.data_ptr(), st.rstd.data_ptr(),
                                                                  st.bsums.data_ptr(), M, C, sp)),
            "bn_bwd_apply": (6, lambda i: lib.vdqn_bn_bwd_apply(xs[i % nbuf].data_ptr(), xs[(i + 1) % nbuf].data_ptr(),
                                                                st.mean.data_ptr(), st.rstd.data_ptr(), g.data_ptr(),
                                                                st.bsums.data_ptr(), dg.data_ptr(), db.data_ptr(),
                                                                ys[i % 2].data_ptr(), M, C, sp)),
        }
        for name, (bpe, fn) in cases.items():
            for i in range(10):
                L.check(fn(i), name)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(50):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 50
            gbs = bpe * M * C / (ms * 1e-3) / 1e9
            out["kernels"].append({"kernel": name, "M": M, "C": C, "us": round(ms * 1e3, 2), "bytes": bpe * M * C,
                                   "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak, 3)})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
