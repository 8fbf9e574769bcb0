"""TD epilogue at B = 2^20 (the size at which the kernel is HBM-measurable): staged variant
(16-byte aligned tensors) against the direct one (forced by tensors that start 4 bytes off alignment), CUDA
events over 20 launches with preallocated outputs; also checks that the two produce identical dQ / y / best.
Algorithmic bytes per sample: 3 x 60 (Q) + 8 (act) + 40 + 40 (rew, term) + 60 (dQ) = 328.

    python tools/td_bandwidth.py > gpurun_out/td_bandwidth.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_dqn_b200 import ops  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    peak = 6556.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    nb = 1 << 20
    g = torch.Generator(device=dev).manual_seed(0)
    act = torch.randint(0, 3, (nb,), device=dev, generator=g)
    rw = (torch.rand(nb, 5, device=dev, generator=g) < 0.1).long()
    out = {"peak_gbs": peak, "bytes_per_sample": 328, "B": nb}
    res = {}
    # aligned tensors take the streaming (bulk-copy) kernel, or with VDQN_TD_BULK=0 the staged per-thread one
    aligned = "staged" if os.environ.get("VDQN_TD_BULK", "1") == "0" else "bulk"
    for name, off in ((aligned, 0), ("direct", 1)):
        bufs = []
        for k in range(4):                     # same values in both variants, only the start address differs
            data = torch.randn(nb * 15, device=dev, generator=torch.Generator(device=dev).manual_seed(7 + k))
            buf = torch.empty(nb * 15 + 4, device=dev)
            buf[off:off + nb * 15].copy_(data)
            bufs.append(buf)
            del data
        q = [b[off:off + nb * 15].view(nb, 5, 3) for b in bufs[:3]]
        dq = bufs[3][off:off + nb * 15].view(nb, 5, 3)
        loss = torch.zeros(1, device=dev)
        best = torch.empty(nb, 5, device=dev, dtype=torch.int64)
        y = torch.empty(nb, 5, device=dev)
        ops.td_epilogue(q[0], q[1], q[2], act, rw, rw, dq=dq, loss=loss, best=best, y=y, want_aux=True)
        torch.cuda.synchronize()
        res[name] = (dq.clone(), best.clone(), y.clone(), loss.item())
        for _ in range(3):
            ops.td_epilogue(q[0], q[1], q[2], act, rw, rw, dq=dq, loss=loss)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            ops.td_epilogue(q[0], q[1], q[2], act, rw, rw, dq=dq, loss=loss)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        gbs = nb * 328 / (ms * 1e-3) / 1e9
        out[name] = {"us": round(ms * 1e3, 2), "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak, 3)}
        del bufs, q, dq
    a, b = res[aligned], res["direct"]
    out["identical"] = {"dq": bool(torch.equal(a[0], b[0])), "best": bool(torch.equal(a[1], b[1])),
                        "y": bool(torch.equal(a[2], b[2])), "loss": [a[3], b[3]]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
