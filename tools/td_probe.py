"""TD epilogue timing probe: eager launches vs CUDA-graph replays (is the host the bound?), two batch
sizes, and a plain device copy of the same byte count for scale."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_dqn_b200 import ops  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


def main():
    dev = torch.device("cuda:0")
    out = {"env_bulk": os.environ.get("VDQN_TD_BULK", "1")}
    for lb in (20, 22):
        nb = 1 << lb
        q = [torch.randn(nb, 5, 3, device=dev) for _ in range(3)]
        act = torch.randint(0, 3, (nb,), device=dev)
        rw = (torch.rand(nb, 5, device=dev) < 0.1).long()
        dq, loss = torch.empty_like(q[0]), torch.zeros(1, device=dev)
        fn = lambda: ops.td_epilogue(q[0], q[1], q[2], act, rw, rw, dq=dq, loss=loss)  # noqa: E731
        eager = timeit(fn)
        g = torch.cuda.CUDAGraph()
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(10):
                fn()
        graph = timeit(g.replay, 5) / 10
        byts = nb * 328
        # a plain copy moving the same number of bytes (half read, half written)
        a = torch.empty(byts // 8, device=dev); b = torch.empty_like(a)
        cp = timeit(lambda: b.copy_(a))
        out[f"B=2^{lb}"] = {"eager_us": eager, "graph_us": graph, "copy_same_bytes_us": cp,
                            "eager_gbs": byts / eager / 1e3, "graph_gbs": byts / graph / 1e3, "copy_gbs": byts / cp / 1e3}
        del q, act, rw, dq, a, b
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
