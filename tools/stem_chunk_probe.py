"""Does the stem region (pack -> stem conv -> max-pool) run faster when it is walked in image chunks
small enough for the 1.6 MB/image pre-pool activation to stay in L2 (one reused scratch buffer)?
Times each chunk size as a replayed CUDA graph over n = 768 frames (the 3B pass at B = 256)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import qstep
from video_dqn_b200 import engine, ops

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
plan = engine.make_plan(3, 5)
sd = {k: v.to(dev) for k, v in qstep.init_state(seed=4, randomize_bn=True).items()}
W = engine.PreparedWeights(plan, dev)
W.prepare(sd)
frames = torch.randint(0, 256, (n, 224, 224, 3), dtype=torch.uint8, device=dev)
bf16 = torch.bfloat16
p_ref = None
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(chunk, with_idx, what="all"):
    xp = torch.empty(chunk if what != "nopack" else n, 112, 112, 16, device=dev, dtype=bf16)
    s = torch.empty(chunk, 112, 112, 64, device=dev, dtype=bf16)
    p = torch.empty(n, 56, 56, 64, device=dev, dtype=bf16)
    idx = torch.empty(n, 56, 56, 64, device=dev, dtype=torch.uint8) if with_idx else None

    def body():
        for c0 in range(0, n, chunk):
            c1 = min(n, c0 + chunk)
            m = c1 - c0
            ops.stem_pack(frames[c0:c1], xp[:m])
            engine._conv(W, plan.stem, xp[:m], s[:m], relu=True)
            ops.maxpool_fwd(s[:m], p[c0:c1], None if idx is None else idx[c0:c1])

    body()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    ts = []
    for _ in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], p


for with_idx in (False, True):
    for chunk in (n, 256, 128, 96, 64, 48, 32, 24, 16, 8):
        if chunk > n:
            continue
        ms, p = run(chunk, with_idx)
        if p_ref is None:
            p_ref = p.clone()
        same = bool(torch.equal(p, p_ref))
        print(f"idx={int(with_idx)} chunk={chunk:4d}  {ms * 1e3:8.1f} us  ({ms * 1e3 / n:6.3f} us/frame)  identical={same}",
              flush=True)
