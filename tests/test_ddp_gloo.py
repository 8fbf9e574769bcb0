"""World-size-2 test of the data-parallel gradient exchange on CPU (gloo): the bucketed
all-reduce launched stage by stage from the backward pass covers the flat gradient arena exactly
once, in reverse parameter order, and yields the sum over ranks (Adam's grad_scale = 1/world turns
it into the global-batch mean).  The CUDA path uses the same class with NCCL."""
import os
import sys
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _fake_learner():
    from video_dqn_b200 import engine as E
    from video_dqn_b200.optim import FlatArena
    from video_dqn_b200.qnet import HabitatDQNMultiAction, grad_param_names
    m = HabitatDQNMultiAction(3, 5, extra_capacity=True, panorama=False)
    names = grad_param_names()
    mpd = dict(m.named_parameters())
    arena = FlatArena([mpd[n].shape for n in names], "cpu")
    G = dict(zip(names, arena.views()))
    plan = E.make_plan(3)
    lr = types.SimpleNamespace(opt=types.SimpleNamespace(grad_arena=arena.flat), model=m, G=G, plan=plan)
    return lr, names, arena


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from video_dqn_b200.ddp import GradSync
    lr, names, arena = _fake_learner()
    g = torch.Generator().manual_seed(100 + rank)
    arena.flat.copy_(torch.randn(arena.flat.shape, generator=g))
    local = arena.flat.clone()
    sync = GradSync(lr, bucket_bytes=8 << 20)
    stages = ["head"] + [b.conv1.name for b in reversed(lr.plan.blocks)] + ["stem"]
    for s in stages:
        sync.on_stage(s)
    sync.finish()
    # expected: sum over ranks
    other = torch.randn(arena.flat.shape, generator=torch.Generator().manual_seed(100 + (1 - rank)))
    ok_sum = torch.allclose(arena.flat, local + other, rtol=1e-6, atol=1e-6)
    spans = sync.launched
    covered = sorted(spans)
    contiguous = covered[0][0] == 0 and covered[-1][1] == arena.flat.numel() and \
        all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    reverse = all(a[0] >= b[1] for a, b in zip(spans, spans[1:]))
    q.put((rank, ok_sum, contiguous, reverse, len(spans)))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
    for rank, ok_sum, contiguous, reverse, n in res:
        assert ok_sum, f"rank {rank}: all-reduce result is not the sum over ranks"
        assert contiguous, f"rank {rank}: buckets do not tile the arena exactly once"
        assert reverse, f"rank {rank}: buckets not launched in reverse parameter order"
        assert 3 <= n <= 12, n
