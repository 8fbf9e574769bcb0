"""NVLink gradient exchange (csrc/nvl_allreduce.cu) against NCCL on N GPUs: correctness, bit-identity across
ranks, time for the 49.7 MB arena.   torchrun --nproc-per-node N tools/ddp_probe.py"""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_dqn_b200 import _lib as L  # noqa: E402
from video_dqn_b200.ddp import NvlGradSync  # noqa: E402


class _FakeLearner:
    class _Opt:
        pass

    def __init__(self, arena):
        self.opt = self._Opt()
        self.opt.grad_arena = arena


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 12426416
    alloc = NvlGradSync.allocator()
    buf = alloc(n, dev)
    arena = buf[:n]
    out = {}
    for use_mc in (True, False):
        sync = NvlGradSync(_FakeLearner(arena), alloc, use_multicast=use_mc)
        tag = "multimem" if sync.multicast else "p2p"
        if use_mc and not sync.multicast:
            if rank == 0:
                print("no multicast support on this fabric: multimem variant skipped")
            continue
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        src = torch.randn(n, device=dev, generator=g)
        ref = src.clone()
        dist.all_reduce(ref)
        ok = True
        for it in range(3):
            arena.copy_(src)
            buf[n:].zero_()
            torch.cuda.synchronize(); dist.barrier()
            sync.finish()
            torch.cuda.synchronize()
            err = ((arena - ref).abs().max() / ref.abs().max()).item()
            same = arena.clone()
            dist.broadcast(same, 0)
            ident = bool(torch.equal(same, arena))
            ok &= err < 1e-6 and ident
            if rank == 0:
                print(f"{tag} it {it}: max rel err vs NCCL {err:.2e}, bit-identical to rank 0: {ident}")
        # timing: the exchange alone, inputs already in place (the sum is re-summed: values grow, timing unaffected)
        arena.copy_(src * 1e-3)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            arena.mul_(1.0 / world)
            sync.finish()
        e1.record(); torch.cuda.synchronize()
        t_both = e0.elapsed_time(e1) / 20
        e0.record()
        for _ in range(20):
            arena.mul_(1.0 / world)
        e1.record(); torch.cuda.synchronize()
        t_mul = e0.elapsed_time(e1) / 20
        t = torch.tensor([t_both - t_mul], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[tag] = (ok, t.item())
        if rank == 0:
            print(f"{tag}: {'OK' if ok else 'FAILED'}; exchange of {n * 4 / 1e6:.1f} MB on {world} GPUs: {t.item() * 1e3:.1f} us "
                  f"(max over ranks)")
    arena.zero_()
    auto = NvlGradSync(_FakeLearner(arena), alloc)
    if rank == 0:
        print("autotune:", auto.tuning, "-> multicast" if auto.multicast else "-> p2p")
    x = arena.clone()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        dist.all_reduce(x)
    e0.record()
    for _ in range(20):
        dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"NCCL all_reduce of the same arena, alone on the GPU: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
    dist.barrier()
    dist.destroy_process_group()
    return 0 if all(v[0] for v in out.values()) else 1


if __name__ == "__main__":
    sys.exit(main())
