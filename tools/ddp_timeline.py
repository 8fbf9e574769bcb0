"""Where does the early gradient exchange run?  One eager step per rank with CUDA events on both streams.
torchrun --nproc-per-node N tools/ddp_timeline.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    model, target, lr = bench.make_dp_learner(dev, world, 256, False, "nvl")
    sync = lr.grad_sync
    batch = [t.to(dev) for t in bench.synthetic_quads(256, seed=3 + rank, pinned=False)]
    for _ in range(3):
        lr.step(batch)
    torch.cuda.synchronize(); dist.barrier()
    ev = {}
    orig_on, orig_fin = sync.on_stage, sync.finish

    def on_stage(stage):
        if sync.wants(stage):
            ev["main@early_stage"] = torch.cuda.Event(enable_timing=True); ev["main@early_stage"].record()
        orig_on(stage)
        if sync.wants(stage):
            with torch.cuda.stream(sync.side):
                ev["side:early_done"] = torch.cuda.Event(enable_timing=True); ev["side:early_done"].record()

    def finish():
        ev["main@backward_end"] = torch.cuda.Event(enable_timing=True); ev["main@backward_end"].record()
        orig_fin()
        ev["main@exchange_end"] = torch.cuda.Event(enable_timing=True); ev["main@exchange_end"].record()
    sync.on_stage, sync.finish = on_stage, finish
    t0 = torch.cuda.Event(enable_timing=True); t0.record()
    lr.step(batch)
    t1 = torch.cuda.Event(enable_timing=True); t1.record()
    torch.cuda.synchronize()
    out = {k: round(t0.elapsed_time(v), 3) for k, v in ev.items()}
    out["step_end"] = round(t0.elapsed_time(t1), 3)
    print(f"rank {rank} (eager step, ms from step start): {out}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
