#!/usr/bin/env python
"""Benchmark of the Q-learning training step (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # ours (CUDA kernels via the C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on host cores

A step = one full iteration of the reference loop body (train_q_network.py:211-229) on one batch
of synthetic quadruplets: 3 forwards (online s, online s', target s'), Double-DQN TD loss, backward,
Adam, weight refresh.  Workload = BASELINE.json configs[1]: HabitatDQNMultiAction (extra_capacity,
1 frame, 3 actions), per-GPU batch 256, bf16 operands / fp32 accumulation, random-init weights.
`value` times the step with inputs resident in HBM; `e2e` goes through the public API from pinned
HOST buffers (async H2D of the batch + D2H of the loss inside the timed region).
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Q-learning train steps/sec & frames/sec"
BATCH_PER_GPU = 256
# algorithmic work of the reference network per frame (hook-counted)
CONV_FWD_FLOP = 2 * (1821888256 - 954112)          # all 21 convs (7x7x3 stem), per frame
STEM_FLOP = 2 * 112 * 112 * 64 * 147
MLP_FLOP = 2 * 954112
IGEMM_FLOP_PER_QUAD = 3 * CONV_FWD_FLOP + (CONV_FWD_FLOP - STEM_FLOP)   # 3 fwd + dgrad (no stem dgrad)
WGRAD_FLOP_PER_QUAD = CONV_FWD_FLOP
STEP_FLOP_PER_QUAD = 17982854656                   # SURVEY.md 8d, reference semantics
N_PARAMS = 12426383


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region.  A timed region of K graph replays lasts a
    fraction of a second, too short for `nvidia-smi -lms`, so a thread polls NVML (nvidia-ml-py) every
    5 ms with host time stamps; `stop(t0, t1)` keeps the samples taken between the two.  Falls back to
    the nvidia-smi loop of /opt/skills/guides/B200_PROFILING.md when NVML cannot be opened."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.p, self.thread, self.samples, self._stop = index, None, None, [], False

    def _nvml_loop(self, nv, h):
        R = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
             "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        while not self._stop:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.samples.append((time.perf_counter(), float(mhz), [k for k, b in R.items() if bits & b]))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: map through CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.index])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self, t0=None, t1=None):
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            inside = [x for x in self.samples if t0 is None or t0 <= x[0] <= t1]
            use = inside if inside else self.samples
            sm = [x[1] for x in use]
            reasons = sorted({r for x in use for r in x[2]})
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.smax, "reasons": reasons,
                    "samples": len(inside), "samples_total": len(self.samples), "source": "nvml thread, 5 ms period, "
                    "samples between the first launch and the final synchronize of the timed region"}
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        load = [c for c in sm if c > 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(load) if load else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 100"}


def synthetic_quads(B, seed, pinned=True):
    """uint8 HWC frames + loader-typed labels (dataloaders/q_learning_real.py:75-98)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    mk = lambda *s, dt: torch.empty(*s, dtype=dt, pin_memory=pinned)  # noqa: E731
    before = mk(B, 224, 224, 3, dt=torch.uint8); after = mk(B, 224, 224, 3, dt=torch.uint8)
    before.random_(0, 256, generator=g); after.random_(0, 256, generator=g)
    act = mk(B, dt=torch.int64); act.random_(0, 3, generator=g)
    rew = mk(B, 5, dt=torch.int64); rew.copy_((torch.rand(B, 5, generator=g) < 0.1).long())
    term = mk(B, 5, dt=torch.int64); term.copy_(rew)
    valid = mk(B, 5, dt=torch.int64); valid.fill_(1)
    gt = torch.full((B,), float("nan"), dtype=torch.float64)
    return before, after, act, rew, term, gt, valid


# ----------------------------------------------------------------------------------------------
def cpu_reference_arm(steps, warmup, batch, threads):
    """The reference algorithm (oracle port, fp32, torch CPU kernels) on the host cores."""
    import torch
    from oracle import qstep
    torch.set_num_threads(threads)
    tr = qstep.OracleTrainer(qstep.init_state(seed=4))
    data = [qstep.synthetic_batch(batch, seed=1 + i) for i in range(2)]
    for i in range(warmup):
        tr.step(data[i % 2])
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        tr.step(data[i % 2])
        times.append(time.perf_counter() - t0)
    return statistics.median(times) if times else 0.0      # median of the timed steps


def make_config(world, B, graph):
    return {"workload": "HabitatDQNMultiAction (extra_capacity, 1 frame, 3 actions) Double-DQN "
                        "training step, synthetic quadruplets (BASELINE configs[1])",
            "batch_per_gpu": B, "global_batch": world * B, "frame": "3x224x224",
            "parallelism": f"dp{world}", "cuda_graph": graph,
            "l2": "inputs rotate over 3 device-resident batches (231 MB) and the step streams "
                  ">3 GB of activations: working set larger than the 126 MB L2",
            "random_init_weights": True}


def find_reference():
    """Directory holding the UNMODIFIED reference sources of the path (archs/HabitatDQNMultiAction.py and
    train_q_network.py): `baseline/_ref` (git-ignored copy made by __graft_entry__.build(); travels to the
    GPU box with the snapshot) or the container's /root/reference.  None: only the oracle port is left."""
    for d in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.exists(os.path.join(d, "archs", "HabitatDQNMultiAction.py")) and \
                os.path.exists(os.path.join(d, "train_q_network.py")):
            return d
    return None


def cpu_reference_real(ref_dir, steps, warmup, batch, threads):
    """The reference ITSELF on the host cores: its own `HabitatDQNMultiAction` (imported unmodified; only
    the ImageNet download is patched out), its own `process_batch` (lifted out of `run_train` with ast, it is
    a closure) and `torch.optim.Adam`, driven by the loop body of train_q_network.py:221-229."""
    import types
    import torch
    os.environ["VDQN_REFERENCE"] = ref_dir
    from oracle import make_goldens, qstep
    make_goldens.REF = ref_dir
    torch.set_num_threads(threads)
    RefNet, restore = make_goldens.import_reference_model()
    cfg = qstep.StepConfig()
    refcfg = types.SimpleNamespace(device="cpu", LINEAR=cfg.LINEAR, GAMMA=cfg.GAMMA, LOSS_CLIP=cfg.LOSS_CLIP,
                                   VALUE_LEARNING=False, REMOVE_BEFORE_REWARD=cfg.REMOVE_BEFORE_REWARD)
    model = RefNet(3, 5, extra_capacity=True, panorama=False)
    target = RefNet(3, 5, extra_capacity=True, panorama=False)
    restore()
    model.load_state_dict(qstep.init_state(seed=4), strict=True)
    target.load_state_dict(model.state_dict())
    target.eval()
    opt = torch.optim.Adam(model.parameters(), lr=cfg.LEARNING_RATE)
    process_batch = make_goldens.lift_process_batch(model, target, refcfg)
    data = [qstep.synthetic_batch(batch, seed=1 + i) for i in range(2)]

    def one(b):
        model.set_train()
        opt.zero_grad()
        loss = process_batch(b)
        loss.backward()
        opt.step()
        return loss.item()
    for i in range(warmup):
        one(data[i % 2])
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        one(data[i % 2])
        times.append(time.perf_counter() - t0)
    return statistics.median(times) if times else 0.0


def run_reference(a, out_stream=sys.stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = 8
    steps, warmup = max(1, a.steps), max(0, a.warmup)       # as given; a B = 8 step is ~0.1 s on 16 cores
    ref_dir = find_reference()
    kind, dt = "port", None
    if ref_dir is not None:
        try:
            dt = cpu_reference_real(ref_dir, steps, warmup, B, threads)
            kind = "reference"
        except Exception as exc:                            # noqa: BLE001  (report the port instead)
            print(f"reference arm: could not run the reference from {ref_dir}: {exc!r}", file=sys.stderr)
    if dt is None:
        dt = cpu_reference_arm(steps, warmup, B, threads)
    fps = 2 * B / dt
    what = ("the reference's own HabitatDQNMultiAction + process_batch + torch.optim.Adam"
            if kind == "reference" else "fp32 oracle port")
    sample = (f"median of {steps} steps of {B} quadruplets after {warmup} warm-up steps (of the {BATCH_PER_GPU}-quadruplet "
              f"workload), {what}, fp32, all host cores")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": a.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": dt * 1e3, "steps_per_sec": 1.0 / dt,
        "quadruplets_per_sec": B / dt, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(make_config(a.gpus, BATCH_PER_GPU, False), cuda_graph=None, parallelism="cpu",
                       sample=sample, threads=threads),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), file=out_stream, flush=True)


def _inverse_block(dev):
    """Inverse-dynamics model (BASELINE configs[4]): training step on frame pairs at the reference's batch
    (train_inverse_model.py:21,93-110) and the labelling forward (dataset/process_episodes_real.py:171-179)."""
    import torch
    from video_dqn_b200.inverse import InverseActionModule, InverseActionRunner, InverseModelTrainer
    nb = 128
    torch.manual_seed(7)
    isd = InverseActionModule().state_dict()    # random init in the reference module's key layout
    g = torch.Generator().manual_seed(11)
    hk = torch.randn(nb, 3, 224, 224, generator=g).pin_memory()
    hk1 = torch.randn(nb, 3, 224, 224, generator=g).pin_memory()
    hact = torch.randint(0, 3, (nb,), generator=g).pin_memory()
    dk, dk1, dact = hk.to(dev), hk1.to(dev), hact.to(dev)
    tr = InverseModelTrainer(isd, nb, lr=1e-4, device=dev)
    for _ in range(5):
        tr.step(dk, dk1, dact)
    torch.cuda.synchronize()
    iters = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        tr.step(dk, dk1, dact)
    e1.record(); torch.cuda.synchronize()
    train_ms = e0.elapsed_time(e1) / iters
    t0 = time.perf_counter()
    for _ in range(iters):
        lt = tr.step(hk, hk1, hact)             # H2D of the fp32 pair batch (154 MB) + step
    _ = lt.item()
    train_e2e_ms = (time.perf_counter() - t0) / iters * 1e3
    loss_f32 = float(lt.item())
    # the same from uint8 HWC frames (decoder output; normalisation fused into the first kernel): 39 MB H2D
    del tr
    tr = InverseModelTrainer(isd, nb, lr=1e-4, device=dev, frames_uint8=True)
    hu = torch.empty(nb, 224, 224, 3, dtype=torch.uint8, pin_memory=True).random_(0, 256)
    hu1 = torch.empty(nb, 224, 224, 3, dtype=torch.uint8, pin_memory=True).random_(0, 256)
    for _ in range(5):
        tr.step(hu, hu1, hact)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        lt = tr.step(hu, hu1, hact)
    _ = lt.item()
    train_e2e_u8_ms = (time.perf_counter() - t0) / iters * 1e3
    run = InverseActionRunner(isd, nb, dev)
    for _ in range(3):
        run.label(dk, dk1)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        run.label(dk, dk1)
    e1.record(); torch.cuda.synchronize()
    label_ms = e0.elapsed_time(e1) / iters
    # per pair: two trunk forwards + head forward (7.31 GFLOP) + head backward (2 x 29.0 MMAC x 2)
    flop_pair = 7.31e9 + 0.116e9
    return {"batch": nb, "train_step_ms": train_ms, "pairs_per_sec_device": nb / train_ms * 1e3,
               "train_step_e2e_ms": train_e2e_ms, "pairs_per_sec_e2e": nb / train_e2e_ms * 1e3,
               "train_step_e2e_uint8_ms": train_e2e_u8_ms, "pairs_per_sec_e2e_uint8": nb / train_e2e_u8_ms * 1e3,
               "train_tflops": flop_pair * nb / (train_ms * 1e-3) / 1e12,
               "label_ms": label_ms, "pairs_per_sec_label": nb / label_ms * 1e3,
               "frames": "fp32 NCHW pairs (the reference loader's output); *_uint8: uint8 HWC pairs",
               "loss": loss_f32}


def _basic_block(dev, B):
    """Training step of the `basic` architecture (train-mode BatchNorm) at the bench batch:
    CUDA-graph replay (and eager launches next to it), device-resident fp32 frames."""
    import torch
    from video_dqn_b200.learner import StepConfig
    from video_dqn_b200.learner_basic import BasicQLearner
    from video_dqn_b200.qnet import HabitatDQNMultiAction
    batches = [[t.to(dev) for t in synthetic_quads(B, seed=50 + i, pinned=False)] for i in range(2)]
    res = {}
    for mode, graph in (("eager", False), ("graph", True)):
        torch.manual_seed(4)
        nets = [HabitatDQNMultiAction(3, 5, extra_capacity=False, panorama=False).to(dev) for _ in range(2)]
        nets[1].load_state_dict(nets[0].state_dict())
        lr = BasicQLearner(nets[0], nets[1], StepConfig(), batch_size=B, use_graph=graph)
        for i in range(3):
            lr.step(batches[i % 2])
        torch.cuda.synchronize()
        iters = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            loss = lr.step(batches[i % 2])
        e1.record(); torch.cuda.synchronize()
        res[mode] = (e0.elapsed_time(e1) / iters, float(loss.item()))
        del lr, nets
        torch.cuda.empty_cache()
    ms = res["graph"][0]
    out = {"batch": B, "ms_per_step": ms, "frames_per_sec": 2 * B / ms * 1e3, "loss": res["graph"][1],
           "ms_per_step_eager": res["eager"][0],
           "mode": "train-mode BatchNorm (batch statistics), CUDA-graph replay (ms_per_step_eager: eager launches)"}
    del batches
    torch.cuda.empty_cache()
    return out


def make_dp_learner(dev, world, B, use_graph, kind):
    """One learner of the data-parallel job: `kind` = "nvl" (the gradient arena in NVLink symmetric memory, the
    exchange is this library's kernel) or "nccl" (bucketed NCCL all-reduce from inside the backward pass)."""
    import torch
    from video_dqn_b200.ddp import GradSync, NvlGradSync
    from video_dqn_b200.learner import QLearner, StepConfig
    from video_dqn_b200.qnet import HabitatDQNMultiAction
    torch.manual_seed(4)                      # SEED of configs/experiments/real_data/config.yml:11
    model = HabitatDQNMultiAction(3, 5, extra_capacity=True, panorama=False).to(dev)
    target = HabitatDQNMultiAction(3, 5, extra_capacity=True, panorama=False).to(dev)
    target.load_state_dict(model.state_dict())
    target.eval()
    alloc = NvlGradSync.allocator() if (world > 1 and kind == "nvl") else None
    learner = QLearner(model, target, StepConfig(), batch_size=B, frames_uint8=True, use_graph=use_graph,
                       world_size=world, grad_alloc=alloc)
    if world > 1:
        learner.grad_sync = NvlGradSync(learner, alloc) if alloc is not None else GradSync(learner)
    return model, target, learner


def dp_check(dev, rank, world, kind):
    """On-hardware data-parallel correctness (the reference has no DP, train_q_network.py:275):
    every rank takes ONE small step on its own 8 quadruplets with the gradient exchange of the timed run,
    then (a) the parameters of all ranks must be bit-identical and (b) the exchanged (averaged) gradient is
    compared with the gradient a single process computes on the concatenated global batch (the loss is a
    mean over B*5 entries, so the global-batch gradient is the average of the per-rank ones)."""
    import torch
    import torch.distributed as dist
    Bs = 8
    m, t, lr = make_dp_learner(dev, world, Bs, False, kind)
    lr.step([x.to(dev) for x in synthetic_quads(Bs, seed=500 + rank, pinned=False)])
    torch.cuda.synchronize()
    p = lr.opt.param_arena
    ref = p.clone()
    dist.broadcast(ref, 0)
    diff = (p - ref).abs().max().reshape(1)
    dist.all_reduce(diff, op=dist.ReduceOp.MAX)
    g_avg = (lr.opt.grad_arena / world).clone()
    loss_mean = lr.loss.clone()
    dist.all_reduce(loss_mean, op=dist.ReduceOp.SUM)
    out = {"batch_per_rank": Bs, "max_param_diff_across_ranks": float(diff.item()),
           "exchange": type(lr.grad_sync).__name__, "multicast": getattr(lr.grad_sync, "multicast", None),
           "exchange_us": getattr(lr.grad_sync, "tuning", None)}
    if hasattr(lr.grad_sync, "close"):
        lr.grad_sync.close()
    del lr, m, t
    if rank == 0:
        m, t, big = make_dp_learner(dev, 1, world * Bs, False, kind)
        parts = [synthetic_quads(Bs, seed=500 + r, pinned=False) for r in range(world)]
        big.step([torch.cat([pt[i] for pt in parts]).to(dev) for i in range(7)])
        torch.cuda.synchronize()
        g = big.opt.grad_arena
        out["avg_grad_rel_l2_vs_global_batch"] = float(((g_avg - g).norm() / g.norm()).item())
        out["loss_mean_over_ranks"] = float(loss_mean.item() / world)
        out["loss_global_batch"] = float(big.loss.item())
        del big, m, t
    torch.cuda.empty_cache()
    dist.barrier()
    return out


# ----------------------------------------------------------------------------------------------
def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its
    version banner on stdout when NCCL_DEBUG is set in the environment), so keep a private handle on
    the real stdout for the JSON line and point file descriptor 1 at stderr for everything else."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    out_stream = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--detail", action="store_true", help="also print per-shape conv timings to stderr")
    ap.add_argument("--no-inference", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        return run_reference(a, out_stream)
    a.warmup = max(a.warmup, 3)

    import torch
    import torch.distributed as dist
    from video_dqn_b200 import _lib, ops
    from video_dqn_b200.ddp import GradSync
    from video_dqn_b200.learner import QLearner, StepConfig
    from video_dqn_b200.qnet import HabitatDQNMultiAction
    from video_dqn_b200.staging import BatchStager

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    _lib.check(lib.vdqn_init(local), "init")
    B = a.batch

    # gradient exchange: VDQN_DDP=nvl (default: this library's NVLink kernel) | nccl
    dp_kind = os.environ.get("VDQN_DDP", "nvl") if world > 1 else None
    dpc = None
    if world > 1:
        try:
            dpc = dp_check(dev, rank, world, dp_kind)
        except Exception as exc:                      # symmetric memory unavailable on this fabric: NCCL buckets
            if dp_kind != "nvl":
                raise
            print(f"bench: NVLink exchange unavailable ({exc!r}); falling back to NCCL", file=sys.stderr)
            dp_kind = "nccl"
            dpc = dp_check(dev, rank, world, dp_kind)
    model, target, learner = make_dp_learner(dev, world, B, not a.no_graph, dp_kind)
    host = [synthetic_quads(B, seed=1 + rank + 97 * i) for i in range(3)]
    pool = [[t.to(dev) for t in b] for b in host]          # device-resident batches (> L2 together)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---------------- N > 1: every rank's step time WITHOUT the exchange (its own GPU at its own clocks).  A
    # data-parallel step ends when the slowest rank is done: t(rank 0) / max_r t(r) bounds the weak-scaling
    # efficiency whatever the exchange costs (B200s under a power cap differ by a few percent).
    local_ms = None
    if world > 1:
        sync, learner.grad_sync = learner.grad_sync, None
        for i in range(4):
            learner.load_batch(pool[i % len(pool)]); learner.step()
        torch.cuda.synchronize(); dist.barrier()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for i in range(20):
            learner.load_batch(pool[i % len(pool)]); learner.step()
        l1.record(); torch.cuda.synchronize()
        mine = torch.tensor([l0.elapsed_time(l1) / 20], device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        local_ms = [round(float(t.item()), 4) for t in allr]
        learner.grad_sync = sync
        learner._graphs.clear()               # the captured step had no exchange in it: capture again
        learner._eager_steps = 0

    # ---------------- device-resident timing ("value")
    sampler = ClockSampler(local)
    sampler.start()
    learner.model._state(); learner.target_net._state()
    torch.cuda.synchronize()
    lc0 = lib.vdqn_launch_count()
    learner.load_batch(pool[0])
    learner.step()                             # first step runs eagerly: counts one step's launches
    per_step_launches = lib.vdqn_launch_count() - lc0
    for i in range(1, a.warmup):
        learner.load_batch(pool[i % len(pool)])
        learner.step()
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_region0 = time.perf_counter()
    e0.record()
    for i in range(a.steps):
        learner.load_batch(pool[i % len(pool)])
        learner.step()
    e1.record()
    barrier()
    t_region1 = time.perf_counter()
    ms = max_over_ranks(e0.elapsed_time(e1)) / a.steps
    clocks = sampler.stop(t_region0, t_region1)
    loss_dev = float(learner.loss.item())

    # ---------------- per-kernel roofline (eager instrumented steps, CUDA events on the launch stream)
    roof, roof_other, breakdown, conv_detail = None, [], None, None
    pk, pk_kind = peaks()
    if rank == 0:
        learner_use_graph, saved_sync = learner.use_graph, learner.grad_sync
        learner.use_graph = False
        learner.grad_sync = None           # rank-local pass: no collectives (other ranks are not in it)
        ops.PROFILE = []
        nprof = 10
        marks = []
        for i in range(nprof):
            learner.load_batch(pool[i % len(pool)])
            learner.step()
            marks.append(len(ops.PROFILE))
        torch.cuda.synchronize()
        # per kernel family: the MEDIAN over the instrumented steps of that step's summed launch time
        # (x nprof, so the per-step / per-launch arithmetic below is unchanged)
        per_step, lo = [], 0
        for hi in marks:
            d = {}
            for kind, tag, s, e in ops.PROFILE[lo:hi]:
                t, n = d.get(kind, (0.0, 0))
                d[kind] = (t + s.elapsed_time(e), n + 1)
            per_step.append(d)
            lo = hi
        agg = {k: (statistics.median(d[k][0] for d in per_step if k in d) * nprof,
                   sum(d[k][1] for d in per_step if k in d)) for k in per_step[0]}
        learner.use_graph, learner.grad_sync = learner_use_graph, saved_sync
        step_ms_eager = sum(t for t, _ in agg.values()) / nprof
        breakdown = {k: {"ms_per_step": round(t / nprof, 4), "launches_per_step": n / nprof}
                     for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])}
        conv_detail = {}
        for kind, tag, s_, e_ in ops.PROFILE:
            if kind in ("igemm", "wgrad"):
                key = f"{kind}:{tag}"
                t0, n0 = conv_detail.get(key, (0.0, 0))
                conv_detail[key] = (t0 + s_.elapsed_time(e_), n0 + 1)
        conv_detail = {k: [round(t / n * 1e3, 1), n // nprof] for k, (t, n) in conv_detail.items()}

        def entry(kind, bound, work_per_step, unit_scale, peak, unit):
            t, n = agg[kind]
            per_launch_ms = t / n
            launches_per_step = n / nprof
            achieved = work_per_step / launches_per_step / (per_launch_ms * 1e-3) / unit_scale
            return {"kernel": kind, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                    "frac": achieved / peak, "traffic": None, "launches_per_step": launches_per_step,
                    "avg_launch_ms": per_launch_ms, "share_of_step": (t / nprof) / ms,
                    "peak_source": f"{pk_kind} ({'bf16_tflops_sustained' if bound == 'tensor' else 'hbm_gbs'})"}

        roof = entry("igemm", "tensor", IGEMM_FLOP_PER_QUAD * B, 1e12, pk["bf16_tflops_sustained"], "TFLOP/s")
        roof_other.append(entry("wgrad", "tensor", WGRAD_FLOP_PER_QUAD * B, 1e12,
                                pk["bf16_tflops_sustained"], "TFLOP/s"))
        roof_other.append(entry("adam", "hbm", 28.0 * N_PARAMS, 1e9, pk["hbm_gbs"], "GB/s"))
        roof_other.append(entry("td", "hbm", 370.0 * B, 1e9, pk["hbm_gbs"], "GB/s"))
        # TD epilogue at an HBM-measurable size (B = 2^20)
        try:
            nb = 1 << 20
            q = [torch.randn(nb, 5, 3, device=dev) for _ in range(3)]
            act = torch.randint(0, 3, (nb,), device=dev); rw = (torch.rand(nb, 5, device=dev) < 0.1).long()
            # outputs preallocated: the timed region holds the kernel launches only (tools/td_bandwidth.py)
            dq_big, loss_big = torch.empty_like(q[0]), torch.zeros(1, device=dev)
            for _ in range(3):
                ops.td_epilogue(q[0], q[1], q[2], act, rw, rw, dq=dq_big, loss=loss_big)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                ops.td_epilogue(q[0], q[1], q[2], act, rw, rw, dq=dq_big, loss=loss_big)
            e.record(); torch.cuda.synchronize()
            t_ms = s.elapsed_time(e) / 10
            del dq_big, loss_big
            byts = nb * (3 * 60 + 8 + 40 + 40 + 60)
            roof_other.append({"kernel": "td@B=2^20", "bound": "hbm", "achieved": byts / t_ms / 1e6,
                               "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": byts / t_ms / 1e6 / pk["hbm_gbs"],
                               "traffic": None, "avg_launch_ms": t_ms})
            del q, act, rw
        except Exception as ex:  # noqa
            roof_other.append({"kernel": "td@B=2^20", "error": repr(ex)})
        ops.PROFILE = None
        ncu = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(ncu):
            tr = json.load(open(ncu))
            if roof["kernel"] in tr:
                roof["traffic"] = tr[roof["kernel"]]
                roof["traffic_source"] = ("profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per "
                                          "launch from the committed ncu launch list of this command (not measured in "
                                          "this run: ncu cannot run inside a timed bench)")
            for r in roof_other:
                if r.get("kernel") in tr:
                    r["traffic"] = tr[r["kernel"]]

    # ---------------- end-to-end through the public API from pinned host buffers
    e2e = None
    if not a.no_e2e:
        stager = BatchStager(learner)
        k = a.steps
        for i in range(2):
            stager.push(host[i % len(host)])
        for i in range(3):                                   # warm-up of the staged path
            stager.pop_into_learner(); learner.step(); _ = learner.loss.item()
            stager.push(host[(i + 2) % len(host)])
        barrier()
        t0 = time.perf_counter()
        base = learner.steps_done
        for i in range(k):
            stager.pop_into_learner()
            learner.step()
            stager.push(host[(i + 5) % len(host)])           # next batch's H2D overlaps this step
            # D2H of the step result, every step (:229): each loss is copied to pinned memory by the
            # step itself and read here one step behind, so the device is not idle while the host reads
            if i > 0:
                _ = learner.loss_value(base + i - 1)
        _ = learner.loss_value(base + k - 1)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0) / k
        e2e = {"value": world * B * 2 / dt, "unit": "frames/s", "h2d_bytes_per_step": stager.h2d_bytes,
               "d2h_bytes_per_step": 4, "ms_per_step": dt * 1e3, "steps_per_sec": 1.0 / dt,
               "frames": "uint8 HWC (to_imgnet fused into the stem-pack kernel)"}

    # ---------------- forward-only callers (BASELINE configs[3]): value-map batches of 32 views
    # (visualize_value.py:78-98) and batch-1 policy scoring of a uint8 frame (evaluate.py:110-114)
    inference = None
    if rank == 0 and world == 1 and not a.no_inference:
        from video_dqn_b200.inference import QValueRunner
        model.eval()
        inference = {}
        for nb in (32, 1):
            run = QValueRunner(model, nb, frames_uint8=True)
            hostf = torch.empty(nb, 224, 224, 3, dtype=torch.uint8, pin_memory=True).random_(0, 256)
            devf = hostf.to(dev)
            for _ in range(5):
                run(devf)
            torch.cuda.synchronize()
            iters = 200
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                run()
            e1.record(); torch.cuda.synchronize()
            dev_ms = e0.elapsed_time(e1) / iters
            t0 = time.perf_counter()
            for _ in range(iters):
                _q, v, _b = run(hostf)                 # H2D of the uint8 views + forward
                _ = v[0, 0].item()                      # D2H of the result, every call (evaluate.py:114)
            host_ms = (time.perf_counter() - t0) / iters * 1e3
            inference[f"batch{nb}"] = {"device_ms": dev_ms, "views_per_sec_device": nb / dev_ms * 1e3,
                                       "e2e_ms": host_ms, "views_per_sec_e2e": nb / host_ms * 1e3}
        model.set_train()

    # ---------------- inverse-dynamics model (BASELINE configs[4]): training step on frame pairs at the
    # reference's batch (train_inverse_model.py:21,93-110) and the labelling forward
    # (dataset/process_episodes_real.py:171-179)
    inverse = None
    if rank == 0 and world == 1 and not a.no_inference:
        try:
            inverse = _inverse_block(dev)
        except Exception as exc:                      # an auxiliary block must not take the headline line down
            inverse = {"error": repr(exc)}
    basic = None
    if rank == 0 and world == 1 and not a.no_inference:
        try:
            basic = _basic_block(dev, B)
        except Exception as exc:
            basic = {"error": repr(exc)}

    # ---------------- CPU baseline (rank 0, N = 1): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        dtc = cpu_reference_arm(steps=5, warmup=3, batch=8, threads=threads)
        cpu = {"value": 16 / dtc, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "median of 5 steps of 8 quadruplets after 3 warm-up steps (of the 256-quadruplet "
                         "workload), fp32 oracle port",
               "ms_per_step_b8": dtc * 1e3}
        # also the reference's real batch (16, train_q_network.py:98) on all cores, and the
        # as-shipped thread setting (torch.set_num_threads(1), :85)
        dt16 = cpu_reference_arm(steps=5, warmup=3, batch=16, threads=threads)
        dt1 = cpu_reference_arm(steps=5, warmup=1, batch=8, threads=1)
        cpu["batch16_all_cores"] = {"value": 32 / dt16, "unit": "frames/s", "ms_per_step": dt16 * 1e3}
        cpu["as_shipped_1_thread"] = {"value": 16 / dt1, "unit": "frames/s", "cores": 1, "ms_per_step_b8": dt1 * 1e3}

    if rank == 0 and a.detail and conv_detail:
        for k, v in sorted(conv_detail.items(), key=lambda kv: -kv[1][0] * kv[1][1]):
            print(f"{v[0]:9.1f} us x{v[1]:2d}  {k}", file=sys.stderr)
    if rank == 0:
        fps = world * B * 2 / (ms * 1e-3)
        out = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "steps_per_sec": 1e3 / ms,
            "quadruplets_per_sec": world * B / (ms * 1e-3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": make_config(world, B, not a.no_graph),
            "tensor_pipe_frac_of_step": STEP_FLOP_PER_QUAD * B / (ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
            "step_tflops": STEP_FLOP_PER_QUAD * B / (ms * 1e-3) / 1e12,
            "loss": loss_dev,
            "clocks": clocks, "gpu_launches": (per_step_launches or 0) * a.steps,
            "gpu_launches_per_step": per_step_launches,
            "dp_check": dpc, "grad_exchange": dp_kind,
            "per_rank_ms_without_exchange": local_ms,
            "slowest_rank_bound_on_efficiency": (min(local_ms) / max(local_ms)) if local_ms else None,
            "e2e": e2e, "roofline": roof, "roofline_kernels": roof_other, "cpu_baseline": cpu,
            "breakdown_eager_ms": breakdown if rank == 0 else None,
            "inference": inference, "inverse_model": inverse, "basic_architecture": basic,
        }
        print(json.dumps(out), file=out_stream, flush=True)
    if world > 1:
        barrier()
        sys.stdout.flush(); sys.stderr.flush()
        if dp_kind == "nvl":
            del learner
            dist.destroy_process_group()          # nothing of NCCL's is captured in the step graph: a clean teardown
        else:
            # NCCL fallback only (VDQN_DDP=nccl, or no symmetric memory): NCCL collectives captured in CUDA graphs
            # keep the communicator busy at teardown -- destroy_process_group() blocks even after the graphs are
            # released, the learner deleted and the device synchronised (measured again in round 2: a 300 s
            # hang at N = 2).  Everything is reported and flushed, so this path leaves without the destructor;
            # the default NVLink-kernel path above tears down normally.
            os._exit(0)


if __name__ == "__main__":
    main()
