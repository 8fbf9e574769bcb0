import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
def log(*a):
    print(f"[rank {os.environ.get('RANK')}] {time.time()%1000:8.2f}", *a, file=sys.stderr, flush=True)
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
log("pg up")
from video_dqn_b200.ddp import GradSync
from video_dqn_b200.learner import QLearner, StepConfig
from video_dqn_b200.qnet import HabitatDQNMultiAction
torch.manual_seed(4)
B = 32
model = HabitatDQNMultiAction(3, 5, extra_capacity=True, panorama=False).to(dev)
target = HabitatDQNMultiAction(3, 5, extra_capacity=True, panorama=False).to(dev)
target.load_state_dict(model.state_dict()); target.eval()
use_graph = os.environ.get("VDQN_GRAPH", "1") == "1"
lr = QLearner(model, target, StepConfig(), batch_size=B, frames_uint8=True, use_graph=use_graph, world_size=world)
lr.grad_sync = GradSync(lr)
log("learner built")
g = torch.Generator().manual_seed(1 + rank)
batch = (torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, generator=g).to(dev),
         torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, generator=g).to(dev),
         torch.randint(0, 3, (B,), generator=g).to(dev), torch.zeros(B, 5, dtype=torch.long, device=dev),
         torch.zeros(B, 5, dtype=torch.long, device=dev), None, torch.ones(B, 5, dtype=torch.long, device=dev))
for i in range(4):
    log("step", i, "begin")
    l = lr.step(batch)
    torch.cuda.synchronize()
    log("step", i, "loss", l.item(), "buckets", lr.grad_sync.launched[-8:])
# gradient agreement across ranks after the all-reduce
gsum = lr.opt.grad_arena.double().sum().item()
t = torch.tensor([gsum], device=dev, dtype=torch.float64)
lst = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(lst, t)
log("grad arena sums per rank", [x.item() for x in lst])
psum = lr.opt.param_arena.double().sum().item()
t = torch.tensor([psum], device=dev, dtype=torch.float64); dist.all_gather(lst, t)
log("param arena sums per rank", [x.item() for x in lst])
dist.barrier(); torch.cuda.synchronize()
log("done")
sys.stderr.flush(); os._exit(0)
