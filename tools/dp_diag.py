"""Where does the data-parallel step time go?  (torchrun, N ranks)  Times, as CUDA-graph replays, max over ranks:
(a) the fwd+bwd step without the gradient exchange, (b) the gradient-bucket all-reduce alone, (c) the full step."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import torch.distributed as dist
import torch_b200 as tlt
from torch_b200.graphs import GraphedStep
from torch_b200.parallel import GradientBucket

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(1234)
up = tlt.FactorizedLinear((8, 8, 12), (8, 12, 32), bias=True, factorization="blocktt", rank=16, device=dev).to(torch.bfloat16)
down = tlt.FactorizedLinear((8, 12, 32), (8, 8, 12), bias=True, factorization="blocktt", rank=16, device=dev).to(torch.bfloat16)
params = list(up.parameters()) + list(down.parameters())
x = torch.randn(32768, 768, device=dev, dtype=torch.bfloat16).requires_grad_(True)
gy = torch.randn(32768, 768, device=dev, dtype=torch.bfloat16)
bucket = GradientBucket(params)


def compute():
    x.grad = None
    for p in params:
        p.grad = None
    down(up(x)).backward(gy)


def exchange():
    bucket.allreduce()


def full():
    compute()
    exchange()


def timed(fn, steps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    run = GraphedStep(fn).replay
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    mn = ms.clone()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    return ms.item(), mn.item()


compute()
res = {"compute": timed(compute), "exchange": timed(exchange), "full": timed(full)}
if rank == 0:
    print({k: (round(v[0], 4), round(v[1], 4)) for k, v in res.items()}, "ms (max over ranks, min over ranks); world", world, flush=True)
os._exit(0)
