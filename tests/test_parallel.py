"""world_size-2 gloo test (CPU) of the data-parallel host logic: batch sharding + flat-bucket gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torch_b200.parallel import GradientBucket, shard_batch


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)),
              torch.nn.Parameter(torch.zeros(2, 2, 2, dtype=torch.bfloat16))]
    full = torch.arange(10, dtype=torch.float32)
    b, e = shard_batch(10, rank, world)
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(full[b:e].sum()) * (i + 1))
    bucket = GradientBucket(params)
    bucket.allreduce(average=False)
    res = [p.grad.float().mean().item() for p in params]
    # asynchronous form (what bench.py uses per ResNet stage): two buckets in flight at once, averaged
    groups = [GradientBucket(params[:2]), GradientBucket(params[2:])]
    handles = [g.allreduce_async(average=True) for g in groups]
    for g, h in zip(groups, handles):
        g.finish(h)
    res += [p.grad.float().mean().item() for p in params]
    out[rank] = res
    dist.destroy_process_group()


def test_gradient_bucket_allreduce_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    total = float(torch.arange(10).sum())
    for rank in (0, 1):
        # after the SUM exchange every rank holds the same gradients, so the AVG exchange that follows leaves them unchanged
        assert out[rank] == [total, 2 * total, 3 * total] * 2


def test_shard_batch_covers_everything_once():
    for total in (0, 1, 7, 32768):
        for world in (1, 2, 3, 8):
            spans = [shard_batch(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
