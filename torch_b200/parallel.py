"""Data-parallel plumbing: one process per GPU, batch sharded across ranks, parameters replicated.

The factorized-layer hot path has exactly one exchange step per iteration: the sum all-reduce of the
small factor / core / bias gradients.  They are packed into ONE flat fp32 bucket so that a single NCCL call (NVLink 5 /
NVSwitch, NVLS in-switch reduction when available) carries them -- at <= 50 MB the exchange is latency-bound, so one
call beats per-tensor calls.  Works with any torch.distributed backend (the CPU tests use gloo, world_size 2).
"""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


class GradientBucket:
    """Flat fp32 bucket holding the gradients of ``params`` (replicated parameters of the factorized layers)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view(p.shape))
            off += p.numel()

    def pack(self):
        """gradients -> flat fp32 bucket.  Grouped by dtype: one concatenation + one cast per group, not one copy per tensor."""
        groups = {}
        for i, p in enumerate(self.params):
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            groups.setdefault(g.dtype, []).append((i, g))
        offs = self._offsets()
        for dtype, items in groups.items():
            contiguous_run = all(items[k + 1][0] == items[k][0] + 1 for k in range(len(items) - 1))
            if contiguous_run and len(items) > 1:
                lo = offs[items[0][0]]
                hi = offs[items[-1][0]] + self.params[items[-1][0]].numel()
                self.flat[lo:hi].copy_(torch.cat([g.reshape(-1) for _, g in items]))
            else:
                for i, g in items:
                    self.views[i].copy_(g)

    def _offsets(self):
        offs, off = [], 0
        for p in self.params:
            offs.append(off)
            off += p.numel()
        return offs

    def unpack(self, scale: float = 1.0):
        if scale != 1.0:
            self.flat.mul_(scale)
        have = [p for p in self.params if p.grad is not None]
        if len(have) == len(self.params):
            torch._foreach_copy_([p.grad for p in self.params], self.views)      # one fused launch per dtype group
            return
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                p.grad = v.to(p.dtype).clone()
            else:
                p.grad.copy_(v)

    def allreduce(self, group: Optional[dist.ProcessGroup] = None, average: bool = True):
        """grads <- sum (or mean) over ranks.  No-op without an initialised process group."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        self.pack()
        world = dist.get_world_size(group)
        if average and dist.get_backend(group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)       # the mean is taken inside the collective
            self.unpack(1.0)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.unpack(1.0 / world if average else 1.0)


    # ---- asynchronous form: issue the exchange as soon as this bucket's gradients are final, finish it later ---------------------
    def allreduce_async(self, group: Optional[dist.ProcessGroup] = None, average: bool = True):
        """Pack and start the all-reduce; returns a handle for ``finish``.  The collective runs on the backend's own stream
        (NCCL) after everything queued so far on the current stream, so kernels launched between this call and ``finish`` --
        the backward pass of the layers whose gradients are not in this bucket -- overlap the exchange.  Capturable in a CUDA
        graph (the fork / join become graph dependencies)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        self.pack()
        world = dist.get_world_size(group)
        if average and dist.get_backend(group) == "nccl":
            return (dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group, async_op=True), 1.0)
        return (dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=True), 1.0 / world if average else 1.0)

    def finish(self, handle):
        """Wait (on the current stream) for the exchange started by ``allreduce_async`` and write the reduced gradients back."""
        if handle is None:
            return
        work, scale = handle
        work.wait()
        self.unpack(scale)


def shard_batch(total: int, rank: int, world_size: int):
    """[begin, end) of this rank's slice of a batch of ``total`` independent samples (ragged tail goes to low ranks)."""
    base, rem = divmod(total, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)
