"""Host-buffer front end for a factorized-layer step: overlaps PCIe copies with the kernels.

The hot path itself keeps every operand in HBM.  When the caller's data lives in HOST memory (the reference's own
calling convention on CPU tensors, and the ``e2e`` leg of bench.py), a step is

    H2D(inputs of step i+1)   |   kernels of step i   |   D2H(results of step i-1)

on three CUDA streams: the two DMA directions of the PCIe link run concurrently with each other and with compute, so
steady-state throughput is bounded by the slower of (host->device bytes / link rate) and the kernel time rather than
by their sum.  Inputs are staged in two device buffers per tensor; the captured step (``GraphedStep``) reads its own
static tensors, which are refreshed with device-to-device copies (HBM speed, ~1 % of a step) on the compute stream.

``replicas``: one (static inputs, step, results) triple PER pipeline slot -- e.g. the same step captured twice on two sets of static
tensors.  The host copies then land directly in a replica's own inputs and leave directly from its own result buffers: no staging
copies at all (at BASELINE config 3 the two staging passes move 446 MB per step, ~6 % of the end-to-end step).
"""
from typing import Callable, List, Sequence

import torch


class HostPipelinedStep:
    """Run ``step()`` -- a closure over the static device tensors ``dev_inputs`` producing ``get_outputs()`` -- on host data.

    ``submit(host_inputs, host_outputs)`` enqueues one step: pinned ``host_inputs[k]`` -> ``dev_inputs[k]``, step, result k
    -> pinned ``host_outputs[k]``.  Everything is asynchronous; ``drain()`` waits for all submitted work.  ``host_outputs``
    of a step are valid after ``drain()`` (or after the step two submissions later has been enqueued and synchronised).
    """

    DEPTH = 2

    def __init__(self, dev_inputs: Sequence[torch.Tensor], step: Callable[[], None],
                 get_outputs: Callable[[], List[torch.Tensor]], replicas=None):
        self.dev_inputs = list(dev_inputs)
        self.step = step
        self.get_outputs = get_outputs
        self.replicas = list(replicas) if replicas else None
        if self.replicas:
            self.DEPTH = len(self.replicas)
        dev = self.dev_inputs[0].device
        self.compute = torch.cuda.current_stream(dev)
        self.h2d = torch.cuda.Stream(dev)
        self.d2h = torch.cuda.Stream(dev)
        self.stage_in = None if self.replicas else [[torch.empty_like(t) for t in self.dev_inputs] for _ in range(self.DEPTH)]
        self.stage_out = [None] * self.DEPTH
        self.ev_loaded = [torch.cuda.Event() for _ in range(self.DEPTH)]      # H2D of slot finished
        self.ev_consumed = [torch.cuda.Event() for _ in range(self.DEPTH)]    # slot's staging inputs copied into the static tensors
        self.ev_computed = [torch.cuda.Event() for _ in range(self.DEPTH)]    # slot's results staged for D2H
        self.ev_stored = [torch.cuda.Event() for _ in range(self.DEPTH)]      # D2H of slot finished
        self.n = 0

    def submit(self, host_inputs: Sequence[torch.Tensor], host_outputs: Sequence[torch.Tensor]):
        slot = self.n % self.DEPTH
        reuse = self.n >= self.DEPTH
        if self.replicas:
            ins, step, get_outputs = self.replicas[slot]
            with torch.no_grad():
                with torch.cuda.stream(self.h2d):
                    if reuse:
                        self.h2d.wait_event(self.ev_computed[slot])     # the replica's previous step has read its inputs
                    for t, h in zip(ins, host_inputs):
                        t.copy_(h, non_blocking=True)
                    self.ev_loaded[slot].record(self.h2d)
                self.compute.wait_event(self.ev_loaded[slot])
                if reuse:
                    self.compute.wait_event(self.ev_stored[slot])       # its previous results have reached the host
            step()
            with torch.no_grad():
                outs = get_outputs()
                self.ev_computed[slot].record(self.compute)
                with torch.cuda.stream(self.d2h):
                    self.d2h.wait_event(self.ev_computed[slot])
                    for h, o in zip(host_outputs, outs):
                        h.copy_(o, non_blocking=True)
                    self.ev_stored[slot].record(self.d2h)
            self.n += 1
            return
        with torch.no_grad():
            # 1. host -> staging (copy stream); the slot's previous contents must have been consumed
            with torch.cuda.stream(self.h2d):
                if reuse:
                    self.h2d.wait_event(self.ev_consumed[slot])
                for s, h in zip(self.stage_in[slot], host_inputs):
                    s.copy_(h, non_blocking=True)
                self.ev_loaded[slot].record(self.h2d)
            # 2. staging -> static tensors, step, results -> output staging (compute stream)
            self.compute.wait_event(self.ev_loaded[slot])
            for t, s in zip(self.dev_inputs, self.stage_in[slot]):
                t.copy_(s)
            self.ev_consumed[slot].record(self.compute)
        self.step()
        with torch.no_grad():
            outs = self.get_outputs()
            if self.stage_out[slot] is None:
                self.stage_out[slot] = [torch.empty_like(o) for o in outs]
            if reuse:
                self.compute.wait_event(self.ev_stored[slot])
            for s, o in zip(self.stage_out[slot], outs):
                s.copy_(o)
            self.ev_computed[slot].record(self.compute)
            # 3. output staging -> host (second copy stream: the other DMA direction)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(self.ev_computed[slot])
                for h, s in zip(host_outputs, self.stage_out[slot]):
                    h.copy_(s, non_blocking=True)
                self.ev_stored[slot].record(self.d2h)
        self.n += 1

    def drain(self):
        self.h2d.synchronize()
        self.compute.synchronize()
        self.d2h.synchronize()
