"""CUDA-graph capture of a whole factorized-layer step.

A fwd+bwd pass of a factorized layer is a chain of short kernels (core contractions, the reconstruction, the GEMMs,
the bias reductions); issued eagerly from Python each launch costs tens of microseconds of host time, which at
BASELINE config 1/2 sizes rivals the GPU time.  Every kernel of libtltorch_cuda.so is launched on torch's current
stream with caller-owned buffers and no host synchronisation, so a step on static input tensors can be captured once
into a CUDA graph and replayed: one host call per step.
"""
import torch


class GraphedStep:
    """Capture ``fn()`` -- a closure over STATIC tensors (inputs are updated in place with ``copy_``) -- and replay it.

    >>> x = torch.randn(B, I, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    >>> def step():
    ...     for p in params: p.grad = None
    ...     x.grad = None
    ...     layer(x).backward(gy)
    >>> g = GraphedStep(step); g.replay()        # x.grad / p.grad hold the results after each replay
    """

    def __init__(self, fn, warmup=3, results=None):
        """``results``: optional callable evaluated right after the capture -- it returns the step's result tensors (``x.grad``,
        ``p.grad`` ...), which at that moment are the graph's own static buffers.  Keep ``self.results`` instead of re-reading the
        ``.grad`` attributes later: any eager step run afterwards rebinds them to fresh tensors the graph never writes."""
        self.fn = fn
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            fn()
        self.results = results() if results is not None else None

    def replay(self):
        self.graph.replay()

    __call__ = replay
