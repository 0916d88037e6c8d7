"""Micro-benchmark of the fused BlockTT kernels and the bias-gradient mat-vec at BASELINE config 2 shapes.

Development aid (run on the GPU box):  python tools/blocktt_bench.py
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from torch_b200 import blocktt_ops as bo
from torch_b200 import ops

dev = torch.device("cuda:0")


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, rows, cols in (("up   W(3072,768)", (8, 12, 32), (8, 8, 12)), ("down W(768,3072)", (8, 8, 12), (8, 12, 32))):
    for dtype in (torch.bfloat16, torch.float32):
        shapes = [(1, rows[0], cols[0], 16), (16, rows[1], cols[1], 16), (16, rows[2], cols[2], 1)]
        cores = [torch.randn(s, device=dev).to(dtype) for s in shapes]
        R, Cn = rows[0] * rows[1] * rows[2], cols[0] * cols[1] * cols[2]
        W = torch.empty(R, Cn, device=dev, dtype=dtype)
        dW = torch.randn(R, Cn, device=dev, dtype=torch.float32)
        d = [torch.empty_like(c) for c in cores]
        t_rec = timeit(lambda: bo._execute_btt3_reconstruct(*cores, W, rows, cols, 16, 16))
        t_grad = timeit(lambda: bo._execute_btt3_core_grads(*cores, dW, *d, rows, cols, 16, 16))
        print(f"{name} {str(dtype)[6:]:9s} reconstruct {t_rec:7.1f} us   core_grads (memset + A + B) {t_grad:7.1f} us", flush=True)

for n_out in (3072, 768):
    gy = torch.randn(32768, n_out, device=dev, dtype=torch.bfloat16)
    t = timeit(lambda: ops.reduce_to(gy, "bo", "o"))
    print(f"bias grad colsum (32768 x {n_out}) bf16: {t:7.1f} us  = {gy.numel() * 2 / t / 1e3:7.1f} GB/s", flush=True)
