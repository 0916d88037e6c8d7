"""Time tlt_pointwise_tc alone (CUDA events) on config 3's 56^2 / 28^2 1x1 shapes (NCHW activations); TLT_PW_T=on|off forces the
transposed-role kernel."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_b200 import pointwise_ops  # noqa: E402

dev = torch.device("cuda:0")


def timed(fn, steps=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for (B, K, N, S) in [(256, 64, 69, 3136), (256, 69, 64, 3136), (256, 128, 141, 784), (256, 141, 128, 784), (32, 64, 69, 3136), (64, 64, 69, 3136),
                     (512, 64, 69, 3136)]:
    x = torch.randn(B, K, S, device=dev, dtype=torch.bfloat16)
    w = (torch.randn(N, K, device=dev) * 0.1).bfloat16()
    y = torch.empty(B, N, S, device=dev, dtype=torch.bfloat16)
    ms = timed(lambda: pointwise_ops._execute_pointwise(x, w, None, y))
    byts = (x.numel() + y.numel()) * 2
    print(json.dumps({"case": f"pointwise_tc B={B} K={K} N={N} S={S} T={os.environ.get('TLT_PW_T', 'off')}",
                      "ms": ms, "GBps": byts / ms / 1e6}), flush=True)
