"""Micro-benchmark of individual contraction launches (development aid; run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from torch_b200.ops import contract

dev = torch.device("cuda:0")
CASES = [
    # BlockTT (8,12,32)x(8,8,12) r16 reconstruction and its backward
    ("XcdY,YefZ->XcedfZ", (1, 8, 8, 16), (16, 12, 8, 16)),
    ("XcedfY,YghZ->XcegdfhZ", (1, 8, 12, 8, 8, 16), (16, 32, 12, 1)),
    ("XcegdfhZ,YghZ->XcedfY", (1, 8, 12, 32, 8, 8, 12, 1), (16, 32, 12, 1)),
    ("XcedfY,XcegdfhZ->YghZ", (1, 8, 12, 8, 8, 16), (1, 8, 12, 32, 8, 8, 12, 1)),
    ("XcedfZ,YefZ->XcdY", (1, 8, 12, 8, 8, 16), (16, 12, 8, 16)),
    ("XcdY,XcedfZ->YefZ", (1, 8, 8, 16), (1, 8, 12, 8, 8, 16)),
    ("bo,b->o", (32768, 3072), (32768,)),
    ("bo,b->o", (32768, 768), (32768,)),
]
for dtype in (torch.bfloat16, torch.float32):
    for spec, sa, sb in CASES:
        a = torch.randn(sa, device=dev, dtype=dtype)
        b = torch.randn(sb, device=dev, dtype=dtype)
        for _ in range(3):
            contract(spec, a, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            contract(spec, a, b)
        e1.record()
        torch.cuda.synchronize()
        print(f"{str(dtype):16s} {spec:28s} {e0.elapsed_time(e1) / n * 1e3:9.1f} us")
