"""Fused 64->64 @56^2 CP conv layer: fwd and fwd+bwd graph time against the batch size (fixed cost per launch vs cost per image)."""
import math, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
import torch
import torch_b200 as tlt
import bench
dev = torch.device("cuda:0")
torch.manual_seed(1)
conv = tlt.FactorizedConv(64, 64, 3, order=2, padding=1, factorization="cp", rank=0.25, bias=False, device=dev)
conv.reset_parameters(std=1.0 / math.sqrt(9 * 64))
conv = conv.to(torch.bfloat16)
params = list(conv.parameters())
for B in (16, 32, 64, 128, 256, 512):
    x = torch.randn(B, 64, 56, 56, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    gy = torch.randn(B, 64, 56, 56, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    def step():
        y = conv(x)
        torch.autograd.grad(y, [x] + params, gy)
    def fwd():
        with torch.no_grad():
            conv(x)
    res = []
    for fn in (fwd, step):
        run, mode, launches, graphed = bench.time_graph(fn, 20, 3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(50): run()
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 50 * 1e3)
        del run, graphed
    print(f"batch {B:4d}: fwd {res[0]:7.1f} us   fwd+bwd {res[1]:7.1f} us   bwd {res[1]-res[0]:7.1f} us", flush=True)
