"""Smallest end-to-end case of every path the headline uses, for compute-sanitizer:
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_case.py
fused channels-last CP conv (fwd + both backward kernels), the unfused rows chain (GEMM / separable sweep / GEMM, stride 1 and 2)
and a BlockTT linear, one image / a few rows each, fwd + bwd."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import torch_b200 as tlt

dev = torch.device("cuda:0")
torch.manual_seed(0)


def run_conv(cin, cout, hw, stride, batch=1):
    conv = tlt.FactorizedConv(cin, cout, 3, order=2, stride=stride, padding=1, factorization="cp", rank=0.25, bias=True, device=dev)
    conv.reset_parameters(std=0.1)
    conv = conv.to(torch.bfloat16)
    x = torch.randn(batch, cin, hw, hw, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = conv(x)
    g = torch.autograd.grad(y, [x] + list(conv.parameters()), torch.ones_like(y))
    torch.cuda.synchronize()
    assert all(torch.isfinite(t.float()).all() for t in g)
    print(f"conv {cin}->{cout} @{hw} s{stride}: ok", flush=True)


run_conv(64, 64, 56, 1)          # tlt_cpconv_rows_* (second-generation kernels)
run_conv(64, 128, 56, 2)         # chain: GEMM, stride-2 sweep, GEMM
run_conv(128, 128, 14, 1)        # chain, stride 1
lin = tlt.FactorizedLinear((8, 8, 12), (8, 12, 32), bias=True, factorization="blocktt", rank=16, device=dev).to(torch.bfloat16)
x = torch.randn(128, 768, device=dev, dtype=torch.bfloat16, requires_grad=True)
y = lin(x)
torch.autograd.grad(y, [x] + list(lin.parameters()), torch.ones_like(y))
torch.cuda.synchronize()
print("blocktt linear: ok", flush=True)
