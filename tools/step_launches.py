"""Runs a few eager fwd+bwd steps of one named layer (for `ncu --metrics gpu__time_duration.sum` launch lists).

    python tools/step_launches.py {trl|tucker_conv|cp_conv:<cin>:<cout>:<hw>:<stride>|tucker_linear|cp_linear} [steps]
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import torch_b200 as tlt

dev = torch.device("cuda:0")
torch.manual_seed(2)
what = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
if what == "trl":
    layer = tlt.TRL((2048, 7, 7), (1000,), bias=True, factorization="tucker", rank=(64, 4, 4, 10), device=dev, dtype=torch.float32)
    with torch.no_grad():
        for p in layer.parameters():
            p.mul_(0.05)
    layer = layer.to(torch.bfloat16)
    x = torch.randn(512, 2048, 7, 7, device=dev, dtype=torch.bfloat16)
elif what in ("tucker_conv", "tucker_conv_cl"):
    layer = tlt.FactorizedConv(512, 512, 3, order=2, padding=1, factorization="tucker", rank=0.5, bias=True, device=dev)
    layer.reset_parameters()
    layer = layer.to(torch.bfloat16)
    x = torch.randn(512, 512, 7, 7, device=dev, dtype=torch.bfloat16)
    if what.endswith("_cl"):
        x = x.contiguous(memory_format=torch.channels_last)
elif what.startswith("cp_conv"):
    _, cin, cout, hw, stride = what.split(":")
    layer = tlt.FactorizedConv(int(cin), int(cout), 3, order=2, padding=1, stride=int(stride), factorization="cp", rank=0.25, bias=False,
                               device=dev)
    layer.reset_parameters()
    layer = layer.to(torch.bfloat16)
    x = torch.randn(256, int(cin), int(hw), int(hw), device=dev, dtype=torch.bfloat16)
    if what.startswith("cp_conv_cl"):          # channels-last activations: the rows route of config 3
        x = x.contiguous(memory_format=torch.channels_last)
elif what == "tucker_linear":
    layer = tlt.FactorizedLinear((16, 16, 16), (16, 16, 16), bias=True, factorization="tucker", rank=0.75, device=dev).to(torch.bfloat16)
    x = torch.randn(8192, 4096, device=dev, dtype=torch.bfloat16)
elif what == "cp_linear":
    layer = tlt.FactorizedLinear((4, 8, 16), (4, 8, 16), bias=True, factorization="cp", rank=0.5, device=dev)
    x = torch.randn(128, 512, device=dev)
else:
    raise SystemExit(f"unknown layer {what}")
x.requires_grad_(True)
gy = None
for _ in range(steps):
    y = layer(x)
    if gy is None:
        gy = torch.randn_like(y)
    torch.autograd.grad(y, [x] + list(layer.parameters()), gy)
torch.cuda.synchronize()
print("ok", what, steps)
