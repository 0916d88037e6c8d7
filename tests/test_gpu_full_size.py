"""Full-size property tests (run on the B200 with -m gpu): the BASELINE.json configurations at their real sizes, checked
through size-independent properties -- agreement with the reconstructed dense weight on a sample of rows / images in fp64,
linearity in x, and gradient identities -- because the CPU oracle cannot finish these sizes in seconds.
(config 2 lives in test_gpu_parity.py::test_config2_full_size_properties.)"""
import pytest
import torch
import torch.nn.functional as F

import _golden as G

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def test_config1_cp_linear_fp32_full_size():
    """FactorizedLinear 512->512, CP rank 0.5 (R = 2341), batch 128, fp32: y == x to_matrix()^T + b and the x / bias gradients."""
    import torch_b200 as tlt
    torch.manual_seed(1)
    layer = tlt.FactorizedLinear((4, 8, 16), (4, 8, 16), bias=True, factorization="cp", rank=0.5, device=dev())
    assert int(layer.weight.rank) == 2341
    x = torch.randn(128, 512, device=dev(), requires_grad=True)
    y = layer(x)
    W = layer.weight.to_matrix().double()
    want = x.detach().double() @ W.t() + layer.bias.double()
    assert G.rel_err(y, want.cpu()) < 1e-5
    gy = torch.randn_like(y)
    gx, gb = torch.autograd.grad(y, [x, layer.bias], gy)
    assert G.rel_err(gx, (gy.double() @ W).cpu()) < 1e-5
    assert G.rel_err(gb, gy.double().sum(0).cpu()) < 1e-5


def test_config5_tucker_linear_full_size():
    """FactorizedLinear 4096->4096, Tucker rank 0.75 (ranks 15^6), 8192 tokens, bf16 (fused mmd16 -> tcgen05 -> mmd16 route)."""
    import torch_b200 as tlt
    torch.manual_seed(5)
    ts = (16, 16, 16)
    layer = tlt.FactorizedLinear(ts, ts, bias=True, factorization="tucker", rank=0.75, device=dev()).to(torch.bfloat16)
    B = 8192
    x = torch.randn(B, 4096, device=dev(), dtype=torch.bfloat16, requires_grad=True)
    y = layer(x)
    W = layer.weight.to_matrix().double()
    rows = torch.randint(0, B, (48,), device=dev())
    want = x.detach()[rows].double() @ W.t() + layer.bias.double()
    assert G.rel_err(y[rows], want.cpu()) < 2e-2
    y2 = layer(2 * x.detach())                                         # linearity (scaling by 2 is exact in bf16)
    assert G.rel_err(y2.float() - layer.bias.float(), (2 * (y.detach().float() - layer.bias.float())).double().cpu()) < 2e-2
    gy = torch.randn_like(y)
    gx, gb = torch.autograd.grad(y, [x, layer.bias], gy)
    assert G.rel_err(gx[rows], (gy[rows].double() @ W).cpu()) < 2e-2
    assert G.rel_err(gb, gy.double().sum(0).cpu()) < 2e-2


def test_config4_tucker_conv_full_size():
    """Tucker FactorizedConv 512->512 3x3 on (512, 512, 7, 7), rank 0.5 -> (388, 388, 2, 2), bf16 (channels-last rows route)."""
    import torch_b200 as tlt
    torch.manual_seed(4)
    conv = tlt.FactorizedConv(512, 512, 3, order=2, padding=1, factorization="tucker", rank=0.5, bias=True, device=dev())
    conv.reset_parameters(std=0.2)
    conv.bias.data.normal_()
    conv = conv.to(torch.bfloat16)
    assert tuple(conv.weight.rank) == (388, 388, 2, 2)
    x = torch.randn(512, 512, 7, 7, device=dev(), dtype=torch.bfloat16, requires_grad=True)
    y = conv(x)
    Wd = conv.weight.to_tensor().double()
    imgs = torch.randint(0, 512, (6,), device=dev())
    want = F.conv2d(x.detach()[imgs].double(), Wd, conv.bias.double(), padding=1)
    assert G.rel_err(y[imgs], want.cpu()) < 2e-2
    gy = torch.randn_like(y)
    gx, gb = torch.autograd.grad(y, [x, conv.bias], gy)
    want_gx = F.conv_transpose2d(gy[imgs].double(), Wd, padding=1)
    assert G.rel_err(gx[imgs], want_gx.cpu()) < 2e-2
    assert G.rel_err(gb, gy.double().sum((0, 2, 3)).cpu()) < 2e-2


def test_config4_trl_full_size():
    """TRL (2048,7,7) -> 1000, Tucker rank (64,4,4,10), batch 512, bf16 (rows projection route)."""
    import torch_b200 as tlt
    torch.manual_seed(6)
    trl = tlt.TRL((2048, 7, 7), (1000,), bias=True, factorization="tucker", rank=(64, 4, 4, 10), device=dev())
    with torch.no_grad():
        for p in trl.parameters():
            p.mul_(0.3)
    trl = trl.to(torch.bfloat16)
    x = torch.randn(512, 2048, 7, 7, device=dev(), dtype=torch.bfloat16, requires_grad=True)
    y = trl(x)
    Wd = trl.weight.to_tensor().double().reshape(-1, 1000)
    rows = torch.randint(0, 512, (24,), device=dev())
    want = x.detach()[rows].double().reshape(24, -1) @ Wd + trl.bias.double()
    assert G.rel_err(y[rows], want.cpu()) < 2e-2
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, [x], gy)
    want_gx = (gy[rows].double() @ Wd.t()).reshape(24, 2048, 7, 7)
    assert G.rel_err(gx[rows], want_gx.cpu()) < 2e-2


@pytest.mark.parametrize("cin,cout,hw,stride", [(64, 64, 56, 1), (128, 128, 28, 1), (64, 128, 56, 2), (256, 256, 14, 1), (512, 512, 7, 1)])
def test_config3_cp_conv_full_size(cin, cout, hw, stride):
    """CP FactorizedConv layers of ResNet-18 (rank 0.25) at batch 256, bf16: NCHW tensor-core route (56^2), rows route (strided /
    14^2 / 7^2) vs F.conv2d with the reconstructed kernel on a sample of images; dx vs conv_transpose2d."""
    import torch_b200 as tlt
    torch.manual_seed(hw + stride)
    conv = tlt.FactorizedConv(cin, cout, 3, order=2, padding=1, stride=stride, factorization="cp", rank=0.25, bias=False, device=dev())
    conv.reset_parameters(std=0.5)
    conv = conv.to(torch.bfloat16)
    x = torch.randn(256, cin, hw, hw, device=dev(), dtype=torch.bfloat16, requires_grad=True)
    y = conv(x)
    Wd = conv.weight.to_tensor().double()
    imgs = torch.randint(0, 256, (3,), device=dev())
    want = F.conv2d(x.detach()[imgs].double(), Wd, None, stride=stride, padding=1)
    assert G.rel_err(y[imgs], want.cpu()) < 2e-2
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, [x], gy)
    want_gx = F.conv_transpose2d(gy[imgs].double(), Wd, stride=stride, padding=1, output_padding=stride - 1)
    assert G.rel_err(gx[imgs], want_gx.cpu()) < 2e-2
