"""Development check + timing of the fused CP-conv rows kernels (tlt_cpconv_rows_*), one case per process.

    python tools/cprows_check.py --case fwd1|fwd2|bwd1 [--batch 256] [--steps 20]

Compares against the plain fp32 chain 1x1 -> depthwise 3x3 -> 1x1 (torch ops on the same bf16-rounded inputs) and its autograd,
then times the launches with CUDA events.  The parity tests proper are tests/test_gpu_parity.py (oracle).
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import torch.nn.functional as F
from torch_b200 import cprows_ops

dev = torch.device("cuda:0")


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def reference(x, lam, f_out, f_in, kh, kw, stride):
    """x (B, C, H, W) fp32; returns y (B, O, H/s, W/s) fp32 through the reference's op sequence."""
    r = lam.shape[0]
    z = F.conv2d(x, f_in.t().reshape(r, -1, 1, 1))
    z = F.conv2d(z, kh.t().reshape(r, 1, 3, 1), stride=(stride, 1), padding=(1, 0), groups=r)
    z = F.conv2d(z, kw.t().reshape(r, 1, 1, 3), stride=(1, stride), padding=(0, 1), groups=r)
    return F.conv2d(z, (f_out * lam).reshape(f_out.shape[0], r, 1, 1))


def timed(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default="fwd1")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--check-batch", type=int, default=11)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--height", type=int, default=56)
    args = ap.parse_args()
    stride = 2 if args.case.endswith("2") else 1
    c, o, r = (64, 64, 69) if stride == 1 else (64, 128, 93)
    H, W = args.height, 56
    torch.manual_seed(0)
    q = lambda t: t.to(torch.bfloat16)      # noqa: E731
    lam = q(1.0 + 0.1 * torch.randn(r, device=dev))
    f_out = q(torch.randn(o, r, device=dev) / r ** 0.5)
    f_in = q(torch.randn(c, r, device=dev) / c ** 0.5)
    kh = q(torch.randn(3, r, device=dev) * 0.6)
    kw = q(torch.randn(3, r, device=dev) * 0.6)
    params = [lam, f_out, f_in, kh, kw]

    for B in (args.check_batch, args.batch):
        x = q(torch.randn(B, c, H, W, device=dev)).contiguous(memory_format=torch.channels_last)
        gy = q(torch.randn(B, o, H // stride, W // stride, device=dev)).contiguous(memory_format=torch.channels_last)
        if args.case.startswith("fwd"):
            with torch.no_grad():
                y = cprows_ops.cp_conv_rows_fused(x, lam, [f_out, f_in, kh, kw], None, stride)
                torch.cuda.synchronize()
                if B == args.check_batch:
                    yr = reference(x.float(), *[t.float() for t in params], stride)
                    print(json.dumps({"case": args.case, "batch": B, "err_y": rel(y, yr)}), flush=True)
                else:
                    ms = timed(lambda: cprows_ops.cp_conv_rows_fused(x, lam, [f_out, f_in, kh, kw], None, stride), args.steps)
                    nbytes = (x.numel() + y.numel()) * 2
                    print(json.dumps({"case": args.case, "batch": B, "ms": ms, "alg_GBps": nbytes / ms / 1e6}), flush=True)
        else:
            xs = x.detach().requires_grad_(True)
            ps = [t.detach().requires_grad_(True) for t in params]
            y = cprows_ops.cp_conv_rows_fused(xs, ps[0], ps[1:], None, stride)
            grads = torch.autograd.grad(y, [xs] + ps, gy)
            torch.cuda.synchronize()
            if B == args.check_batch:
                xr = x.float().detach().requires_grad_(True)
                pr = [t.float().detach().requires_grad_(True) for t in params]
                yr = reference(xr, *pr, stride)
                gr = torch.autograd.grad(yr, [xr] + pr, gy.float())
                names = ["dx", "d_lambda", "d_f_out", "d_f_in", "d_kh", "d_kw"]
                print(json.dumps({"case": args.case, "batch": B, "err_y": rel(y, yr), **{n: rel(a, b) for n, a, b in zip(names, grads, gr)}}),
                      flush=True)
            else:
                ws = cprows_ops._fwd_op(x.permute(0, 2, 3, 1), lam, f_out, f_in, kh, kw, None, 1)[1]
                xr_, gr_ = x.permute(0, 2, 3, 1), gy.permute(0, 2, 3, 1)
                ms = timed(lambda: cprows_ops._bwd_op(xr_, gr_, ws, lam, kh, kw, o), args.steps)
                print(json.dumps({"case": args.case, "batch": B, "ms": ms, "alg_GBps": 3 * x.numel() * 2 / ms / 1e6}), flush=True)


if __name__ == "__main__":
    main()
