"""Time the fused CP-convolution kernels alone (CUDA events, tensors larger than L2): forward chain with and without the
intermediate saves, and -- when built -- the one-pass backward.  Prints one JSON line per case with algorithmic GB/s."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_b200 import cpconv_ops  # noqa: E402

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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--bwd", action="store_true")
    args = ap.parse_args()
    B = args.batch
    for (ch, hw, R) in [(64, 56, 69), (128, 28, 141)]:
        torch.manual_seed(0)
        x = torch.randn(B, ch, hw, hw, device=dev, dtype=torch.bfloat16)
        f_in = (torch.randn(ch, R, device=dev) * 0.1).bfloat16()
        f_out = (torch.randn(ch, R, device=dev) * 0.1).bfloat16()
        taps = (torch.randn(9, R, device=dev) * 0.3).bfloat16()
        nbytes = 2 * x.numel() * 2
        for variant in ([0, 1] if ch == 64 else [0]):
            for save in (False, True):
                ms = timed(lambda: cpconv_ops._cpf_op(x, f_in, taps, f_out, None, False, save, variant), args.steps)
                extra = 2 * B * R * hw * hw * 2 if save else 0
                print(json.dumps({"case": f"cpconv fused fwd {ch}@{hw}^2 R={R} B={B} variant={variant} save={save}", "ms": ms,
                                  "alg_GBps": nbytes / ms / 1e6, "moved_GBps": (nbytes + extra) / ms / 1e6}), flush=True)
        if args.bwd and hasattr(cpconv_ops, "_cpf_bwd_op"):
            gy = torch.randn_like(x)
            ms = timed(lambda: cpconv_ops._cpf_bwd_op(x, gy, f_in, taps, f_out), args.steps)
            print(json.dumps({"case": f"cpconv fused bwd {ch}@{hw}^2 R={R} B={B}", "ms": ms, "alg_GBps": 3 * x.numel() * 2 / ms / 1e6}),
                  flush=True)


if __name__ == "__main__":
    main()
