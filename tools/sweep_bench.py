"""Separable depthwise sweeps of config 3's unfused layers, one launch at a time: forward, data gradient, weight gradient of every
layer shape, under the default plan of csrc/dwsep_staged.cu and under the plans named on the command line
(TLT_DWST_PLAN="band rows,ring stages,min segment length").  CUDA events around 20 launches, L2 flushed between launches.

    python tools/sweep_bench.py [--plans 4,2,7 2,3,3 ...] > gpurun_out/x/sweeps.json
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from torch_b200 import rows_ops

SHAPES = [(56, 96, 2), (28, 144, 1), (28, 192, 2), (14, 288, 1), (14, 384, 2), (7, 576, 1)]     # (H = W of the input map, ld, stride)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plans", nargs="*", default=[])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--kinds", default="fwd,dgrad,wgrad")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B = args.batch
    for hw, ld, s in SHAPES:
        ho = hw // s
        z = torch.randn(B * hw * hw, ld, device=dev, dtype=torch.bfloat16)
        g = torch.randn(B * ho * ho, ld, device=dev, dtype=torch.bfloat16)
        taps = torch.randn(6, ld, device=dev, dtype=torch.float32)
        y = torch.empty_like(g)
        dz = torch.empty_like(z)
        acc = torch.empty(rows_ops._dwsep_parts(), 9, ld, device=dev, dtype=torch.float32)
        calls = {"fwd": lambda: rows_ops._execute_dwsep("fwd", z, taps, y, B, hw, hw, s),
                 "dgrad": lambda: rows_ops._execute_dwsep("dgrad", g, taps, dz, B, hw, hw, s),
                 "wgrad": lambda: rows_ops._execute_dwsep("wgrad", z, g, acc, B, hw, hw, s)}
        dw = torch.zeros(9, ld, device=dev, dtype=torch.float32)
        calls["tile_wgrad"] = lambda: rows_ops._execute_dwconv_rows("wgrad", z, g, dw, B, [hw, hw], [ho, ho], [3, 3], [s, s], [1, 1], [1, 1])
        calls["tile_fwd"] = lambda: rows_ops._execute_dwconv_rows("fwd", z, torch.randn(9, ld, device=dev, dtype=torch.bfloat16), y, B, [hw, hw], [ho, ho], [3, 3], [s, s], [1, 1], [1, 1])
        mb = {"tile_wgrad": (z.numel() + g.numel()) * 2 / 1e6, "tile_fwd": (z.numel() + g.numel()) * 2 / 1e6, "fwd": (z.numel() + g.numel()) * 2 / 1e6, "dgrad": (z.numel() + g.numel()) * 2 / 1e6, "wgrad": (z.numel() + g.numel()) * 2 / 1e6}
        ref = {}
        for plan in [""] + args.plans:
            if plan:
                os.environ["TLT_DWST_PLAN"] = plan
            else:
                os.environ.pop("TLT_DWST_PLAN", None)
            out = {"shape": f"{hw}^2 ld {ld} s{s}", "plan": plan or "default"}
            for kind in args.kinds.split(","):
                fn = calls[kind]
                fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t = 0.0
                for _ in range(args.reps):
                    flush.zero_()
                    e0.record()
                    fn()
                    e1.record()
                    torch.cuda.synchronize()
                    t += e0.elapsed_time(e1)
                us = t / args.reps * 1e3
                out[kind + "_us"] = round(us, 1)
                out[kind + "_gbs"] = round(mb[kind] / us * 1e3)
                if kind == "tile_wgrad":
                    dw.zero_(); fn(); torch.cuda.synchronize()
                    out["tile_wgrad_maxdiff_vs_staged"] = float((dw - ref["wgrad"]).abs().max() / ref["wgrad"].abs().max()) if "wgrad" in ref else None
                if kind == "wgrad":
                    if not plan:
                        ref["wgrad"] = acc.sum(0).clone()
                    else:
                        out["wgrad_maxdiff_vs_default"] = float((acc.sum(0) - ref["wgrad"]).abs().max() / ref["wgrad"].abs().max())
                if kind == "fwd":
                    if not plan:
                        ref["fwd"] = y.clone()
                    else:
                        out["fwd_equal_default"] = bool(torch.equal(y, ref["fwd"]))
            print(json.dumps(out), flush=True)
        os.environ.pop("TLT_DWST_PLAN", None)
        del z, g, y, dz, acc


if __name__ == "__main__":
    main()
