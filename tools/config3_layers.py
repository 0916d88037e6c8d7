"""Config 3 layer by layer: fwd+bwd of each distinct ResNet-18 CP conv (channels-last, bf16, 256 images) timed as a CUDA graph,
plus the CUDA-event time of every C entry point of one eager step (torch_b200._cuda.PROFILE).

    python tools/config3_layers.py [--batch 256] [--only 1..7] > gpurun_out/x/layers.json
"""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import torch_b200 as tlt
import bench

SHAPES = [(64, 64, 56, 1, 4), (64, 128, 56, 2, 1), (128, 128, 28, 1, 3), (128, 256, 28, 2, 1), (256, 256, 14, 1, 3), (256, 512, 14, 2, 1),
          (512, 512, 7, 1, 3)]
HBM = bench._peaks()[0].get("hbm_gbs", 6533.5)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--only", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B = args.batch
    total = 0.0
    for i, (cin, cout, hw, s, count) in enumerate(SHAPES):
        if args.only and args.only != i + 1:
            continue
        torch.manual_seed(1)
        conv = tlt.FactorizedConv(cin, cout, 3, order=2, padding=1, stride=s, factorization="cp", rank=0.25, bias=False, device=dev)
        conv.reset_parameters(std=1.0 / math.sqrt(9 * cin))
        conv = conv.to(torch.bfloat16)
        x = torch.randn(B, cin, hw, hw, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        ho = hw // s
        gy = torch.randn(B, cout, ho, ho, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
        params = list(conv.parameters())
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def step():
            y = conv(x)
            torch.autograd.grad(y, [x] + params, gy)

        def fwd_only():
            with torch.no_grad():
                conv(x)

        out = {"layer": f"{cin}->{cout} @{hw}^2 s{s}", "rank": int(conv.weight.rank), "count_in_config3": count}
        for name, fn in (("fwd_bwd", step), ("fwd", fwd_only)):
            run, mode, launches, graphed = bench.time_graph(fn, args.steps, 3)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t = 0.0
            for _ in range(args.steps):
                flush.zero_()                     # cold L2, as inside the 16-layer chain (2 GB of activations per step)
                e0.record()
                run()
                e1.record()
                torch.cuda.synchronize()
                t += e0.elapsed_time(e1)
            out[name + "_ms"] = t / args.steps
            out[name + "_launches"] = launches
            del run, graphed
        nx, ny = B * cin * hw * hw, B * cout * ho * ho
        out["hbm_floor_ms"] = 2 * (3 * nx + 2 * ny) / (HBM * 1e9) * 1e3
        out["frac"] = out["hbm_floor_ms"] / out["fwd_bwd_ms"]
        out["entry_points_us"] = {k: [round(v["ms_per_step"] * 1e3, 1), v["launches_per_step"]] for k, v in bench.profile_entry_points(step).items()}
        total += count * out["fwd_bwd_ms"]
        print(json.dumps(out), flush=True)
        del conv, x, gy, flush
        torch.cuda.empty_cache()
    print(json.dumps({"config3_sum_of_layers_ms": total}), flush=True)


if __name__ == "__main__":
    main()
