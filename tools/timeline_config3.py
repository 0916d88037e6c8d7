"""Kernel timeline of the headline step INSIDE its CUDA graph (CUPTI through torch.profiler; nsys is not in the image): start,
duration and stream of every kernel of one replay, so that gaps between dependent launches and the real overlap of the two
branches are visible (an ncu launch list serialises kernels and flushes caches; CUDA events around eager launches include launch
latency).  Not a bench number: the profiler adds its own overhead.

    python tools/timeline_config3.py [--images 256] > gpurun_out/x/timeline.json
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--config", default="3", help="3 (headline) or a key of bench.other_configs: 1, 2, 4-conv, 4-conv+dropout, 4-TRL, 5-tucker, 5-blocktt")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    if args.config != "3":
        c = bench.other_configs(dev)[args.config]()
        step = bench._fwd_bwd_step(c["layers"], c["x"])[0]
        return timeline(step)
    torch.manual_seed(1234)
    convs = []
    for cin, cout, hw, s in bench.RESNET18:
        conv = tlt.FactorizedConv(cin, cout, 3, order=2, padding=1, stride=s, factorization="cp", rank=0.25, bias=False, device=dev)
        conv.reset_parameters(std=1.0 / math.sqrt(9 * cin))
        convs.append(conv.to(torch.bfloat16))
    params = [p for c in convs for p in c.parameters()]
    B = args.images
    x = torch.randn(B, 64, 56, 56, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    gy = torch.randn(B, 512, 7, 7, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)

    def step():
        x.grad = None
        for p in params:
            p.grad = None
        a = x
        for c in convs:
            a = c(a)
        a.backward(gy)

    timeline(step)


def timeline(step):
    run, mode, launches, graphed = bench.time_graph(step, 5, 3)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            run()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower() and "memset" not in e.name.lower()]
    ev.sort(key=lambda e: e.time_range.start)
    n = len(ev) // 3
    ev = ev[2 * n:]                       # the last replay
    t0 = ev[0].time_range.start
    rows = []
    for e in ev:
        rows.append({"name": e.name[:70], "start_us": round(e.time_range.start - t0, 2), "dur_us": round(e.time_range.end - e.time_range.start, 2),
                     "stream": getattr(e, "device_resource_id", None)})
    end = max(r["start_us"] + r["dur_us"] for r in rows)
    print(json.dumps({"mode": mode, "launches_per_step": launches, "kernels_in_replay": len(rows), "span_us": end,
                      "sum_dur_us": sum(r["dur_us"] for r in rows)}))
    for r in rows:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
