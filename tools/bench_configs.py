"""Per-config measurement of the factorized-layer hot path (BASELINE.json configs 1, 3, 4, 5; config 2 is bench.py).

    python tools/bench_configs.py [--configs 1,3,4,5] [--out gpurun_out/configs.json]

For every config: fwd+bwd through the public API (nn.Module layers), CUDA-graph replayed when capturable, timed with CUDA
events over `--steps` replays after warm-up; reports ms/step, samples/s and the fraction of the roofline SURVEY.md 8(d)
assigns (algorithmic bytes or FLOPs from that table / measured peak in MEASURED_PEAKS.json).  One JSON object per
config on stdout and in --out.  Numbers here are development measurements, not the driver's bench line.
"""
import argparse
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import torch_b200 as tlt
from torch_b200 import ops
from torch_b200.graphs import GraphedStep

dev = torch.device("cuda:0")
REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
try:
    PEAKS = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
except Exception:
    PEAKS = {}
HBM = PEAKS.get("hbm_gbs", 6650.0)
TENSOR = PEAKS.get("bf16_tflops_sustained", 1400.0)
FFMA_TF = 148 * 128 * 2 * 1.965e9 / 1e12        # fp32 FFMA peak (SURVEY 7 "hard parts")


def time_step(step, steps, warmup, graph=True):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    mode = "eager"
    run = step
    if graph:
        try:
            run = GraphedStep(step).replay
            mode = "cuda_graph"
        except Exception as exc:
            torch.cuda.synchronize()
            print(f"  (graph capture failed: {exc!r}; eager)", file=sys.stderr)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    l0 = ops.launch_count()
    step()
    launches = ops.launch_count() - l0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, mode, launches


def fwd_bwd(layers, x, gy_shape_fn=None):
    params = [p for l in layers for p in l.parameters()]
    holder = {}

    def step():
        x.grad = None
        for p in params:
            p.grad = None
        y = x
        for l in layers:
            y = l(y)
        if "gy" not in holder:
            holder["gy"] = torch.randn_like(y)
        y.backward(holder["gy"])
    return step


def report(name, ms, mode, launches, samples, bound, alg_bytes=None, alg_flops=None, extra=None):
    out = {"config": name, "ms_per_step": ms, "samples_per_s": samples / (ms / 1e3), "step_mode": mode, "tlt_launches_per_step": launches,
           "bound": bound}
    if alg_bytes is not None:
        out["algorithmic_bytes"] = alg_bytes
        out["hbm_gbs_achieved"] = alg_bytes / (ms / 1e3) / 1e9
        out["hbm_frac"] = out["hbm_gbs_achieved"] / HBM
    if alg_flops is not None:
        out["algorithmic_flops"] = alg_flops
        out["tflops_achieved"] = alg_flops / (ms / 1e3) / 1e12
        out["tensor_frac"] = out["tflops_achieved"] / TENSOR
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)
    return out


def config1(args):
    """FactorizedLinear 512->512 CP rank 0.5, batch 128, fp32 (latency / FFMA-bound)."""
    torch.manual_seed(0)
    layer = tlt.FactorizedLinear((4, 8, 16), (4, 8, 16), bias=True, factorization="cp", rank=0.5, device=dev, dtype=torch.float32)
    x = torch.randn(128, 512, device=dev, requires_grad=True)
    ms, mode, n = time_step(fwd_bwd([layer], x), args.steps, args.warmup)
    flops = 1.84e9
    return report("1: FactorizedLinear 512->512 cp rank=0.5 B=128 fp32", ms, mode, n, 128, "latency/FFMA", alg_bytes=2.9e6,
                  alg_flops=flops, extra={"ffma_frac": flops / (ms / 1e3) / 1e12 / FFMA_TF, "rank": int(layer.weight.rank)})


RESNET18 = [(64, 64, 56, 1)] * 4 + [(64, 128, 56, 2)] + [(128, 128, 28, 1)] * 3 + [(128, 256, 28, 2)] + [(256, 256, 14, 1)] * 3 + \
           [(256, 512, 14, 2)] + [(512, 512, 7, 1)] * 3


def config3(args):
    """CP separable 3x3 FactorizedConv (rank 0.25) for each of ResNet-18's 16 3x3 convs, batch 256, bf16: per distinct layer."""
    res = []
    total_ms, total_bytes, total_flops = 0.0, 0.0, 0.0
    seen = {}
    B = args.conv_batch
    layers = RESNET18 if not args.only_layer else [RESNET18[i] for i in (0, 4, 5, 8, 9, 12, 13)][args.only_layer - 1:args.only_layer]
    for (cin, cout, hw, stride) in layers:
        key = (cin, cout, hw, stride)
        if key not in seen:
            torch.manual_seed(1)
            conv = tlt.FactorizedConv(cin, cout, 3, order=2, padding=1, stride=stride, factorization="cp", rank=0.25, bias=False,
                                      device=dev, dtype=torch.float32)
            conv.reset_parameters()                       # like the reference, the constructor leaves the factors uninitialised
            conv = conv.to(torch.bfloat16)
            x = torch.randn(B, cin, hw, hw, device=dev, dtype=torch.bfloat16).requires_grad_(True)
            ms, mode, n = time_step(fwd_bwd([conv], x), max(args.steps // 4, 5), 3)
            R = int(conv.weight.rank)
            ho = hw // stride
            nx, ny = B * cin * hw * hw, B * cout * ho * ho
            P = R * (cin + cout + 6 + 1)
            byts = 2 * (3 * nx + 2 * ny) + 3 * 2 * P
            # fwd MACs: 1x1 in (at input res), two depthwise 3-tap (first strided along H... both counted at their output res), 1x1 out
            macs = B * (hw * hw * cin * R + ho * hw * R * 3 + ho * ho * R * 3 + ho * ho * R * cout)
            flops = 3 * 2 * macs
            seen[key] = (ms, byts, flops, R, mode, n)
            report(f"3-layer: cp conv {cin}->{cout} @{hw}^2 s{stride} R={R} B={B} bf16", ms, mode, n, B,
                   "hbm" if flops / byts < 216 else "tensor", alg_bytes=byts, alg_flops=flops)
            del conv, x
            torch.cuda.empty_cache()
        ms, byts, flops, R, mode, n = seen[key]
        total_ms += ms
        total_bytes += byts
        total_flops += flops
    return report(f"3: ResNet-18 16x cp conv3x3 rank 0.25 B={B} bf16 (sum over layers, each timed alone)", total_ms, "cuda_graph", 0, B, "hbm",
                  alg_bytes=total_bytes, alg_flops=total_flops)


def config4(args):
    """Tucker FactorizedConv 512->512 3x3 (rank 0.5) with tensor dropout + Tucker TRL head on (512,2048,7,7), bf16."""
    out = []
    torch.manual_seed(2)
    B = 512
    conv = tlt.FactorizedConv(512, 512, 3, order=2, padding=1, factorization="tucker", rank=0.5, bias=True, device=dev,
                              dtype=torch.float32)
    conv.reset_parameters()
    conv = conv.to(torch.bfloat16)
    x = torch.randn(B, 512, 7, 7, device=dev, dtype=torch.bfloat16).requires_grad_(True)
    ms, mode, n = time_step(fwd_bwd([conv], x), args.steps, args.warmup)
    out.append(report(f"4-conv: tucker conv 512->512 3x3 @7^2 rank={tuple(conv.weight.rank)} B=512 bf16 (no dropout)", ms, mode, n, B, "tensor",
                      alg_bytes=134.4e6, alg_flops=263.8e9))
    conv.weight = tlt.tensor_dropout(conv.weight, p=0.1)
    conv.train()
    ms, mode, n = time_step(fwd_bwd([conv], x), args.steps, args.warmup)     # fixed-shape masks: no host sync, graph-capturable
    out.append(report("4-conv+dropout: same with tensor_dropout(p=0.1) (masks drawn on device with the reference's RNG calls, multiplied into the core)", ms, mode, n, B,
                      "tensor", alg_bytes=134.4e6, alg_flops=263.8e9))
    del conv, x
    torch.cuda.empty_cache()
    trl = tlt.TRL((2048, 7, 7), (1000,), bias=True, factorization="tucker", rank=(64, 4, 4, 10), device=dev, dtype=torch.float32)
    with torch.no_grad():
        for p in trl.parameters():
            p.mul_(0.05)
    trl = trl.to(torch.bfloat16)
    x = torch.randn(B, 2048, 7, 7, device=dev, dtype=torch.bfloat16).requires_grad_(True)
    ms, mode, n = time_step(fwd_bwd([trl], x), args.steps, args.warmup)
    out.append(report("4-TRL: TRL (2048,7,7)->1000 tucker rank (64,4,4,10) B=512 bf16 (layer default path)", ms, mode, n, B, "hbm",
                      alg_bytes=311.2e6, alg_flops=23e9))
    return out


def config5(args):
    """Large-rank Tucker / BlockTT FactorizedLinear 4096->4096 (rank 0.75), 8192 tokens, bf16."""
    out = []
    B = 8192
    for fact, flops, byts in (("tucker", 576.6e9, 403.9e6), ("blocktt", 824.6e9, 411.2e6)):
        torch.manual_seed(3)
        layer = tlt.FactorizedLinear((16, 16, 16), (16, 16, 16), bias=True, factorization=fact, rank=0.75, device=dev,
                                     dtype=torch.float32).to(torch.bfloat16)
        x = torch.randn(B, 4096, device=dev, dtype=torch.bfloat16).requires_grad_(True)
        ms, mode, n = time_step(fwd_bwd([layer], x), max(args.steps // 2, 5), 3)
        out.append(report(f"5-{fact}: FactorizedLinear 4096->4096 {fact} rank 0.75 -> {tuple(layer.weight.rank)} B=8192 bf16", ms, mode, n, B,
                          "tensor", alg_bytes=byts, alg_flops=flops))
        del layer, x
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,3,4,5")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--conv-batch", type=int, default=256)
    ap.add_argument("--only-layer", type=int, default=0, help="config 3: time only the n-th distinct layer (1..7)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    fns = {"1": config1, "3": config3, "4": config4, "5": config5}
    results = []
    for c in args.configs.split(","):
        t0 = time.time()
        try:
            r = fns[c](args)
            results += r if isinstance(r, list) else [r]
        except Exception as exc:
            import traceback
            traceback.print_exc()
            results.append({"config": c, "error": repr(exc)})
            print(json.dumps(results[-1]), flush=True)
            torch.cuda.synchronize()
        print(f"  [config {c}: {time.time() - t0:.1f} s]", file=sys.stderr, flush=True)
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump({"peaks": {"hbm_gbs": HBM, "bf16_tflops_sustained": TENSOR, "ffma_tflops": FFMA_TF}, "results": results},
                  open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
