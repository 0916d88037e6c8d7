"""Config 3 what-if timings: the headline step (16 chained CP convs, one CUDA graph) with one family of launches removed at a time, to
see what each family costs INSIDE the graph (overlap included) rather than as a serialised ncu / CUDA-event sum.  Results of the
reduced steps are garbage by construction; only the times mean anything.

    python tools/whatif_config3.py [--images 256] [--steps 20] > gpurun_out/x/whatif.json
"""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import torch_b200 as tlt
from torch_b200 import rows_ops, blocktt_ops, cprows_ops
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
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

    orig = {"prep": rows_ops._execute_chain_prep, "post": rows_ops._execute_chain_post, "dwsep": rows_ops._execute_dwsep,
            "gemm": blocktt_ops._execute_gemm, "pack": cprows_ops._execute_pack, "fin": cprows_ops._execute_finalize,
            "ffwd": cprows_ops._execute_fwd, "fbwd": cprows_ops._execute_bwd}

    def restore():
        rows_ops._execute_chain_prep, rows_ops._execute_chain_post, rows_ops._execute_dwsep = orig["prep"], orig["post"], orig["dwsep"]
        blocktt_ops._execute_gemm = orig["gemm"]
        cprows_ops._execute_pack, cprows_ops._execute_finalize = orig["pack"], orig["fin"]
        cprows_ops._execute_fwd, cprows_ops._execute_bwd = orig["ffwd"], orig["fbwd"]

    nop = lambda *a, **k: None

    def sweeps_without(kinds):
        def f(kind, *a, **k):
            if kind not in kinds:
                orig["dwsep"](kind, *a, **k)
        return f

    def gemm_without(wgrad):
        def f(A, Bm, bias, Cm, **k):
            is_w = k.get("a_major", 0) == 1 and k.get("b_major", 0) == 1
            if is_w != wgrad:
                orig["gemm"](A, Bm, bias, Cm, **k)
        return f

    variants = {
        "base": lambda: None,
        "no_prep_pack": lambda: (setattr(rows_ops, "_execute_chain_prep", nop), setattr(cprows_ops, "_execute_pack", nop)),
        "no_post_finalize": lambda: (setattr(rows_ops, "_execute_chain_post", nop), setattr(cprows_ops, "_execute_finalize", nop)),
        "no_wgrad_branch": lambda: (setattr(rows_ops, "_execute_dwsep", sweeps_without({"wgrad"})),
                                    setattr(blocktt_ops, "_execute_gemm", gemm_without(True)),
                                    setattr(rows_ops, "_execute_chain_post", nop)),
        "no_wgrad_sweep": lambda: setattr(rows_ops, "_execute_dwsep", sweeps_without({"wgrad"})),
        "no_wgrad_gemms": lambda: setattr(blocktt_ops, "_execute_gemm", gemm_without(True)),
        "no_fwd_dgrad_sweeps": lambda: setattr(rows_ops, "_execute_dwsep", sweeps_without({"fwd", "dgrad"})),
        "no_main_gemms": lambda: setattr(blocktt_ops, "_execute_gemm", gemm_without(False)),
        "no_fused_fwd": lambda: setattr(cprows_ops, "_execute_fwd", nop),
        "no_fused_bwd": lambda: setattr(cprows_ops, "_execute_bwd", nop),
        "no_chain_branches": lambda: setattr(rows_ops, "_BRANCHES", False),
    }
    only = set(args.only.split(",")) if args.only else None
    base = None
    for name, apply in variants.items():
        if only and name not in only and name != "base":
            continue
        restore()
        rows_ops._BRANCHES = True
        apply()
        run, mode, launches, graphed = bench.time_graph(step, args.steps, 3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        if name == "base":
            base = ms
        print(json.dumps({"variant": name, "ms_per_step": round(ms, 4), "delta_ms": round(ms - base, 4), "launches": launches, "mode": mode}),
              flush=True)
        del run, graphed
    restore()


if __name__ == "__main__":
    main()
