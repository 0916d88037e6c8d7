"""Times tlt_mmd16_fwd / tlt_mmd16_bwd alone at BASELINE config 5's shapes (16^3 <-> 15^3, 8192 tokens) and reports the
achieved HBM GB/s against algorithmic bytes (read x + write y; backward: read u, g + write du).

    python tools/mmd16_bench.py [--batch 8192] [--iters 20]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from torch_b200 import tucker_ops as to

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8192)
ap.add_argument("--iters", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda:0")
B = args.batch
torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(args.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for name, din, dout, flags in (("project 16^3->15^3", [16] * 3, [15] * 3, [True] * 3), ("expand 15^3->16^3", [15] * 3, [16] * 3, [False] * 3)):
    n_in, n_out = din[0] ** 3, dout[0] ** 3
    ld_in, ld_out = (n_in + 7) & ~7, (n_out + 7) & ~7
    x = torch.randn(B, ld_in, device=dev, dtype=torch.bfloat16)
    g = torch.randn(B, ld_out, device=dev, dtype=torch.bfloat16)
    mats = [(torch.randn((din[k], dout[k]) if flags[k] else (dout[k], din[k]), device=dev) / 4).to(torch.bfloat16) for k in range(3)]
    y = torch.empty(B, ld_out, device=dev, dtype=torch.bfloat16)
    du = torch.empty_like(x)
    dM = torch.zeros(3, 16, 16, device=dev)
    t_f = timed(lambda: to._execute_mmd16_fwd(x, din, dout, mats, flags, None, y))
    t_b = timed(lambda: to._execute_mmd16_bwd(x, g, din, dout, mats, flags, du, dM))
    t_d = timed(lambda: to._execute_mmd16_bwd(x, g, din, dout, mats, flags, du, None))
    bf = 2.0 * B * (n_in + n_out)
    bb = 2.0 * B * (2 * n_in + n_out)
    print(json.dumps({"case": name, "batch": B, "fwd_us": t_f * 1e3, "fwd_gbs": bf / t_f / 1e6, "bwd_us": t_b * 1e3,
                      "bwd_gbs": bb / t_b / 1e6, "bwd_data_only_us": t_d * 1e3, "bwd_data_only_gbs": bf / t_d / 1e6}), flush=True)
