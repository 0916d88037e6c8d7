"""Micro-benchmark of tlt_gemm_bf16_tc on the hot-path GEMM shapes, beside torch.matmul (cuBLAS) on the same box.

Development aid (run on the GPU box):  python tools/gemm_bench.py [--quick]
Prints one line per (shape, variant): time, TFLOP/s, max relative error vs cuBLAS.
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from torch_b200 import _cuda

dev = torch.device("cuda:0")
lib = _cuda.lib()

# (M, N, K, a_mn, b_mn, out_dtype, split_k)   -- config 2 (BERT FFN, 32768 tokens) fwd / dgrad / wgrad, config 5, square
SHAPES = [
    (32768, 3072, 768, 0, 0, torch.bfloat16, 0),    # h = x W_up^T
    (32768, 768, 3072, 0, 0, torch.bfloat16, 0),    # y = h W_down^T
    (32768, 3072, 768, 0, 1, torch.bfloat16, 0),    # dh = dy W_down
    (32768, 768, 3072, 0, 1, torch.bfloat16, 0),    # dx = dh W_up
    (768, 3072, 32768, 1, 1, torch.bfloat16, 0),    # dW_down = dy^T h   (bf16 out: no split)
    (768, 3072, 32768, 1, 1, torch.float32, 0),     # same, fp32 out: library-chosen split-K
    (768, 3072, 32768, 1, 1, torch.float32, 1),
    (768, 3072, 32768, 1, 1, torch.float32, 2),
    (768, 3072, 32768, 1, 1, torch.float32, 4),
    (3072, 768, 32768, 1, 1, torch.float32, 0),     # dW_up = dh^T x
    (8192, 4096, 4096, 0, 0, torch.bfloat16, 0),    # config 5 dense-equivalent
    (8192, 8192, 8192, 0, 0, torch.bfloat16, 0),    # the shape MEASURED_PEAKS.json quotes cuBLAS on
]


def run(M, N, K, a_mn, b_mn, out_dtype, split, iters=20):
    a = torch.randn((K, M) if a_mn else (M, K), device=dev, dtype=torch.bfloat16)
    b = torch.randn((K, N) if b_mn else (N, K), device=dev, dtype=torch.bfloat16)
    c = torch.empty((M, N), device=dev, dtype=out_dtype)
    g = _cuda.GemmDesc()
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldc = a.shape[1], b.shape[1], N
    g.a_major, g.b_major = a_mn, b_mn
    g.dtype_c = _cuda.dtype_code(out_dtype)
    g.alpha = 1.0
    g.split_k = split

    def ours():
        _cuda.check(lib.tlt_gemm_bf16_tc(C.byref(g), _cuda.ptr(a), _cuda.ptr(b), None, _cuda.ptr(c), _cuda.stream_ptr(c)), "gemm")

    A = a.t() if a_mn else a
    Bt = b if b_mn else b.t()

    def cublas():
        return torch.matmul(A, Bt)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}
    for name, fn in (("ours", ours), ("cublas", cublas)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            flush.zero_()                      # L2 flush between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        res[name] = tot / iters
    ref = cublas().float()
    err = ((c.float() - ref).norm() / ref.norm()).item()
    fl = 2.0 * M * N * K
    print(f"M{M:6d} N{N:5d} K{K:6d} a{a_mn}b{b_mn} {str(out_dtype)[6:]:9s} split={split} | ours {res['ours'] * 1e3:7.1f} us "
          f"{fl / res['ours'] / 1e9:7.1f} TF | cublas {res['cublas'] * 1e3:7.1f} us {fl / res['cublas'] / 1e9:7.1f} TF | "
          f"rel diff {err:.2e}", flush=True)


if __name__ == "__main__":
    
    shapes = SHAPES[:3] if "--quick" in sys.argv else SHAPES
    if "--chain" in sys.argv:      # the six GEMMs of every unfused CP-conv rows chain layer of config 3 (256 images)
        shapes = []
        r8 = lambda v: (v + 7) // 8 * 8
        for cin, cout, hw, s, r in [(64, 128, 56, 2, 93), (128, 128, 28, 1, 141), (128, 256, 28, 2, 189), (256, 256, 14, 1, 285),
                                    (256, 512, 14, 2, 381), (512, 512, 7, 1, 573)]:
            pin, pout, ld = 256 * hw * hw, 256 * (hw // s) ** 2, r8(r)
            bf, f32 = torch.bfloat16, torch.float32
            print(f"--- {cin}->{cout} @{hw}^2 s{s} R={r}")
            for sh in [(pin, ld, cin, 0, 0, bf, 0), (pout, cout, ld, 0, 0, bf, 0), (pout, ld, cout, 0, 1, bf, 0), (pin, cin, ld, 0, 1, bf, 0),
                       (cout, ld, pout, 1, 1, f32, 0), (cin, ld, pin, 1, 1, f32, 0)]:
                run(*sh, iters=10)
        sys.exit(0)
    for s in shapes:
        run(*s)
