import sys, os
sys.path.insert(0, "/root/repo")
import torch, ctypes as C
from torch_b200 import _cuda
lib = _cuda.lib()
dev = torch.device("cuda:0")
def run(M, N, K, a_mn, b_mn, out_dtype):
    a = torch.randn((K, M) if a_mn else (M, K), device=dev, dtype=torch.bfloat16)
    b = torch.randn((K, N) if b_mn else (N, K), device=dev, dtype=torch.bfloat16)
    c = torch.empty((M, N), device=dev, dtype=out_dtype)
    g = _cuda.GemmDesc()
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldc = a.shape[1], b.shape[1], N
    g.a_major, g.b_major = a_mn, b_mn
    g.dtype_c = _cuda.dtype_code(out_dtype)
    g.alpha = 1.0
    for _ in range(3):
        _cuda.check(lib.tlt_gemm_bf16_tc(C.byref(g), _cuda.ptr(a), _cuda.ptr(b), None, _cuda.ptr(c), _cuda.stream_ptr(c)), "gemm")
    torch.cuda.synchronize()
run(200704, 144, 128, 0, 0, torch.bfloat16)
run(200704, 128, 144, 0, 0, torch.bfloat16)
run(128, 144, 200704, 1, 1, torch.float32)
