"""Accuracy experiment: fp32 GEMM as ONE tcgen05 bf16 GEMM on 3-way bf16-split operands (6 K-blocks), fp32 accumulation in TMEM."""
import sys
sys.path.insert(0, "/root/repo")
import torch
from torch_b200.blocktt_ops import _execute_gemm
dev = torch.device("cuda:0")
torch.manual_seed(0)

def split3(t):
    hi = t.to(torch.bfloat16); r = t - hi.float()
    mid = r.to(torch.bfloat16); r = r - mid.float()
    lo = r.to(torch.bfloat16)
    return hi, mid, lo

def catA(t, dim):
    h, m, l = split3(t); return torch.cat([h, h, m, h, m, l], dim).contiguous()
def catB(t, dim):
    h, m, l = split3(t); return torch.cat([h, m, h, l, m, h], dim).contiguous()

for (M, N, K) in [(128, 2344, 512), (128, 512, 2344), (512, 2344, 128)]:
    a = torch.randn(M, K, device=dev) * 0.5
    b = torch.randn(N, K, device=dev) * 0.5
    A6, B6 = catA(a, 1), catB(b, 1)
    out = torch.empty(M, N, device=dev, dtype=torch.float32)
    _execute_gemm(A6, B6, None, out, M=M, N=N, K=6 * K, lda=6 * K, ldb=6 * K, a_major=0, b_major=0)
    ref = a.double() @ b.double().t()
    e = ((out.double() - ref).norm() / ref.norm()).item()
    e32 = (((a @ b.t()).double() - ref).norm() / ref.norm()).item()
    # 3 blocks only
    A3 = torch.cat([split3(a)[0], split3(a)[0], split3(a)[1]], 1).contiguous(); B3 = torch.cat([split3(b)[0], split3(b)[1], split3(b)[0]], 1).contiguous()
    out3 = torch.empty(M, N, device=dev, dtype=torch.float32)
    _execute_gemm(A3, B3, None, out3, M=M, N=N, K=3 * K, lda=3 * K, ldb=3 * K, a_major=0, b_major=0)
    e3 = ((out3.double() - ref).norm() / ref.norm()).item()
    print(f"M{M} N{N} K{K}: 6-block split {e:.2e}   3-block {e3:.2e}   torch fp32 matmul {e32:.2e}")
# MN-major operands (weight-gradient form): out[M,N] = sum_k A[k,M] B[k,N]
M, N, K = 512, 2344, 128
a = torch.randn(K, M, device=dev) * 0.5; b = torch.randn(K, N, device=dev) * 0.5
A6, B6 = catA(a, 0), catB(b, 0)
out = torch.empty(M, N, device=dev, dtype=torch.float32)
_execute_gemm(A6, B6, None, out, M=M, N=N, K=6 * K, lda=M, ldb=N, a_major=1, b_major=1)
ref = a.double().t() @ b.double()
print("MN-major 6-block", ((out.double() - ref).norm() / ref.norm()).item())
