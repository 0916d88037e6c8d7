"""Custom ops (with autograd) for the fused 3-core BlockTT (TT-matrix) path -- BASELINE config 2's hot loop.

  tltorch::blocktt3_to_matrix   cores -> dense (rows, cols) matrix in ONE launch (tlt_blocktt3_reconstruct); backward
                                contracts the matrix gradient back onto the cores (tlt_blocktt3_core_grads, 2 launches)
  tltorch::blocktt3_linear      y = x . W^T (+ bias) with W rebuilt on the fly: reconstruct + tcgen05 GEMM (bias in the
                                epilogue).  Backward: dX = dY . W and dW = dY^T . X on the tensor cores -- dW as fp32
                                split-K partials met by TMA reduce-add -- then the fused core-gradient kernels; the bias
                                gradient is a column-sum mat-vec.

Reference: linear_blocktt (tltorch/functional/factorized_linear.py:39-60) and BlockTT.to_tensor
(tltorch/factorized_tensors/tensorized_matrices.py:354-384); why the dense route is taken at these shapes: DESIGN.md 4.
CPU tensors raise NotImplementedError (no fallback); the launch funnels below are what tests/_emulator.py swaps out.
"""
import ctypes as C
from typing import List, Optional, Tuple

import torch

from . import _cuda
from . import ops

__all__ = ["blocktt3_fits", "blocktt3_to_matrix", "blocktt3_linear", "gemm_supported"]

_SMEM_LIMIT = 200 * 1024


def blocktt3_fits(rows, cols, rank1, rank2, for_grads=True):
    """Mirror of fill_btt3() in csrc/blocktt.cu: do the staged core slices fit shared memory with the smallest chunk?"""
    r4 = lambda v: (v + 3) & ~3
    n_g3 = (rows[2] + 7) // 8
    g3 = (rows[2] + n_g3 - 1) // n_g3
    r2r, i3p = r4(rank2), r4(cols[2])
    i3s = cols[2] if cols[2] % 2 else cols[2] + 1
    need = r4(cols[0] * rank1) + r4(rank1 * cols[1] * rank2) + r4(r2r)
    need += (r4(g3 * cols[2] * r2r) + g3 * i3s) if for_grads else r4(rank2 * g3 * i3p)
    return need * 4 <= _SMEM_LIMIT and rows[0] * rows[1] < 2 ** 31 and n_g3 <= 65535


def gemm_supported(M, N, K, lda, ldb, ldc, out_dtype):
    """Mirror of tlt_gemm_bf16_tc_supported() for contiguous operands."""
    if min(M, N, K) <= 0 or lda % 8 or ldb % 8:
        return False
    return ldc % (4 if out_dtype == torch.float32 else 8) == 0


def _meta(c1, c2, c3):
    if c1.dim() != 4 or c2.dim() != 4 or c3.dim() != 4 or c1.shape[0] != 1 or c3.shape[3] != 1:
        raise ValueError("blocktt3: cores must be (1,r1,c1,a), (a,r2,c2,b), (b,r3,c3,1)")
    if c1.shape[3] != c2.shape[0] or c2.shape[3] != c3.shape[0]:
        raise ValueError("blocktt3: TT ranks of neighbouring cores do not match")
    rows = [c1.shape[1], c2.shape[1], c3.shape[1]]
    cols = [c1.shape[2], c2.shape[2], c3.shape[2]]
    return rows, cols, c1.shape[3], c2.shape[3]


def _desc(dtype, rows, cols, r1, r2):
    d = _cuda.BlockTT3Desc()
    d.dtype = _cuda.dtype_code(dtype)
    for k in range(3):
        d.rows[k] = rows[k]
        d.cols[k] = cols[k]
    d.rank1, d.rank2 = r1, r2
    return d


# =====================================================================================================================
# launch funnels (the only functions here that touch the shared library)
# =====================================================================================================================
def _execute_btt3_reconstruct(c1, c2, c3, W, rows, cols, r1, r2):
    _cuda.require_cuda(c1, c2, c3, W)
    d = _desc(W.dtype, rows, cols, r1, r2)
    _cuda.check(_cuda.lib().tlt_blocktt3_reconstruct(C.byref(d), _cuda.ptr(c1), _cuda.ptr(c2), _cuda.ptr(c3), _cuda.ptr(W),
                                                     _cuda.stream_ptr(W)), "tlt_blocktt3_reconstruct")


def _execute_btt3_core_grads(c1, c2, c3, dW32, d1, d2, d3, rows, cols, r1, r2):
    _cuda.require_cuda(c1, c2, c3, dW32, d1, d2, d3)
    d = _desc(c1.dtype, rows, cols, r1, r2)
    lib = _cuda.lib()
    n = lib.tlt_blocktt3_core_grads_workspace_size(C.byref(d))
    ws = torch.empty(n, dtype=torch.uint8, device=dW32.device)
    _cuda.check(lib.tlt_blocktt3_core_grads(C.byref(d), _cuda.ptr(c1), _cuda.ptr(c2), _cuda.ptr(c3), _cuda.ptr(dW32),
                                            _cuda.ptr(d1), _cuda.ptr(d2), _cuda.ptr(d3), _cuda.ptr(ws), n,
                                            _cuda.stream_ptr(dW32)), "tlt_blocktt3_core_grads")


def _execute_gemm(a, b, bias, out, M, N, K, lda, ldb, a_major, b_major, alpha=1.0, split_k=0):
    """out[M,N] = alpha * A[M,K] . B[N,K]^T (+ bias[N]) on the tcgen05 tensor cores; operand majors as in tlt_gemm_desc."""
    _cuda.require_cuda(a, b, bias, out)
    g = _cuda.GemmDesc()
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldc = lda, ldb, out.stride(0)
    g.a_major, g.b_major = a_major, b_major
    g.dtype_c = _cuda.dtype_code(out.dtype)
    g.accumulate = 0
    g.dtype_bias = _cuda.dtype_code(bias.dtype) if bias is not None else 0
    g.alpha = alpha
    g.split_k = split_k
    prof = ops.PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _cuda.check(_cuda.lib().tlt_gemm_bf16_tc(C.byref(g), _cuda.ptr(a), _cuda.ptr(b), _cuda.ptr(bias), _cuda.ptr(out),
                                             _cuda.stream_ptr(out)), "tlt_gemm_bf16_tc")
    if prof is not None:
        e1.record()
        prof.append((f"gemm_bf16_tc M{M} N{N} K{K} a{a_major}b{b_major} {str(out.dtype)[6:]}", e0, e1, 2.0 * M * N * K,
                     2.0 * (M * K + N * K) + out.element_size() * M * N))


# =====================================================================================================================
# tltorch::blocktt3_to_matrix
# =====================================================================================================================
@torch.library.custom_op("tltorch::blocktt3_to_matrix", mutates_args=())
def _to_matrix_op(c1: torch.Tensor, c2: torch.Tensor, c3: torch.Tensor) -> torch.Tensor:
    rows, cols, r1, r2 = _meta(c1, c2, c3)
    W = torch.empty((rows[0] * rows[1] * rows[2], cols[0] * cols[1] * cols[2]), dtype=c1.dtype, device=c1.device)
    _execute_btt3_reconstruct(c1.contiguous(), c2.contiguous(), c3.contiguous(), W, rows, cols, r1, r2)
    return W


@_to_matrix_op.register_fake
def _(c1, c2, c3):
    return c1.new_empty((c1.shape[1] * c2.shape[1] * c3.shape[1], c1.shape[2] * c2.shape[2] * c3.shape[2]))


@torch.library.custom_op("tltorch::blocktt3_core_grads", mutates_args=())
def _core_grads_op(c1: torch.Tensor, c2: torch.Tensor, c3: torch.Tensor,
                   gW32: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    rows, cols, r1, r2 = _meta(c1, c2, c3)
    d1, d2, d3 = torch.empty_like(c1), torch.empty_like(c2), torch.empty_like(c3)
    _execute_btt3_core_grads(c1.contiguous(), c2.contiguous(), c3.contiguous(), gW32.contiguous(), d1, d2, d3,
                             rows, cols, r1, r2)
    return d1, d2, d3


@_core_grads_op.register_fake
def _(c1, c2, c3, gW32):
    return torch.empty_like(c1), torch.empty_like(c2), torch.empty_like(c3)


def _to_matrix_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _to_matrix_backward(ctx, gW):
    c1, c2, c3 = ctx.saved_tensors
    # fp32 matrix gradient for the kernel (fp64 only ever reaches this point under the CPU descriptor emulator)
    d1, d2, d3 = _core_grads_op(c1, c2, c3, gW if gW.dtype in (torch.float32, torch.float64) else gW.float())
    return d1, d2, d3


_to_matrix_op.register_autograd(_to_matrix_backward, setup_context=_to_matrix_setup)


def blocktt3_to_matrix(c1, c2, c3):
    """(prod rows, prod cols) matrix of a 3-core BlockTT; differentiable w.r.t. the cores (first order)."""
    return _to_matrix_op(c1, c2, c3)


# =====================================================================================================================
# tltorch::blocktt3_linear
# =====================================================================================================================
def _gemm_args(transpose, which, B, R, Cn):
    """Operand majors of the three GEMMs for W (R x Cn) row-major.  transpose=True: y = x W^T (x has Cn features, y has R);
    transpose=False: y = x W (x has R features, y has Cn)."""
    if transpose:
        fin, fout = Cn, R
        if which == "fwd":      # y[B,fout] = x[B,fin] . W[fout,fin]^T
            return dict(M=B, N=fout, K=fin, lda=fin, ldb=fin, a_major=0, b_major=0)
        if which == "dx":       # dx[B,fin] = gy[B,fout] . (W^T)[fin,fout]^T ; element (n=i,k=o) at W[o*fin+i] -> MN-major
            return dict(M=B, N=fin, K=fout, lda=fout, ldb=fin, a_major=0, b_major=1)
        # dW[fout,fin] = gy^T . x ; A (m=o,k=b) at gy[b*fout+o], B (n=i,k=b) at x[b*fin+i]: both MN-major
        return dict(M=fout, N=fin, K=B, lda=fout, ldb=fin, a_major=1, b_major=1)
    fin, fout = R, Cn
    if which == "fwd":          # y[B,fout] = x[B,fin] . W[fin,fout] ; B (n=o,k=i) at W[i*fout+o] -> MN-major
        return dict(M=B, N=fout, K=fin, lda=fin, ldb=fout, a_major=0, b_major=1)
    if which == "dx":           # dx[B,fin] = gy[B,fout] . W[fin,fout]^T ; K-major
        return dict(M=B, N=fin, K=fout, lda=fout, ldb=fout, a_major=0, b_major=0)
    # dW[fin,fout] = x^T . gy
    return dict(M=fin, N=fout, K=B, lda=fin, ldb=fout, a_major=1, b_major=1)


@torch.library.custom_op("tltorch::blocktt3_linear", mutates_args=())
def _linear_op(x: torch.Tensor, c1: torch.Tensor, c2: torch.Tensor, c3: torch.Tensor, bias: Optional[torch.Tensor],
               transpose: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    rows, cols, r1, r2 = _meta(c1, c2, c3)
    R, Cn = rows[0] * rows[1] * rows[2], cols[0] * cols[1] * cols[2]
    x = x.contiguous()
    B = x.shape[0]
    W = torch.empty((R, Cn), dtype=c1.dtype, device=c1.device)
    _execute_btt3_reconstruct(c1.contiguous(), c2.contiguous(), c3.contiguous(), W, rows, cols, r1, r2)
    y = torch.empty((B, R if transpose else Cn), dtype=x.dtype, device=x.device)
    if B:
        _execute_gemm(x, W, bias, y, **_gemm_args(transpose, "fwd", B, R, Cn))
    return y, W


@_linear_op.register_fake
def _(x, c1, c2, c3, bias, transpose):
    R, Cn = c1.shape[1] * c2.shape[1] * c3.shape[1], c1.shape[2] * c2.shape[2] * c3.shape[2]
    return x.new_empty((x.shape[0], R if transpose else Cn)), c1.new_empty((R, Cn))


@torch.library.custom_op("tltorch::blocktt3_linear_bwd", mutates_args=())
def _linear_bwd_op(gy: torch.Tensor, x: torch.Tensor, W: torch.Tensor, c1: torch.Tensor, c2: torch.Tensor,
                   c3: torch.Tensor, transpose: bool, need_dx: bool, need_dw: bool
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    rows, cols, r1, r2 = _meta(c1, c2, c3)
    R, Cn = W.shape
    gy, x = gy.contiguous(), x.contiguous()
    B = x.shape[0]
    dx = torch.empty_like(x) if need_dx else x.new_empty(0)
    if need_dx and B:
        _execute_gemm(gy, W, None, dx, **_gemm_args(transpose, "dx", B, R, Cn))
    if need_dw:
        dW = torch.empty((R, Cn), dtype=torch.float32, device=x.device)      # fp32: split-K partials meet here
        if B:
            a, b = (gy, x) if transpose else (x, gy)
            _execute_gemm(a, b, None, dW, **_gemm_args(transpose, "dw", B, R, Cn))
        else:
            dW.zero_()
        d1, d2, d3 = torch.empty_like(c1), torch.empty_like(c2), torch.empty_like(c3)
        _execute_btt3_core_grads(c1.contiguous(), c2.contiguous(), c3.contiguous(), dW, d1, d2, d3, rows, cols, r1, r2)
    else:
        d1, d2, d3 = c1.new_empty(0), c2.new_empty(0), c3.new_empty(0)
    return dx, d1, d2, d3


@_linear_bwd_op.register_fake
def _(gy, x, W, c1, c2, c3, transpose, need_dx, need_dw):
    return (torch.empty_like(x) if need_dx else x.new_empty(0),
            torch.empty_like(c1) if need_dw else c1.new_empty(0),
            torch.empty_like(c2) if need_dw else c2.new_empty(0),
            torch.empty_like(c3) if need_dw else c3.new_empty(0))


def _linear_setup(ctx, inputs, output):
    x, c1, c2, c3, bias, transpose = inputs
    _, W = output
    ctx.save_for_backward(x, W, c1, c2, c3)
    ctx.transpose = transpose
    ctx.has_bias = bias is not None
    ctx.set_materialize_grads(False)        # W is an auxiliary output: do not build a zero gradient for it


def _linear_backward(ctx, gy, gW_unused):
    if gy is None:
        return None, None, None, None, None, None
    x, W, c1, c2, c3 = ctx.saved_tensors
    need_dx = ctx.needs_input_grad[0]
    need_dw = any(ctx.needs_input_grad[1:4])
    dx, d1, d2, d3 = _linear_bwd_op(gy, x, W, c1, c2, c3, ctx.transpose, need_dx, need_dw)
    db = ops.reduce_to(gy, "bo", "o") if (ctx.has_bias and ctx.needs_input_grad[4]) else None
    return (dx if need_dx else None,
            d1 if need_dw else None, d2 if need_dw else None, d3 if need_dw else None, db, None)


_linear_op.register_autograd(_linear_backward, setup_context=_linear_setup)


def blocktt3_linear(x2, c1, c2, c3, bias=None, transpose=True):
    """x2 (B, in) bf16 -> (B, out): dense-regime BlockTT linear with the weight rebuilt on chip-resident cores."""
    y, _ = _linear_op(x2, c1, c2, c3, bias, transpose)
    return y


def linear_eligible(x2, cores, bias, transpose):
    """Can the fused tensor-core route take this call?  (bf16, 3 cores, smem fit, TMA alignment rules)"""
    if len(cores) != 3 or x2.dtype != torch.bfloat16 or any(c.dtype != torch.bfloat16 for c in cores):
        return False
    if bias is not None and bias.dtype not in (torch.bfloat16, torch.float32):
        return False
    rows = [c.shape[1] for c in cores]
    cols = [c.shape[2] for c in cores]
    if cores[0].shape[0] != 1 or cores[2].shape[3] != 1:
        return False
    if not blocktt3_fits(rows, cols, cores[0].shape[3], cores[1].shape[3], True):
        return False
    R, Cn = rows[0] * rows[1] * rows[2], cols[0] * cols[1] * cols[2]
    return R % 8 == 0 and Cn % 8 == 0 and x2.shape[0] > 0
