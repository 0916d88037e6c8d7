"""fp32 CP-factorized linear layer on the bf16 tensor cores (3-way bf16-split operands, fp32 accumulation in TMEM).

  tltorch::cplinear_tc_fwd(x (B, I), lambda (R), factors_out, factors_in, bias | None) -> [y (B, O), saved operands...]
  tltorch::cplinear_tc_bwd(...)                                                       -> [dx, dlambda, dbias, dV_1.., dU_1..]

Every product of linear_cp -> tensor_dot_cp (tltorch/functional/factorized_linear.py:23-36, factorized_tensordot.py:70-103) and of its
autograd -- z = x KR_in, y = z (lambda KR_out)^T, dz = dy (lambda KR_out), dx = dz KR_in^T, dKR_in = x^T dz, dKR_out = dy^T z -- runs as
ONE `tlt_gemm_bf16_tc` launch whose reduction dimension holds the six partial products of the split (tlt_split3_bf16, csrc/split3.cu:
measured 5e-7 .. 1.5e-6 relative error, inside the layer's 1e-5 tolerance, where plain bf16 / TF32 products are 1e-3).  The
unsplit chain ran these six products as fp32 SIMT contractions at ~12 TFLOP/s on 16-74 of the 148 SMs (55 % of the step at BASELINE
config 1).  Khatri-Rao operands and their gradients stay on `tlt_khatri_rao_*`.
CPU tensors raise NotImplementedError (no fallback); tests/_emulator.py swaps the funnels.
"""
import ctypes as C
import os
from typing import List, Optional

import torch

from . import _cuda
from . import kr_ops
from . import streams

__all__ = ["cp_linear_tc", "eligible"]


def _r8(v):
    return (v + 7) // 8 * 8


def eligible(x2, weights, factors_out, factors_in, bias):
    if os.environ.get("TLT_CPLINEAR_TC", "1") == "0":
        return False
    ts = [x2, weights] + list(factors_out) + list(factors_in) + ([bias] if bias is not None else [])
    if any(t.dtype != torch.float32 for t in ts) or not (1 <= len(factors_in) <= 4 and 1 <= len(factors_out) <= 4):
        return False
    i = 1
    for f in factors_in:
        i *= f.shape[0]
    o = 1
    for f in factors_out:
        o *= f.shape[0]
    # 16-byte pitches for every TMA operand (bf16 rows of I, O and round8(R) elements) and enough work to pay for the split launches
    return x2.dim() == 2 and x2.shape[1] == i and i % 8 == 0 and o % 8 == 0 and x2.shape[0] >= 16 and weights.shape[0] >= 64


def _execute_split(src, rows, cols, ld, pat_cols, out_cols, pat_rows, out_rows):
    _cuda.require_cuda(src, out_cols, out_rows)
    _cuda.check(_cuda.lib().tlt_split3_bf16(_cuda.ptr(src), rows, cols, ld, pat_cols, _cuda.ptr(out_cols), pat_rows, _cuda.ptr(out_rows),
                                            _cuda.stream_ptr(src)), "tlt_split3_bf16")


def _execute_kr_bwd_ld(factors, weights, g, ldg, dfs, dw):
    """every factor gradient (+ the weight gradient) of one Khatri-Rao product in one launch, g at leading dimension ldg"""
    _cuda.require_cuda(g, weights, dw, *factors, *dfs)
    n, dims, ptrs = kr_ops._args(factors)
    dptrs = (C.c_void_p * n)(*[d.data_ptr() for d in dfs])
    _cuda.check(_cuda.lib().tlt_khatri_rao_bwd_all(_cuda.dtype_code(g.dtype), n, dims, factors[0].shape[1], ptrs, _cuda.ptr(weights),
                                                   _cuda.ptr(g), ldg, dptrs, _cuda.ptr(dw), _cuda.stream_ptr(g)), "tlt_khatri_rao_bwd_all")


def _execute_kr_split(factors, weights, pat_cols, out_cols, pat_rows, out_rows):
    _cuda.require_cuda(weights, out_cols, out_rows, *factors)
    n, dims, ptrs = kr_ops._args(factors)
    _cuda.check(_cuda.lib().tlt_khatri_rao_split3(n, dims, factors[0].shape[1], ptrs, _cuda.ptr(weights), pat_cols, _cuda.ptr(out_cols),
                                                  pat_rows, _cuda.ptr(out_rows), _cuda.stream_ptr(factors[0])), "tlt_khatri_rao_split3")


def _kr_split(factors, weights, rows, R):
    """Khatri-Rao matrix (rows, R) of fp32 factors as B-side split operands: cols form (rows, 6 R8) and rows form (6 rows, R8)."""
    R8 = _r8(R)
    oc = torch.empty((rows, 6 * R8), dtype=torch.bfloat16, device=factors[0].device)
    orr = torch.empty((6 * rows, R8), dtype=torch.bfloat16, device=factors[0].device)
    _execute_kr_split(factors, weights, 1, oc, 1, orr)
    return oc, orr


def _fork(dev):
    return streams.fork(dev, "TLT_CPLINEAR_BRANCHES")


_join = streams.join


def _split(src, cols=False, rows=False, pat_cols=0, pat_rows=0):
    """src fp32 (n, c) [row stride = src.stride(0)] -> (cols form (n, 6 round8(c)) | None, rows form (6 n, round8(c)) | None), bf16."""
    n, c = src.shape
    cp = _r8(c)
    oc = torch.empty((n, 6 * cp), dtype=torch.bfloat16, device=src.device) if cols else None
    orr = torch.empty((6 * n, cp), dtype=torch.bfloat16, device=src.device) if rows else None
    _execute_split(src, n, c, src.stride(0), pat_cols, oc, pat_rows, orr)
    return oc, orr


def _gemm(a, b, bias, out, M, N, K, lda, ldb, a_major, b_major):
    from .blocktt_ops import _execute_gemm
    _execute_gemm(a, b, bias, out, M=M, N=N, K=K, lda=lda, ldb=ldb, a_major=a_major, b_major=b_major)


def _prod(fs):
    p = 1
    for f in fs:
        p *= f.shape[0]
    return p


@torch.library.custom_op("tltorch::cplinear_tc_fwd", mutates_args=())
def _fwd_op(x: torch.Tensor, lam: torch.Tensor, factors_out: List[torch.Tensor], factors_in: List[torch.Tensor],
            bias: Optional[torch.Tensor]) -> List[torch.Tensor]:
    """-> [y, z (B, R8) fp32, x in rows form, z in rows form, KR_in in cols form, lambda KR_out in rows form]."""
    x, lam = x.contiguous(), lam.contiguous()
    f_in, f_out = [f.contiguous() for f in factors_in], [f.contiguous() for f in factors_out]
    B, I, O, R = x.shape[0], _prod(f_in), _prod(f_out), lam.shape[0]
    R8 = _r8(R)
    dev, f32 = x.device, torch.float32
    # every buffer is allocated here, on the current stream; the output-side Khatri-Rao operand is built on a side stream beside
    # the first product (each of these launches is ~10 us of mostly latency at config 1: two independent branches halve the chain)
    bf = torch.bfloat16
    x_c, x_r = torch.empty((B, 6 * I), dtype=bf, device=dev), torch.empty((6 * B, I), dtype=bf, device=dev)
    krin_c, krin_r = torch.empty((I, 6 * R8), dtype=bf, device=dev), torch.empty((6 * I, R8), dtype=bf, device=dev)
    kro_c, kro_r = torch.empty((O, 6 * R8), dtype=bf, device=dev), torch.empty((6 * O, R8), dtype=bf, device=dev)
    z_c, z_r = torch.empty((B, 6 * R8), dtype=bf, device=dev), torch.empty((6 * B, R8), dtype=bf, device=dev)
    z = torch.empty((B, R8), dtype=f32, device=dev)                  # z = x . KR_in   (pad columns: exact zeros)
    y = torch.empty((B, O), dtype=f32, device=dev)                   # y = z . (lambda KR_out)^T + bias
    side, main = _fork(dev)
    if side is not None:
        with torch.cuda.stream(side):
            _execute_kr_split(f_out, lam, 1, kro_c, 1, kro_r)        # lambda folded in
    else:
        _execute_kr_split(f_out, lam, 1, kro_c, 1, kro_r)
    _execute_split(x, B, I, x.stride(0), 0, x_c, 0, x_r)
    _execute_kr_split(f_in, None, 1, krin_c, 1, krin_r)
    _gemm(x_c, krin_r, None, z, M=B, N=R8, K=6 * I, lda=6 * I, ldb=R8, a_major=0, b_major=1)
    _execute_split(z, B, R8, R8, 0, z_c, 1, z_r)
    _join(side, main)
    _gemm(z_c, kro_c, None if bias is None else bias.contiguous(), y, M=B, N=O, K=6 * R8, lda=6 * R8, ldb=6 * R8, a_major=0, b_major=0)
    return [y, x_r, z_r, krin_c, kro_r]


@_fwd_op.register_fake
def _(x, lam, factors_out, factors_in, bias):
    B, I, O, R8 = x.shape[0], _prod(factors_in), _prod(factors_out), _r8(lam.shape[0])
    bf = torch.bfloat16
    return [x.new_empty((B, O)), x.new_empty((6 * B, I), dtype=bf), x.new_empty((6 * B, R8), dtype=bf), x.new_empty((I, 6 * R8), dtype=bf),
            x.new_empty((6 * O, R8), dtype=bf)]


@torch.library.custom_op("tltorch::cplinear_tc_bwd", mutates_args=())
def _bwd_op(gy: torch.Tensor, x_r: torch.Tensor, z_r: torch.Tensor, krin_c: torch.Tensor, kro_r: torch.Tensor, lam: torch.Tensor,
            factors_out: List[torch.Tensor], factors_in: List[torch.Tensor], need_bias: bool) -> List[torch.Tensor]:
    """-> [dx, dlambda, dbias | empty, dV_1.., dU_1..] (all fp32)."""
    gy, lam = gy.contiguous(), lam.contiguous()
    f_in, f_out = [f.contiguous() for f in factors_in], [f.contiguous() for f in factors_out]
    B, O = gy.shape
    I, R = krin_c.shape[0], lam.shape[0]
    R8 = _r8(R)
    dev, f32 = gy.device, torch.float32
    bf = torch.bfloat16
    gy_c, gy_r = torch.empty((B, 6 * O), dtype=bf, device=dev), torch.empty((6 * B, O), dtype=bf, device=dev)
    dz_c, dz_r = torch.empty((B, 6 * R8), dtype=bf, device=dev), torch.empty((6 * B, R8), dtype=bf, device=dev)
    dz = torch.empty((B, R8), dtype=f32, device=dev)                 # dz = dy . (lambda KR_out)
    dx = torch.empty((B, I), dtype=f32, device=dev)                  # dx = dz . KR_in^T
    dkr_in = torch.empty((I, R8), dtype=f32, device=dev)             # dKR_in = x^T . dz
    dkr_out = torch.empty((O, R8), dtype=f32, device=dev)            # d(lambda KR_out) = dy^T . z
    df_in, df_out = [torch.empty_like(f) for f in f_in], [torch.empty_like(f) for f in f_out]
    dlam = torch.zeros_like(lam)
    _execute_split(gy, B, O, gy.stride(0), 0, gy_c, 0, gy_r)
    # output-side gradients (dy^T z -> dV_m, dlambda, dbias) need nothing from the input side: a side branch beside dz -> dx -> dU_m
    side, main = _fork(dev)

    def out_side():
        _gemm(gy_r, z_r, None, dkr_out, M=O, N=R8, K=6 * B, lda=O, ldb=R8, a_major=1, b_major=1)
        _execute_kr_bwd_ld(f_out, lam, dkr_out, R8, df_out, dlam)
        if need_bias:
            from .ops import reduce_to
            return reduce_to(gy, "bo", "o")
        return torch.empty(0, dtype=f32, device=dev)

    if side is not None:
        with torch.cuda.stream(side):
            dbias = out_side()
    _gemm(gy_c, kro_r, None, dz, M=B, N=R8, K=6 * O, lda=6 * O, ldb=R8, a_major=0, b_major=1)
    _execute_split(dz, B, R8, R8, 0, dz_c, 1, dz_r)
    _gemm(dz_c, krin_c, None, dx, M=B, N=I, K=6 * R8, lda=6 * R8, ldb=6 * R8, a_major=0, b_major=0)
    _gemm(x_r, dz_r, None, dkr_in, M=I, N=R8, K=6 * B, lda=I, ldb=R8, a_major=1, b_major=1)
    _execute_kr_bwd_ld(f_in, None, dkr_in, R8, df_in, None)
    if side is None:
        dbias = out_side()
    _join(side, main)
    if side is not None and dbias.numel():
        dbias.record_stream(main)                                    # allocated while the side stream was current
    return [dx, dlam, dbias] + df_out + df_in


@_bwd_op.register_fake
def _(gy, x_r, z_r, krin_c, kro_r, lam, factors_out, factors_in, need_bias):
    return ([gy.new_empty((gy.shape[0], krin_c.shape[0])), torch.empty_like(lam), gy.new_empty(gy.shape[1] if need_bias else 0)]
            + [torch.empty_like(f) for f in factors_out] + [torch.empty_like(f) for f in factors_in])


def _setup(ctx, inputs, output):
    x, lam, factors_out, factors_in, bias = inputs
    ctx.n_out, ctx.n_in, ctx.has_bias = len(factors_out), len(factors_in), bias is not None
    ctx.save_for_backward(output[1], output[2], output[3], output[4], lam, *factors_out, *factors_in)
    ctx.set_materialize_grads(False)


def _backward(ctx, grads):
    if grads[0] is None:
        return None, None, None, None, None
    saved = list(ctx.saved_tensors)
    x_r, z_r, krin_c, kro_r, lam = saved[:5]
    f_out, f_in = saved[5:5 + ctx.n_out], saved[5 + ctx.n_out:]
    need_bias = ctx.has_bias and ctx.needs_input_grad[4]
    outs = _bwd_op(grads[0], x_r, z_r, krin_c, kro_r, lam, f_out, f_in, need_bias)
    return outs[0], outs[1], list(outs[3:3 + ctx.n_out]), list(outs[3 + ctx.n_out:]), (outs[2] if need_bias else None)


_fwd_op.register_autograd(_backward, setup_context=_setup)


def cp_linear_tc(x2, weights, factors_out, factors_in, bias):
    """x2 (B, I) fp32 -> (B, O): y = ((x2 . KR(factors_in)) * weights) . KR(factors_out)^T + bias on the tensor cores (split bf16)."""
    return _fwd_op(x2, weights, list(factors_out), list(factors_in), bias)[0]
