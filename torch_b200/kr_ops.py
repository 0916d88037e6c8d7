"""Custom op (with autograd) for the weighted Khatri-Rao product  KR[(i_1..i_n), r] = w_r prod_k F_k[i_k, r].

One element-wise launch forward, one launch per factor backward (tlt_khatri_rao_fwd / _bwd) instead of a chain of pairwise
contractions in which the rank index is a batch index (R = 2341 blocks per step at BASELINE config 1).
Reference: the Khatri-Rao operands of tensor_dot_cp / linear_cp (tltorch/functional/factorized_tensordot.py:70-103) and
tl.cp_to_tensor inside CPTensor.to_tensor (factorized_tensors/factorized_tensors.py:129-130).
CPU tensors raise NotImplementedError (no fallback); tests/_emulator.py swaps the two funnels.
"""
import ctypes as C
from typing import List, Optional

import torch

from . import _cuda

__all__ = ["khatri_rao", "eligible"]

MAX_FACTORS = 4


def eligible(factors, weights=None):
    if not 1 <= len(factors) <= MAX_FACTORS:
        return False
    dt = factors[0].dtype
    if dt not in (torch.float32, torch.bfloat16):
        return False
    ts = list(factors) + ([weights] if weights is not None else [])
    return all(t.dtype == dt for t in ts) and all(f.dim() == 2 and f.shape[1] == factors[0].shape[1] for f in factors)


def _args(factors):
    n = len(factors)
    dims = (C.c_int64 * n)(*[f.shape[0] for f in factors])
    ptrs = (C.c_void_p * n)(*[f.data_ptr() for f in factors])
    return n, dims, ptrs


def _execute_kr_fwd(factors, weights, out):
    _cuda.require_cuda(out, weights, *factors)
    n, dims, ptrs = _args(factors)
    _cuda.check(_cuda.lib().tlt_khatri_rao_fwd(_cuda.dtype_code(out.dtype), n, dims, factors[0].shape[1], ptrs, _cuda.ptr(weights),
                                               _cuda.ptr(out), _cuda.stream_ptr(out)), "tlt_khatri_rao_fwd")


def _execute_kr_bwd(factors, weights, g, dfs, dw):
    """dfs: list of fp32 (d_k, R) outputs (None to skip); dw: fp32 (R) accumulated into, or None."""
    _cuda.require_cuda(g, weights, dw, *factors, *dfs)
    n, dims, ptrs = _args(factors)
    dptrs = (C.c_void_p * n)(*[None if d is None else d.data_ptr() for d in dfs])
    _cuda.check(_cuda.lib().tlt_khatri_rao_bwd(_cuda.dtype_code(g.dtype), n, dims, factors[0].shape[1], ptrs, _cuda.ptr(weights),
                                               _cuda.ptr(g), dptrs, _cuda.ptr(dw), _cuda.stream_ptr(g)), "tlt_khatri_rao_bwd")


@torch.library.custom_op("tltorch::khatri_rao", mutates_args=())
def _kr_op(factors: List[torch.Tensor], weights: Optional[torch.Tensor]) -> torch.Tensor:
    factors = [f.contiguous() for f in factors]
    rows = 1
    for f in factors:
        rows *= f.shape[0]
    out = torch.empty((rows, factors[0].shape[1]), dtype=factors[0].dtype, device=factors[0].device)
    if out.numel():
        _execute_kr_fwd(factors, None if weights is None else weights.contiguous(), out)
    return out


@_kr_op.register_fake
def _(factors, weights):
    rows = 1
    for f in factors:
        rows *= f.shape[0]
    return factors[0].new_empty((rows, factors[0].shape[1]))


@torch.library.custom_op("tltorch::khatri_rao_bwd", mutates_args=())
def _kr_bwd_op(factors: List[torch.Tensor], weights: Optional[torch.Tensor], g: torch.Tensor, need: List[bool],
               need_w: bool) -> List[torch.Tensor]:
    factors = [f.contiguous() for f in factors]
    dev = g.device
    need = list(need)
    if need_w:
        need[0] = True                                   # the lambda gradient is reduced from factor 0's pass
    dfs = [torch.empty(f.shape, dtype=torch.float32, device=dev) if nd else None for f, nd in zip(factors, need)]
    dw = torch.zeros(factors[0].shape[1], dtype=torch.float32, device=dev) if need_w else None
    if g.numel():
        _execute_kr_bwd(factors, None if weights is None else weights.contiguous(), g.contiguous(), dfs, dw)
    else:
        for d in dfs:
            if d is not None:
                d.zero_()
    empty = torch.empty(0, dtype=torch.float32, device=dev)
    return [empty if d is None else d for d in dfs] + [empty if dw is None else dw]


@_kr_bwd_op.register_fake
def _(factors, weights, g, need, need_w):
    need = list(need)
    if need_w:
        need[0] = True
    e = g.new_empty(0, dtype=torch.float32)
    return [g.new_empty(f.shape, dtype=torch.float32) if nd else e for f, nd in zip(factors, need)] + \
           [g.new_empty(factors[0].shape[1], dtype=torch.float32) if need_w else e]


def _kr_setup(ctx, inputs, output):
    factors, weights = inputs
    ctx.n = len(factors)
    ctx.has_w = weights is not None
    ctx.save_for_backward(*factors, *([weights] if weights is not None else []))


def _kr_backward(ctx, g):
    saved = list(ctx.saved_tensors)
    factors, weights = saved[:ctx.n], (saved[ctx.n] if ctx.has_w else None)
    need = [bool(v) for v in ctx.needs_input_grad[0]] if isinstance(ctx.needs_input_grad[0], (list, tuple)) else [bool(ctx.needs_input_grad[0])] * ctx.n
    need_w = ctx.has_w and bool(ctx.needs_input_grad[1])
    outs = _kr_bwd_op(factors, weights, g, need, need_w)
    gf = [outs[k].to(factors[k].dtype) if need[k] else None for k in range(ctx.n)]
    gw = outs[ctx.n].to(weights.dtype) if need_w else None
    return gf, gw


_kr_op.register_autograd(_kr_backward, setup_context=_kr_setup)


def khatri_rao(factors, weights=None):
    """(prod d_k, R) weighted Khatri-Rao product (first factor slowest), differentiable in the factors and the weights."""
    return _kr_op(list(factors), weights)
