"""Custom ops (with autograd) for the fused CP-factorized linear layer (fp32).

  tltorch::cplinear_fwd(x (B, I), lambda (R), factors_out [(e_m, R)], factors_in [(d_m, R)], bias (O) | None) -> [y (B, O), z (B, round4(R))]
  tltorch::cplinear_bwd(x, gy, z, lambda, factors_out, factors_in, need_bias) -> [dx, dlambda, dbias, dV_1.., dU_1..]

Two launches forward, four backward (tlt_cplinear_fwd / _bwd, torch_b200/csrc/cplinear.cu): project the activation onto the rank
space with the Khatri-Rao tile of the input factors rebuilt in shared memory, expand with the lambda-weighted tile of the output
factors (+ bias); the backward pass gives dx, every factor gradient, dlambda and dbias without materialising a Khatri-Rao matrix
or a dense weight gradient.  Reference: linear_cp (tltorch/functional/factorized_linear.py:23-36) -> tensor_dot_cp
(factorized_tensordot.py:70-103) and their autograd.  CPU tensors raise NotImplementedError (no fallback); tests/_emulator.py swaps
the two funnels.
"""
import ctypes as C
from typing import List, Optional

import torch

from . import _cuda

__all__ = ["cp_linear_fused", "eligible"]


def enabled():
    """Opt-in (TLT_CPLINEAR_FUSED=1).  The fused op is parity-green (1e-5) and cuts the step from 20 launches to 6, but its fp32 SIMT
    kernels are latency-bound at config 1's size (rank tiles of 16 columns give one 4-8 warp block per SM): 0.31 ms per step
    against 0.24 ms for the unfused chain (profiles/r02_cplinear.md), so the chain stays the default."""
    import os
    return os.environ.get("TLT_CPLINEAR_FUSED", "0") == "1"


def eligible(x2, weights, factors_out, factors_in, bias):
    if not enabled():
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
    if i % 4 or o % 4 or x2.dim() != 2 or x2.shape[1] != i or x2.shape[0] < 1:
        return False
    return max(sum(f.shape[0] for f in factors_in), sum(f.shape[0] for f in factors_out)) <= 512


def _desc(batch, rank, in_dims, out_dims):
    d = _cuda.CpLinearDesc()
    d.batch, d.rank, d.n_in, d.n_out = batch, rank, len(in_dims), len(out_dims)
    for k, v in enumerate(in_dims):
        d.in_dims[k] = v
    for k, v in enumerate(out_dims):
        d.out_dims[k] = v
    return d


def _ptr_array(ts):
    return (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])


def _execute_fwd(d, x, f_in, f_out, lam, bias, z, y):
    _cuda.require_cuda(x, lam, bias, z, y, *f_in, *f_out)
    _cuda.check(_cuda.lib().tlt_cplinear_fwd(C.byref(d), _cuda.ptr(x), _ptr_array(f_in), _ptr_array(f_out), _cuda.ptr(lam), _cuda.ptr(bias),
                                             _cuda.ptr(z), _cuda.ptr(y), _cuda.stream_ptr(x)), "tlt_cplinear_fwd")


def _execute_bwd(d, x, gy, z, f_in, f_out, lam, g, dx, df_in, df_out, dlam, dbias):
    _cuda.require_cuda(x, gy, z, lam, g, dx, dlam, dbias, *f_in, *f_out, *df_in, *df_out)
    _cuda.check(_cuda.lib().tlt_cplinear_bwd(C.byref(d), _cuda.ptr(x), _cuda.ptr(gy), _cuda.ptr(z), _ptr_array(f_in), _ptr_array(f_out),
                                             _cuda.ptr(lam), _cuda.ptr(g), _cuda.ptr(dx), _ptr_array(df_in), _ptr_array(df_out),
                                             _cuda.ptr(dlam), _cuda.ptr(dbias), _cuda.stream_ptr(x)), "tlt_cplinear_bwd")


def _r4(v):
    return (v + 3) // 4 * 4


@torch.library.custom_op("tltorch::cplinear_fwd", mutates_args=())
def _fwd_op(x: torch.Tensor, lam: torch.Tensor, factors_out: List[torch.Tensor], factors_in: List[torch.Tensor],
            bias: Optional[torch.Tensor]) -> List[torch.Tensor]:
    x, lam = x.contiguous(), lam.contiguous()
    f_in, f_out = [f.contiguous() for f in factors_in], [f.contiguous() for f in factors_out]
    b, r = x.shape[0], lam.shape[0]
    in_dims, out_dims = [f.shape[0] for f in f_in], [f.shape[0] for f in f_out]
    o = 1
    for v in out_dims:
        o *= v
    z = torch.empty((b, _r4(r)), dtype=torch.float32, device=x.device)
    y = torch.empty((b, o), dtype=torch.float32, device=x.device)
    _execute_fwd(_desc(b, r, in_dims, out_dims), x, f_in, f_out, lam, None if bias is None else bias.contiguous(), z, y)
    return [y, z]


@_fwd_op.register_fake
def _(x, lam, factors_out, factors_in, bias):
    o = 1
    for f in factors_out:
        o *= f.shape[0]
    return [x.new_empty((x.shape[0], o)), x.new_empty((x.shape[0], _r4(lam.shape[0])))]


@torch.library.custom_op("tltorch::cplinear_bwd", mutates_args=())
def _bwd_op(x: torch.Tensor, gy: torch.Tensor, z: torch.Tensor, lam: torch.Tensor, factors_out: List[torch.Tensor],
            factors_in: List[torch.Tensor], need_bias: bool) -> List[torch.Tensor]:
    x, gy, lam = x.contiguous(), gy.contiguous(), lam.contiguous()
    f_in, f_out = [f.contiguous() for f in factors_in], [f.contiguous() for f in factors_out]
    b, r = x.shape[0], lam.shape[0]
    in_dims, out_dims = [f.shape[0] for f in f_in], [f.shape[0] for f in f_out]
    dev = x.device
    g = torch.empty_like(z)
    dx = torch.empty_like(x)
    df_in, df_out = [torch.empty_like(f) for f in f_in], [torch.empty_like(f) for f in f_out]
    dlam = torch.empty_like(lam)
    dbias = torch.empty(gy.shape[1], dtype=torch.float32, device=dev) if need_bias else None
    _execute_bwd(_desc(b, r, in_dims, out_dims), x, gy, z, f_in, f_out, lam, g, dx, df_in, df_out, dlam, dbias)
    return [dx, dlam, dbias if need_bias else torch.empty(0, dtype=torch.float32, device=dev)] + df_out + df_in


@_bwd_op.register_fake
def _(x, gy, z, lam, factors_out, factors_in, need_bias):
    return ([torch.empty_like(x), torch.empty_like(lam), x.new_empty(gy.shape[1] if need_bias else 0)]
            + [torch.empty_like(f) for f in factors_out] + [torch.empty_like(f) for f in factors_in])


def _setup(ctx, inputs, output):
    x, lam, factors_out, factors_in, bias = inputs
    ctx.n_out, ctx.n_in, ctx.has_bias = len(factors_out), len(factors_in), bias is not None
    ctx.save_for_backward(x, output[1], lam, *factors_out, *factors_in)
    ctx.set_materialize_grads(False)


def _backward(ctx, grads):
    if grads[0] is None:
        return None, None, None, None, None
    saved = list(ctx.saved_tensors)
    x, z, lam = saved[:3]
    f_out, f_in = saved[3:3 + ctx.n_out], saved[3 + ctx.n_out:]
    need_bias = ctx.has_bias and ctx.needs_input_grad[4]
    outs = _bwd_op(x, grads[0], z, lam, f_out, f_in, need_bias)
    dx, dlam, dbias = outs[0], outs[1], (outs[2] if need_bias else None)
    return dx, dlam, list(outs[3:3 + ctx.n_out]), list(outs[3 + ctx.n_out:]), dbias


_fwd_op.register_autograd(_backward, setup_context=_setup)


def cp_linear_fused(x2, weights, factors_out, factors_in, bias):
    """x2 (B, I) fp32 -> (B, O): y = ((x2 . KR(factors_in)) * weights) . KR(factors_out)^T + bias.  Differentiable in everything."""
    return _fwd_op(x2, weights, list(factors_out), list(factors_in), bias)[0]
