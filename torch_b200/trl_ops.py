"""Custom op (with autograd) for the tail of the Tucker tensor-regression layer, and the fused route of tucker_trl built on it.

  tltorch::trl_tail(zp (B, J, ldz), core (rc, q.., ro), f_out (O, ro), bias (O) | None) -> [y (B, O), v (B, ro) fp32]

tucker_trl_fused = spatial_project (one pass over the activation where it lies) -> channel-mode GEMM on tcgen05 -> trl_tail
(tlt_trl_tail_fwd / _bwd, torch_b200/csrc/trl_tail.cu: core contraction + output expansion + bias in one launch; backward: dv /
dzp in one launch, every parameter gradient of the tail in another).  Reference: tucker_trl (tltorch/functional/
tensor_regression.py:37-49).  CPU tensors raise NotImplementedError (no fallback); tests/_emulator.py swaps the two funnels.
"""
import ctypes as C
import math
from typing import List, Optional

import torch

from . import _cuda

__all__ = ["tucker_trl_fused", "eligible", "trl_tail"]


def _desc(batch, j, rc, ldz, ro, out):
    d = _cuda.TrlTailDesc()
    d.batch, d.j, d.rc, d.ldz, d.ro, d.out = batch, j, rc, ldz, ro, out
    return d


def _execute_fwd(d, zp, core, fo, bias, y, v):
    _cuda.require_cuda(zp, core, fo, bias, y, v)
    _cuda.check(_cuda.lib().tlt_trl_tail_fwd(C.byref(d), _cuda.ptr(zp), _cuda.ptr(core), _cuda.ptr(fo), _cuda.ptr(bias), _cuda.ptr(y),
                                             _cuda.ptr(v), _cuda.stream_ptr(zp)), "tlt_trl_tail_fwd")


def _execute_bwd(d, zp, gy, core, fo, v, dv, dzp, dcore, dfo, dbias):
    _cuda.require_cuda(zp, gy, core, fo, v, dv, dzp, dcore, dfo, dbias)
    _cuda.check(_cuda.lib().tlt_trl_tail_bwd(C.byref(d), _cuda.ptr(zp), _cuda.ptr(gy), _cuda.ptr(core), _cuda.ptr(fo), _cuda.ptr(v),
                                             _cuda.ptr(dv), _cuda.ptr(dzp), _cuda.ptr(dcore), _cuda.ptr(dfo), _cuda.ptr(dbias),
                                             _cuda.stream_ptr(zp)), "tlt_trl_tail_bwd")


def _execute_kron_fwd(a, b, k):
    _cuda.require_cuda(a, b, k)
    _cuda.check(_cuda.lib().tlt_kron2_fwd(_cuda.ptr(a), _cuda.ptr(b), a.shape[0], a.shape[1], b.shape[0], b.shape[1], _cuda.ptr(k),
                                          _cuda.stream_ptr(k)), "tlt_kron2_fwd")


def _execute_kron_bwd(a, b, dk, d_a, d_b):
    _cuda.require_cuda(a, b, dk, d_a, d_b)
    _cuda.check(_cuda.lib().tlt_kron2_bwd(_cuda.ptr(a), _cuda.ptr(b), _cuda.ptr(dk), a.shape[0], a.shape[1], b.shape[0], b.shape[1],
                                          _cuda.ptr(d_a), _cuda.ptr(d_b), _cuda.stream_ptr(dk)), "tlt_kron2_bwd")


@torch.library.custom_op("tltorch::kron2", mutates_args=())
def _kron_op(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    a, b = a.contiguous(), b.contiguous()
    k = torch.empty((a.shape[0] * b.shape[0], a.shape[1] * b.shape[1]), dtype=a.dtype, device=a.device)
    _execute_kron_fwd(a, b, k)
    return k


@_kron_op.register_fake
def _(a, b):
    return a.new_empty((a.shape[0] * b.shape[0], a.shape[1] * b.shape[1]))


@torch.library.custom_op("tltorch::kron2_bwd", mutates_args=())
def _kron_bwd_op(a: torch.Tensor, b: torch.Tensor, dk: torch.Tensor) -> List[torch.Tensor]:
    a, b, dk = a.contiguous(), b.contiguous(), dk.contiguous()
    d_a, d_b = torch.empty_like(a), torch.empty_like(b)
    _execute_kron_bwd(a, b, dk, d_a, d_b)
    return [d_a, d_b]


@_kron_bwd_op.register_fake
def _(a, b, dk):
    return [torch.empty_like(a), torch.empty_like(b)]


def _kron_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _kron_backward(ctx, dk):
    a, b = ctx.saved_tensors
    d_a, d_b = _kron_bwd_op(a, b, dk)
    return d_a, d_b


_kron_op.register_autograd(_kron_backward, setup_context=_kron_setup)


def kron2(a, b):
    """(da, pa) x (db, pb) bf16 -> (da db, pa pb): k[(i, j), (p, q)] = a[i, p] b[j, q]; differentiable in both."""
    return _kron_op(a, b)


def _geom(zp, core, fo):
    b, j, ldz = zp.shape
    rc, ro = core.shape[0], core.shape[-1]
    return _desc(b, j, rc, ldz, ro, fo.shape[0])


@torch.library.custom_op("tltorch::trl_tail", mutates_args=())
def _tail_op(zp: torch.Tensor, core: torch.Tensor, f_out: torch.Tensor, bias: Optional[torch.Tensor]) -> List[torch.Tensor]:
    zp, core, f_out = zp.contiguous(), core.contiguous(), f_out.contiguous()
    y = torch.empty((zp.shape[0], f_out.shape[0]), dtype=zp.dtype, device=zp.device)
    v = torch.empty((zp.shape[0], core.shape[-1]), dtype=torch.float32, device=zp.device)
    _execute_fwd(_geom(zp, core, f_out), zp, core, f_out, None if bias is None else bias.contiguous(), y, v)
    return [y, v]


@_tail_op.register_fake
def _(zp, core, f_out, bias):
    return [zp.new_empty((zp.shape[0], f_out.shape[0])), zp.new_empty((zp.shape[0], core.shape[-1]), dtype=torch.float32)]


@torch.library.custom_op("tltorch::trl_tail_bwd", mutates_args=())
def _tail_bwd_op(zp: torch.Tensor, gy: torch.Tensor, core: torch.Tensor, f_out: torch.Tensor, v: torch.Tensor,
                 need_bias: bool) -> List[torch.Tensor]:
    zp, gy, core, f_out = zp.contiguous(), gy.contiguous(), core.contiguous(), f_out.contiguous()
    dzp, dcore, dfo = torch.empty_like(zp), torch.empty_like(core), torch.empty_like(f_out)
    dbias = torch.empty(f_out.shape[0], dtype=zp.dtype, device=zp.device) if need_bias else None
    dv = torch.empty_like(v)
    _execute_bwd(_geom(zp, core, f_out), zp, gy, core, f_out, v, dv, dzp, dcore, dfo, dbias)
    return [dzp, dcore, dfo, dbias if need_bias else zp.new_empty(0)]


@_tail_bwd_op.register_fake
def _(zp, gy, core, f_out, v, need_bias):
    return [torch.empty_like(zp), torch.empty_like(core), torch.empty_like(f_out), zp.new_empty(f_out.shape[0] if need_bias else 0)]


def _setup(ctx, inputs, output):
    zp, core, f_out, bias = inputs
    ctx.has_bias = bias is not None
    ctx.save_for_backward(zp, core, f_out, output[1])
    ctx.set_materialize_grads(False)


def _backward(ctx, grads):
    if grads[0] is None:
        return None, None, None, None
    zp, core, f_out, v = ctx.saved_tensors
    need_bias = ctx.has_bias and ctx.needs_input_grad[3]
    dzp, dcore, dfo, dbias = _tail_bwd_op(zp, grads[0], core, f_out, v, need_bias)
    return dzp, dcore, dfo, (dbias if need_bias else None)


_tail_op.register_autograd(_backward, setup_context=_setup)


def trl_tail(zp, core, f_out, bias=None):
    """zp (B, J, ldz) bf16 (channels [rc, ldz) zero), core (rc, q_1.., ro) with prod(q) = J, f_out (O, ro) -> y (B, O)."""
    return _tail_op(zp, core, f_out, bias)[0]


def eligible(x, weight, bias):
    """bf16 Tucker TRL with one output mode, a large channel mode and 1-3 small spatial modes (the spatial-first projection plan)."""
    from . import spatial_ops
    n_in = x.dim() - 1
    if x.dtype != torch.bfloat16 or not x.is_contiguous() or n_in < 2 or n_in > 4 or weight.order != n_in + 1:
        return False
    ts = [weight.core] + list(weight.factors) + ([bias] if bias is not None else [])
    if any(t.dtype != torch.bfloat16 for t in ts) or (bias is not None and (bias.dim() != 1 or bias.shape[0] != weight.shape[-1])):
        return False
    b, c = x.shape[0], x.shape[1]
    rc, ro = weight.core.shape[0], weight.core.shape[-1]
    S, J = math.prod(x.shape[2:]), math.prod(weight.core.shape[1:-1])
    if ro > 16 or c < 256 or rc * 4 > c or b * S < 1024 or b > 65535:
        return False
    return spatial_ops.eligible(x, S, J)


def tucker_trl_fused(x, weight, bias=None):
    """x (B, C, s_1..) bf16, weight a TuckerTensor of shape (C, s_1.., O) -> (B, O); differentiable in everything."""
    from . import spatial_ops
    from .tucker_ops import padded_linear
    factors = list(weight.factors)
    b, c = x.shape[0], x.shape[1]
    S = math.prod(x.shape[2:])
    k = factors[1]                                                   # (s_1, q_1) ... Kronecker product of the spatial factors
    for m in factors[2:-1]:
        k = kron2(k, m)
    t = spatial_ops.spatial_project(x.reshape(b, c, S), k)          # (B, J, C)
    J = k.shape[1]
    z = padded_linear(t.reshape(b * J, c), factors[0].t(), None, True)   # (B J, round8(rc)), pad columns zero
    return trl_tail(z.reshape(b, J, z.shape[1]), weight.core, factors[-1], bias)
