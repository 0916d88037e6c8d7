"""Custom op (with autograd) for the tail of the Tucker tensor-regression layer, and the fused route of tucker_trl built on it.

  tltorch::trl_tail(zp (B, J, ldz), core (rc, q.., ro), f_out (O, ro), bias (O) | None) -> [y (B, O), v (B, ro) fp32]

tucker_trl_fused = spatial_project (one pass over the activation where it lies) -> channel-mode GEMM on tcgen05 -> trl_tail
(tlt_trl_tail_fwd / _bwd, torch_b200/csrc/trl_tail.cu: core contraction + output expansion + bias in one launch; backward: dv /
dzp in one launch, every parameter gradient of the tail in another).  Reference: tucker_trl (tltorch/functional/
tensor_regression.py:37-49).  CPU tensors raise NotImplementedError (no fallback); tests/_emulator.py swaps the two funnels.
"""
import ctypes as C
import math
import os
from typing import List, Optional

import torch

from . import _cuda
from . import streams

__all__ = ["tucker_trl_fused", "eligible", "trl_tail"]


def _desc(batch, j, rc, ldz, ro, out):
    d = _cuda.TrlTailDesc()
    d.batch, d.j, d.rc, d.ldz, d.ro, d.out = batch, j, rc, ldz, ro, out
    return d


def _execute_fwd(d, zp, core, fo, bias, y, v):
    _cuda.require_cuda(zp, core, fo, bias, y, v)
    _cuda.check(_cuda.lib().tlt_trl_tail_fwd(C.byref(d), _cuda.ptr(zp), _cuda.ptr(core), _cuda.ptr(fo), _cuda.ptr(bias), _cuda.ptr(y),
                                             _cuda.ptr(v), _cuda.stream_ptr(zp)), "tlt_trl_tail_fwd")


def _execute_bwd(d, zp, gy, core, fo, v, dv, dzp, dcore, dfo, dbias, phases=3):
    """phases: 1 = dv / dzp, 2 = dcore / dF_o / dbias (reads dv), 3 = both launches"""
    _cuda.require_cuda(zp, gy, core, fo, v, dv, dzp, dcore, dfo, dbias)
    _cuda.check(_cuda.lib().tlt_trl_tail_bwd(C.byref(d), phases, _cuda.ptr(zp), _cuda.ptr(gy), _cuda.ptr(core), _cuda.ptr(fo), _cuda.ptr(v),
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
    _execute_bwd(_geom(zp, core, f_out), zp, gy, core, f_out, v, dv, dzp, dcore, dfo, dbias, phases=3)
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


# =====================================================================================================================
# the whole fused route as ONE op per direction: every launch is issued here, so the backward pass can put the parameter gradients
# (tail parameters, channel factor) on a side stream beside dt -> spatial backward -> Kronecker backward
# =====================================================================================================================
def _fork(dev):
    return streams.fork(dev, "TLT_TRL_BRANCHES")


def _r8(v):
    return (v + 7) // 8 * 8


def _kron_chain(spatial_factors):
    """Kronecker product of the spatial factors, left to right; returns (k, intermediates) for the backward pass."""
    k, chain = spatial_factors[0], []
    for m in spatial_factors[1:]:
        out = torch.empty((k.shape[0] * m.shape[0], k.shape[1] * m.shape[1]), dtype=k.dtype, device=k.device)
        _execute_kron_fwd(k, m, out)
        chain.append((k, m))
        k = out
    return k, chain


@torch.library.custom_op("tltorch::tucker_trl_fwd", mutates_args=())
def _trl_fwd_op(x: torch.Tensor, core: torch.Tensor, factors: List[torch.Tensor], bias: Optional[torch.Tensor]) -> List[torch.Tensor]:
    """-> [y (B, O), t (B, J, C), zp (B, J, ldz), wp (ldz, round8(C)), v (B, ro) fp32, k (S, J)]"""
    from . import spatial_ops
    from .blocktt_ops import _execute_gemm
    from .tucker_ops import _execute_pack_kmajor
    x, core = x.contiguous(), core.contiguous()
    fs = [f.contiguous() for f in factors]
    b, c = x.shape[0], x.shape[1]
    S = math.prod(x.shape[2:])
    dev, bf = x.device, x.dtype
    k, _ = _kron_chain(fs[1:-1])
    if len(fs) == 3:
        k = k.clone()                                                 # one spatial mode: k IS a factor; an op may not return an input
    J, rc = k.shape[1], fs[0].shape[1]
    t = torch.empty((b, J, c), dtype=bf, device=dev)
    spatial_ops._execute_spatial_fwd(x.reshape(b, c, S), k, t)
    ldz, cp = _r8(rc), _r8(c)
    wp = torch.empty((ldz, cp), dtype=bf, device=dev)                 # F_c^T re-pitched K-major, pad rows zero
    _execute_pack_kmajor(fs[0].t(), wp[:rc])
    if ldz != rc:
        wp[rc:].zero_()
    zp = torch.empty((b * J, ldz), dtype=bf, device=dev)
    _execute_gemm(t.reshape(b * J, c), wp, None, zp, M=b * J, N=ldz, K=cp if cp == c else c, lda=c, ldb=cp, a_major=0, b_major=0)
    fo = fs[-1]
    y = torch.empty((b, fo.shape[0]), dtype=bf, device=dev)
    v = torch.empty((b, core.shape[-1]), dtype=torch.float32, device=dev)
    zp3 = zp.reshape(b, J, ldz)
    _execute_fwd(_geom(zp3, core, fo), zp3, core, fo, None if bias is None else bias.contiguous(), y, v)
    return [y, t, zp3, wp, v, k]


@_trl_fwd_op.register_fake
def _(x, core, factors, bias):
    b, c = x.shape[0], x.shape[1]
    J = math.prod(f.shape[1] for f in factors[1:-1])
    S = math.prod(f.shape[0] for f in factors[1:-1])
    ldz = _r8(factors[0].shape[1])
    return [x.new_empty((b, factors[-1].shape[0])), x.new_empty((b, J, c)), x.new_empty((b, J, ldz)), x.new_empty((ldz, _r8(c))),
            x.new_empty((b, core.shape[-1]), dtype=torch.float32), x.new_empty((S, J))]


@torch.library.custom_op("tltorch::tucker_trl_bwd", mutates_args=())
def _trl_bwd_op(gy: torch.Tensor, x: torch.Tensor, t: torch.Tensor, zp: torch.Tensor, wp: torch.Tensor, v: torch.Tensor, k: torch.Tensor,
                core: torch.Tensor, factors: List[torch.Tensor], need_bias: bool) -> List[torch.Tensor]:
    """-> [dx, dcore, dbias | empty, dF_c, dF_s1.., dF_o]"""
    from . import spatial_ops
    from .blocktt_ops import _execute_gemm
    gy, x, core = gy.contiguous(), x.contiguous(), core.contiguous()
    fs = [f.contiguous() for f in factors]
    b, c = x.shape[0], x.shape[1]
    S, J = k.shape
    rc, ldz, cp = fs[0].shape[1], zp.shape[2], wp.shape[1]
    dev, bf, f32 = x.device, x.dtype, torch.float32
    fo = fs[-1]
    d = _geom(zp, core, fo)
    dv = torch.empty_like(v)
    dzp = torch.empty_like(zp)
    dcore, dfo = torch.empty_like(core), torch.empty_like(fo)
    dbias = torch.empty(fo.shape[0], dtype=bf, device=dev) if need_bias else None
    dt = torch.empty((b * J, c), dtype=bf, device=dev)
    dwp = torch.empty((rc, cp), dtype=f32, device=dev)               # dF_c^T: fp32 so that the GEMM may split the B J reduction
    dfc = torch.empty_like(fs[0])
    dx = torch.empty_like(x)
    dk32 = torch.zeros((S, J), dtype=f32, device=dev)
    _execute_bwd(d, zp, gy, core, fo, v, dv, dzp, None, None, None, phases=1)          # dv, dzp
    side, main = _fork(dev)

    def params():
        _execute_bwd(d, zp, gy, core, fo, v, dv, None, dcore, dfo, dbias, phases=2)    # dcore, dF_o, dbias
        _execute_gemm(dzp.reshape(b * J, ldz), t.reshape(b * J, c), None, dwp, M=rc, N=c, K=b * J, lda=ldz, ldb=c, a_major=1, b_major=1)
        dfc.copy_(dwp[:, :c].t())                                                      # (C, rc) bf16 <- (rc, C) fp32

    if side is not None:
        with torch.cuda.stream(side):
            params()
    # dt[b j, c] = sum_r dzp[b j, r] F_c[c, r]:  B operand (n = c, k = r) at wp[r * ld + c] -> MN-major
    _execute_gemm(dzp.reshape(b * J, ldz), wp, None, dt, M=b * J, N=c, K=ldz, lda=ldz, ldb=cp, a_major=0, b_major=1)
    spatial_ops._execute_spatial_bwd(x.reshape(b, c, S), k, dt.reshape(b, J, c), dx.reshape(b, c, S), dk32)
    # Kronecker chain backward, right to left
    spatial = fs[1:-1]
    chain, acc = [], spatial[0]                                      # (left operand, factor) of every product; the left operands
    for idx, m in enumerate(spatial[1:]):                            # beyond the first are intermediates, recomputed here
        chain.append((acc, m))
        if idx + 2 < len(spatial):
            nxt = torch.empty((acc.shape[0] * m.shape[0], acc.shape[1] * m.shape[1]), dtype=bf, device=dev)
            _execute_kron_fwd(acc, m, nxt)
            acc = nxt
    dk = dk32.to(bf)
    grads_s = [None] * len(spatial)
    for idx in range(len(chain) - 1, -1, -1):
        a, m = chain[idx]
        d_a, d_m = torch.empty_like(a), torch.empty_like(m)
        _execute_kron_bwd(a, m, dk, d_a, d_m)
        grads_s[idx + 1] = d_m
        dk = d_a
    grads_s[0] = dk
    if side is None:
        params()
    else:
        main.wait_stream(side)
    return [dx, dcore, dbias if need_bias else x.new_empty(0), dfc] + grads_s + [dfo]


@_trl_bwd_op.register_fake
def _(gy, x, t, zp, wp, v, k, core, factors, need_bias):
    return ([torch.empty_like(x), torch.empty_like(core), x.new_empty(factors[-1].shape[0] if need_bias else 0)]
            + [torch.empty_like(f) for f in factors])


def _trl_setup(ctx, inputs, output):
    x, core, factors, bias = inputs
    ctx.n = len(factors)
    ctx.has_bias = bias is not None
    ctx.save_for_backward(x, output[1], output[2], output[3], output[4], output[5], core, *factors)
    ctx.set_materialize_grads(False)


def _trl_backward(ctx, grads):
    if grads[0] is None:
        return None, None, None, None
    saved = list(ctx.saved_tensors)
    x, t, zp, wp, v, k, core = saved[:7]
    factors = saved[7:]
    need_bias = ctx.has_bias and ctx.needs_input_grad[3]
    outs = _trl_bwd_op(grads[0], x, t, zp, wp, v, k, core, factors, need_bias)
    return outs[0], outs[1], list(outs[3:3 + ctx.n]), (outs[2] if need_bias else None)


_trl_fwd_op.register_autograd(_trl_backward, setup_context=_trl_setup)


def tucker_trl_fused(x, weight, bias=None):
    """x (B, C, s_1..) bf16, weight a TuckerTensor of shape (C, s_1.., O) -> (B, O); differentiable in everything.  One op per
    direction: Kronecker product of the spatial factors -> spatial projection (one pass over the activation) -> channel-mode GEMM
    -> tail (core contraction + output expansion + bias); the backward pass runs the parameter gradients on a side stream."""
    return _trl_fwd_op(x, weight.core, list(weight.factors), bias)[0]
