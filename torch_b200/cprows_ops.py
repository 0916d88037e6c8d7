"""Custom ops (with autograd) for the op-level fused CP convolution on channels-last rows.

  tltorch::cpconv_rows_fwd(x (B, H, W, C), lambda (R), f_out (O, R), f_in (C, R), kh (3, R), kw (3, R), bias (O) | None, stride)
      -> [y (B, H/s, W/s, O), ws (packed operands, reused by the backward op)]
  tltorch::cpconv_rows_bwd(x, gy, ws, lambda, kh, kw, O)  -> [dx, g_lambda, g_f_out, g_f_in, g_kh, g_kw]

One launch per direction for the whole chain 1x1 -> depthwise 3x3 -> 1x1 (tlt_cpconv_rows_fwd / _bwd, torch_b200/csrc/cpconv_rows.cu):
the R-channel intermediates live in TMEM / shared memory, the weight gradients are accumulated on chip and flushed once per CTA.
Around them one operand-packing launch (forward) and one gradient-finalisation launch (backward); no Khatri-Rao launch (the
separable taps go in as they are), no transposition (the activation is channels-last), no cast / fill kernels.
Reference: cp_conv / cp_conv_mobilenet (tltorch/functional/convolution.py:231-345) and their autograd.
CPU tensors raise NotImplementedError (no fallback); tests/_emulator.py swaps the funnels.
"""
import ctypes as C
from typing import List, Optional

import torch

from . import _cuda

__all__ = ["cp_conv_rows_fused", "eligible"]


def _desc(batch, c_in, c_out, rank, height, width, stride):
    d = _cuda.CpConvRowsDesc()
    d.batch, d.in_channels, d.out_channels, d.rank, d.height, d.width, d.stride = batch, c_in, c_out, rank, height, width, stride
    return d


def supported(batch, c_in, c_out, rank, height, width, stride, training):
    """Shapes the kernels are instantiated for (tlt_cpconv_rows_supported); the stride-2 layer has no fused backward yet."""
    if width != 56 or height % 2 or rank > 128:
        return False
    if stride == 1:
        return (c_in, c_out) == (64, 64)
    if stride == 2:
        return (c_in, c_out) == (64, 128) and height % 8 == 0 and not training
    return False


def eligible(x, weights, factors, bias, kernel, stride, padding, dilation):
    """Channels-last bf16 activation, 3x3 / padding 1 / dilation 1, equal strides, a supported (C, O, W, R)."""
    if x.dim() != 4 or x.dtype != torch.bfloat16 or len(factors) != 4:     # (a CPU tensor fails loudly in the launch funnel)
        return False
    if any(t.dtype != torch.bfloat16 for t in [weights] + list(factors)):
        return False
    if bias is not None and bias.dtype not in (torch.bfloat16, torch.float32):
        return False
    if list(kernel) != [3, 3] or list(padding) != [1, 1] or list(dilation) != [1, 1] or stride[0] != stride[1]:
        return False
    b, c, h, w = x.shape
    if not x.permute(0, 2, 3, 1).is_contiguous():
        return False
    training = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in [x, weights, bias] + list(factors))
    return supported(b, c, factors[0].shape[0], factors[0].shape[1], h, w, stride[0], training)


# =====================================================================================================================
# launch funnels (one per C entry point)
# =====================================================================================================================
def _ws_bytes(d):
    return int(_cuda.lib().tlt_cpconv_rows_workspace_size(C.byref(d)))


def _acc_bytes(d):
    return int(_cuda.lib().tlt_cpconv_rows_grad_workspace_size(C.byref(d)))


def _execute_pack(d, f_in, f_out, kh, kw, lam, ws):
    _cuda.require_cuda(f_in, f_out, kh, kw, lam, ws)
    _cuda.check(_cuda.lib().tlt_cpconv_rows_pack(C.byref(d), _cuda.ptr(f_in), _cuda.ptr(f_out), _cuda.ptr(kh), _cuda.ptr(kw),
                                                 _cuda.ptr(lam), _cuda.ptr(ws), _cuda.stream_ptr(ws)), "tlt_cpconv_rows_pack")


def _execute_fwd(d, x, ws, bias, y):
    _cuda.require_cuda(x, ws, bias, y)
    _cuda.check(_cuda.lib().tlt_cpconv_rows_fwd(C.byref(d), _cuda.ptr(x), _cuda.ptr(ws), _cuda.ptr(bias), _cuda.ptr(y),
                                                _cuda.stream_ptr(x)), "tlt_cpconv_rows_fwd")


def _execute_bwd(d, x, gy, ws, dx, acc):
    _cuda.require_cuda(x, gy, ws, dx, acc)
    _cuda.check(_cuda.lib().tlt_cpconv_rows_bwd(C.byref(d), _cuda.ptr(x), _cuda.ptr(gy), _cuda.ptr(ws), _cuda.ptr(dx), _cuda.ptr(acc),
                                                _cuda.stream_ptr(x)), "tlt_cpconv_rows_bwd")


def _execute_finalize(d, acc, kh, kw, lam, g_fout, g_fin, g_kh, g_kw, g_lam):
    _cuda.require_cuda(acc, kh, kw, lam, g_fout, g_fin, g_kh, g_kw, g_lam)
    _cuda.check(_cuda.lib().tlt_cpconv_rows_finalize(C.byref(d), _cuda.ptr(acc), _cuda.ptr(kh), _cuda.ptr(kw), _cuda.ptr(lam),
                                                     _cuda.ptr(g_fout), _cuda.ptr(g_fin), _cuda.ptr(g_kh), _cuda.ptr(g_kw),
                                                     _cuda.ptr(g_lam), _cuda.stream_ptr(acc)), "tlt_cpconv_rows_finalize")


# =====================================================================================================================
# ops
# =====================================================================================================================
@torch.library.custom_op("tltorch::cpconv_rows_fwd", mutates_args=())
def _fwd_op(x: torch.Tensor, lam: torch.Tensor, f_out: torch.Tensor, f_in: torch.Tensor, kh: torch.Tensor, kw: torch.Tensor,
            bias: Optional[torch.Tensor], stride: int) -> List[torch.Tensor]:
    x = x.contiguous()
    b, h, w, c = x.shape
    o, r = f_out.shape
    d = _desc(b, c, o, r, h, w, stride)
    ws = torch.empty(_ws_bytes(d), dtype=torch.uint8, device=x.device)
    _execute_pack(d, f_in.contiguous(), f_out.contiguous(), kh.contiguous(), kw.contiguous(), lam.contiguous(), ws)
    y = torch.empty((b, h // stride, w // stride, o), dtype=x.dtype, device=x.device)
    _execute_fwd(d, x, ws, None if bias is None else bias.float().contiguous(), y)
    return [y, ws]


@_fwd_op.register_fake
def _(x, lam, f_out, f_in, kh, kw, bias, stride):
    b, h, w, c = x.shape
    return [x.new_empty((b, h // stride, w // stride, f_out.shape[0])), x.new_empty((0,), dtype=torch.uint8)]


@torch.library.custom_op("tltorch::cpconv_rows_bwd", mutates_args=())
def _bwd_op(x: torch.Tensor, gy: torch.Tensor, ws: torch.Tensor, lam: torch.Tensor, kh: torch.Tensor, kw: torch.Tensor,
            out_channels: int) -> List[torch.Tensor]:
    x, gy = x.contiguous(), gy.contiguous()
    b, h, w, c = x.shape
    o, r = out_channels, lam.shape[0]
    d = _desc(b, c, o, r, h, w, 1)
    dx = torch.empty_like(x)
    acc = torch.empty(_acc_bytes(d) // 4, dtype=torch.float32, device=x.device)
    _execute_bwd(d, x, gy, ws, dx, acc)
    g_lam, g_kh, g_kw = torch.empty_like(lam), torch.empty_like(kh), torch.empty_like(kw)
    g_fout = torch.empty((o, r), dtype=lam.dtype, device=x.device)
    g_fin = torch.empty((c, r), dtype=lam.dtype, device=x.device)
    _execute_finalize(d, acc, kh.contiguous(), kw.contiguous(), lam.contiguous(), g_fout, g_fin, g_kh, g_kw, g_lam)
    return [dx, g_lam, g_fout, g_fin, g_kh, g_kw]


@_bwd_op.register_fake
def _(x, gy, ws, lam, kh, kw, out_channels):
    r = lam.shape[0]
    return [torch.empty_like(x), torch.empty_like(lam), lam.new_empty((out_channels, r)), lam.new_empty((x.shape[3], r)),
            torch.empty_like(kh), torch.empty_like(kw)]


def _setup(ctx, inputs, output):
    x, lam, f_out, f_in, kh, kw, bias, stride = inputs
    ctx.has_bias = bias is not None
    ctx.bias_dtype = None if bias is None else bias.dtype
    ctx.stride, ctx.out_channels = stride, f_out.shape[0]
    ctx.save_for_backward(x, output[1], lam, kh, kw)
    ctx.set_materialize_grads(False)       # the workspace output never gets a gradient: no zero tensor for it


def _backward(ctx, grads):
    if grads[0] is None:
        return (None,) * 8
    gy = grads[0]
    x, ws, lam, kh, kw = ctx.saved_tensors
    if ctx.stride != 1:
        raise NotImplementedError("tltorch::cpconv_rows_fwd: the stride-2 layer has no fused backward (eligible() keeps training off it)")
    dx, g_lam, g_fout, g_fin, g_kh, g_kw = _bwd_op(x, gy, ws, lam, kh, kw, ctx.out_channels)
    g_bias = None
    if ctx.has_bias and ctx.needs_input_grad[6]:
        g_bias = gy.reshape(-1, gy.shape[-1]).sum(0, dtype=torch.float32).to(ctx.bias_dtype)
    return dx, g_lam, g_fout, g_fin, g_kh, g_kw, g_bias, None


_fwd_op.register_autograd(_backward, setup_context=_setup)


def cp_conv_rows_fused(x, weights, factors, bias, stride):
    """x (B, C, H, W) channels-last bf16 -> y (B, O, H/s, W/s) channels-last.  factors = [F_out (O, R), F_in (C, R), kh (3, R),
    kw (3, R)], weights = lambda (R): the CP tensor of the (O, C, 3, 3) kernel.  Differentiable in everything."""
    xr = x.permute(0, 2, 3, 1)                                      # the same memory as (B, H, W, C) rows
    y = _fwd_op(xr, weights, factors[0], factors[1], factors[2], factors[3], bias, int(stride))[0]
    return y.permute(0, 3, 1, 2)
