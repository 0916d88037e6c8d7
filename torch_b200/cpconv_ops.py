"""Custom op (with autograd) for the fused CP convolution  y = F_out . dw3x3(F_in^T . x) (+ bias)  on NCHW bf16 activations.

  tltorch::cpconv3x3_fused(x (B, C, H, W), f_in (C, R), taps_t (9, R), f_out (O, R), bias (O) | None, flip, save, variant)
      -> [y (B, O, H, W), z1 (B, R, H, W) | empty, z2 (B, R, H, W) | empty]

One launch per direction (tlt_cpconv3x3_fused, torch_b200/csrc/cpconv_fused.cu): the R-channel intermediates of the chain
stay in shared memory.  The data gradient is the same launch with the two channel factors swapped and the taps flipped.
Reference: cp_conv / cp_conv_mobilenet (tltorch/functional/convolution.py:231-345) -- 1x1 conv -> depthwise convs -> 1x1
conv, each a full HBM round trip of the R-channel tensor -- and their autograd.
CPU tensors raise NotImplementedError (no fallback).
"""
import ctypes as C
import os
from typing import List, Optional

import torch

from . import _cuda, ops, pointwise_ops

__all__ = ["cp_conv_fused", "eligible"]


def _desc(x_shape, out_channels, rank, flip, bias, variant):
    d = _cuda.CpConvDesc()
    d.batch, d.in_channels, d.height, d.width = x_shape
    d.out_channels, d.rank = out_channels, rank
    d.flip_taps = 1 if flip else 0
    d.dtype_bias = _cuda.dtype_code(bias.dtype) if bias is not None else 0
    d.variant = variant
    return d


def _variant():
    return int(os.environ.get("TLT_CPF_VARIANT", "0"))


def eligible(x, f_in, taps_t, f_out, bias, kernel, stride, padding, dilation):
    """bf16 CUDA tensors, 3x3 / stride 1 / padding 1 / dilation 1, and a (C, O, W) the kernel is instantiated for -- in BOTH
    directions (the data gradient runs the same kernel with C and O swapped)."""
    # opt-in: on B200 the fused chain is issue-bound by the fp32 depthwise stage (profiles/r01_cpconv_fused.md) and does not yet
    # beat the unfused route, which stays the default
    if os.environ.get("TLT_CP_FUSED", "off") != "on":
        return False
    if x.dim() != 4 or not x.is_cuda or x.dtype != torch.bfloat16:
        return False
    if any(t.dtype != torch.bfloat16 for t in (f_in, taps_t, f_out)):
        return False
    if bias is not None and bias.dtype not in (torch.bfloat16, torch.float32):
        return False
    if list(kernel) != [3, 3] or list(stride) != [1, 1] or list(padding) != [1, 1] or list(dilation) != [1, 1]:
        return False
    lib = _cuda.lib()
    fwd = _desc(tuple(x.shape), f_out.shape[0], f_in.shape[1], False, bias, 0)
    bwd = _desc((x.shape[0], f_out.shape[0], x.shape[2], x.shape[3]), x.shape[1], f_in.shape[1], True, None, 0)
    return bool(lib.tlt_cpconv3x3_fused_supported(C.byref(fwd))) and bool(lib.tlt_cpconv3x3_fused_supported(C.byref(bwd)))


# =====================================================================================================================
# launch funnel
# =====================================================================================================================
def _execute_cpconv_fused(x, f_in, taps_t, f_out, bias, flip, y, z1, z2, variant):
    _cuda.require_cuda(x, f_in, taps_t, f_out, bias, y, z1, z2)
    d = _desc(tuple(x.shape), f_out.shape[0], f_in.shape[1], flip, bias, variant)
    _cuda.check(_cuda.lib().tlt_cpconv3x3_fused(C.byref(d), _cuda.ptr(x), _cuda.ptr(f_in), _cuda.ptr(taps_t), _cuda.ptr(f_out),
                                                _cuda.ptr(bias), _cuda.ptr(y), _cuda.ptr(z1), _cuda.ptr(z2),
                                                _cuda.stream_ptr(x)), "tlt_cpconv3x3_fused")


# =====================================================================================================================
# ops
# =====================================================================================================================
@torch.library.custom_op("tltorch::cpconv3x3_fused", mutates_args=())
def _cpf_op(x: torch.Tensor, f_in: torch.Tensor, taps_t: torch.Tensor, f_out: torch.Tensor, bias: Optional[torch.Tensor],
            flip: bool, save: bool, variant: int) -> List[torch.Tensor]:
    x, f_in, taps_t, f_out = x.contiguous(), f_in.contiguous(), taps_t.contiguous(), f_out.contiguous()
    B, _, H, W = x.shape
    R, O = f_in.shape[1], f_out.shape[0]
    y = torch.empty((B, O, H, W), dtype=x.dtype, device=x.device)
    z1 = torch.empty((B, R, H, W), dtype=x.dtype, device=x.device) if save else None
    z2 = torch.empty((B, R, H, W), dtype=x.dtype, device=x.device) if save else None
    _execute_cpconv_fused(x, f_in, taps_t, f_out, None if bias is None else bias.contiguous(), flip, y, z1, z2, variant)
    return [y, x.new_empty(0) if z1 is None else z1, x.new_empty(0) if z2 is None else z2]


@_cpf_op.register_fake
def _(x, f_in, taps_t, f_out, bias, flip, save, variant):
    B, _, H, W = x.shape
    R, O = f_in.shape[1], f_out.shape[0]
    z = (B, R, H, W) if save else (0,)
    return [x.new_empty((B, O, H, W)), x.new_empty(z), x.new_empty(z)]


def _cpf_setup(ctx, inputs, output):
    x, f_in, taps_t, f_out, bias, flip, save, variant = inputs
    ctx.has_bias = bias is not None
    ctx.flip, ctx.variant = flip, variant
    ctx.save_for_backward(x, f_in, taps_t, f_out, output[1], output[2])


def _cpf_backward(ctx, grads):
    gy = grads[0]
    x, f_in, taps_t, f_out, z1, z2 = ctx.saved_tensors
    if z1.numel() == 0:
        raise RuntimeError("tltorch::cpconv3x3_fused: backward needs the op to be called with save=True")
    gy = gy.contiguous()
    B, Cc, H, W = x.shape
    R, O = f_in.shape[1], f_out.shape[0]
    # data-gradient chain: gy -> dz2 = F_out^T gy -> dz1 = dw^T(dz2) -> dx = F_in dz1, one launch; its two intermediates
    # feed the three weight-gradient reductions
    dx, dz2, dz1 = _cpf_op(gy, f_out, taps_t, f_in, None, not ctx.flip, True, ctx.variant)
    need = ctx.needs_input_grad
    g_in = g_taps = g_out = g_bias = None
    gy3 = gy.reshape(B, O, H * W)
    if need[1]:
        g_in = pointwise_ops._pw_wgrad_op(x.reshape(B, Cc, H * W), dz1.reshape(B, R, H * W))          # (C, R)
    if need[2]:
        g = ops._dwconv_wgrad_op(z1, dz2, [3, 3], [1, 1], [1, 1], [1, 1]).reshape(R, 9)
        if ctx.flip:
            g = g.flip(1)
        g_taps = g.t().to(taps_t.dtype)                                                              # (9, R)
    if need[3]:
        g_out = pointwise_ops._pw_wgrad_op(gy3, z2.reshape(B, R, H * W))                              # (O, R)
    if ctx.has_bias and need[4]:
        g_bias = ops.reduce_to(gy3, "bns", "n")
    return (dx if need[0] else None), g_in, g_taps, g_out, g_bias, None, None, None


_cpf_op.register_autograd(_cpf_backward, setup_context=_cpf_setup)


def cp_conv_fused(x, f_in, taps_t, f_out, bias=None):
    """x (B, C, H, W) bf16; f_in (C, R); taps_t (9, R) merged depthwise kernel (lambda folded in here or into f_out);
    f_out (O, R).  Differentiable in everything."""
    need_grad = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x, f_in, taps_t, f_out, bias))
    return _cpf_op(x, f_in, taps_t, f_out, bias, False, need_grad, _variant())[0]
