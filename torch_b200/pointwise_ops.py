"""Custom op (with autograd) for 1x1 channel mixing of channels-first bf16 activations on the tcgen05 tensor cores.

  tltorch::pointwise_tc(x (B, K, S), w_nk (N, K) any strides, bias (N) | None) -> y (B, N, S)

      y[b, n, s] = sum_k w_nk[n, k] * x[b, k, s] (+ bias[n])

Reference call sites: the F.conv1d(kernel_size=1) stages of cp_conv / cp_conv_mobilenet / tucker_conv / tt_conv
(tltorch/functional/convolution.py:154,168,207,223,262,278,311,340).  Forward and data gradient are tlt_pointwise_tc
(the data gradient with the transposed weight view), the weight gradient is tlt_pointwise_wgrad_tc (split over all SM
pairs, fp32 partial sums met by TMA reduce-add), the bias gradient a column-sum contraction.  Only the parameter-sized
weight matrix is repacked (K-major, pitch padded to 8) per call; activations are read where they lie in NCHW memory.
Eligibility (``eligible``): CUDA bf16 tensors with S % 8 == 0; everything else stays on the generic contraction.
The launch funnels below are what tests/_emulator.py swaps out.
"""
import ctypes as C
from typing import Optional, Tuple

import torch

from . import _cuda
from . import ops

__all__ = ["pointwise_tc", "eligible"]

MIN_WORK = 1 << 22          # below this many multiply-adds the generic kernel's single launch is as good


def eligible(x3, w_nk, bias):
    if x3.dtype != torch.bfloat16 or w_nk.dtype != torch.bfloat16:
        return False
    if bias is not None and bias.dtype not in (torch.bfloat16, torch.float32):
        return False
    B, K, S = x3.shape
    N = w_nk.shape[0]
    return S % 8 == 0 and S >= 8 and B >= 1 and B * S * K * N >= MIN_WORK


def _desc(B, K, N, S, w_ld, bias):
    d = _cuda.PointwiseDesc()
    d.batch, d.in_channels, d.out_channels, d.spatial, d.w_ld = B, K, N, S, w_ld
    d.dtype_bias = _cuda.dtype_code(bias.dtype) if bias is not None else 0
    return d


# =====================================================================================================================
# launch funnels
# =====================================================================================================================
def _execute_pointwise(x3, w_nk, bias, y):
    """y[b,n,s] = sum_k w_nk[n,k] x3[b,k,s] (+ bias[n]).  w_nk: (N, K) view with arbitrary strides."""
    _cuda.require_cuda(x3, w_nk, bias, y)
    B, K, S = x3.shape
    N = w_nk.shape[0]
    lib = _cuda.lib()
    ld = (K + 7) // 8 * 8
    wt = torch.empty((N, ld), dtype=torch.bfloat16, device=x3.device)
    st = _cuda.stream_ptr(x3)
    _cuda.check(lib.tlt_pack_kmajor(_cuda.ptr(w_nk), _cuda.dtype_code(w_nk.dtype), N, K, w_nk.stride(0), w_nk.stride(1),
                                    _cuda.ptr(wt), ld, st), "tlt_pack_kmajor")
    d = _desc(B, K, N, S, ld, bias)
    _cuda.check(lib.tlt_pointwise_tc(C.byref(d), _cuda.ptr(x3), _cuda.ptr(wt), _cuda.ptr(bias), _cuda.ptr(y), st),
                "tlt_pointwise_tc")


def _execute_pointwise_wgrad(gy3, x3, dw32):
    """dw32[n,k] = sum_{b,s} gy3[b,n,s] x3[b,k,s]; dw32 is an fp32 (N, ld) buffer with ld % 4 == 0."""
    _cuda.require_cuda(gy3, x3, dw32)
    B, K, S = x3.shape
    N = gy3.shape[1]
    d = _desc(B, K, N, S, dw32.stride(0), None)
    _cuda.check(_cuda.lib().tlt_pointwise_wgrad_tc(C.byref(d), _cuda.ptr(gy3), _cuda.ptr(x3), _cuda.ptr(dw32),
                                                   _cuda.stream_ptr(x3)), "tlt_pointwise_wgrad_tc")


# =====================================================================================================================
# ops
# =====================================================================================================================
@torch.library.custom_op("tltorch::pointwise_tc", mutates_args=())
def _pw_op(x3: torch.Tensor, w_nk: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    x3 = x3.contiguous()
    B, K, S = x3.shape
    N = w_nk.shape[0]
    y = torch.empty((B, N, S), dtype=x3.dtype, device=x3.device)
    _execute_pointwise(x3, w_nk, bias, y)
    return y


@_pw_op.register_fake
def _(x3, w_nk, bias):
    return x3.new_empty((x3.shape[0], w_nk.shape[0], x3.shape[2]))


@torch.library.custom_op("tltorch::pointwise_wgrad_tc", mutates_args=())
def _pw_wgrad_op(gy3: torch.Tensor, x3: torch.Tensor) -> torch.Tensor:
    gy3, x3 = gy3.contiguous(), x3.contiguous()
    N, K = gy3.shape[1], x3.shape[1]
    ld = (K + 3) // 4 * 4
    dw = torch.empty((N, ld), dtype=torch.float32, device=x3.device)
    _execute_pointwise_wgrad(gy3, x3, dw)
    return dw[:, :K].to(x3.dtype)


@_pw_wgrad_op.register_fake
def _(gy3, x3):
    return x3.new_empty((gy3.shape[1], x3.shape[1]))


def _pw_setup(ctx, inputs, output):
    x3, w_nk, bias = inputs
    ctx.save_for_backward(x3, w_nk)
    ctx.has_bias = bias is not None


def _pw_backward(ctx, gy):
    x3, w_nk = ctx.saved_tensors
    gx = gw = gb = None
    if ctx.needs_input_grad[0]:
        gx = _pw_op(gy, w_nk.t(), None)
    if ctx.needs_input_grad[1]:
        gw = _pw_wgrad_op(gy, x3)
    if ctx.has_bias and ctx.needs_input_grad[2]:
        gb = ops.reduce_to(gy, "bns", "n")
    return gx, gw, gb


_pw_op.register_autograd(_pw_backward, setup_context=_pw_setup)


def pointwise_tc(x3, w_nk, bias=None):
    return _pw_op(x3, w_nk, bias)
