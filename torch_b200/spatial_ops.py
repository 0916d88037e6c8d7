"""Custom op (with autograd) for the spatial-modes-first projection of channels-first bf16 feature maps.

  tltorch::spatial_project(x3 (B, C, S), k (S, J)) -> t (B, J, C)        t[b, j, c] = sum_s x3[b, c, s] k[s, j]

One pass over the activation where it lies in NCHW memory (tlt_spatial_project_fwd / _bwd, torch_b200/csrc/spatial_project.cu);
t comes out channels-last so the channel mode is a K-major tcgen05 GEMM.  Reference call sites: the input projection of
tucker_trl (tltorch/functional/tensor_regression.py:40) and TCL.forward (factorized_layers/tensor_contraction_layers.py:75).
CPU tensors raise NotImplementedError (no fallback).
"""
from typing import List

import torch

from . import _cuda

__all__ = ["spatial_project", "eligible"]


def eligible(x, S, J):
    if not x.is_cuda or x.dtype != torch.bfloat16 or not x.is_contiguous():
        return False
    return bool(_cuda.lib().tlt_spatial_project_supported(x.shape[0], x.shape[1], S, J))


def _execute_spatial_fwd(x3, k, t):
    _cuda.require_cuda(x3, k, t)
    B, Cc, S = x3.shape
    _cuda.check(_cuda.lib().tlt_spatial_project_fwd(B, Cc, S, k.shape[1], _cuda.ptr(x3), _cuda.ptr(k), _cuda.ptr(t),
                                                    _cuda.stream_ptr(x3)), "tlt_spatial_project_fwd")


def _execute_spatial_bwd(x3, k, gt, dx, dk):
    _cuda.require_cuda(x3, k, gt, dx, dk)
    B, Cc, S = x3.shape
    _cuda.check(_cuda.lib().tlt_spatial_project_bwd(B, Cc, S, k.shape[1], _cuda.ptr(x3), _cuda.ptr(k), _cuda.ptr(gt), _cuda.ptr(dx),
                                                    _cuda.ptr(dk), _cuda.stream_ptr(x3)), "tlt_spatial_project_bwd")


@torch.library.custom_op("tltorch::spatial_project", mutates_args=())
def _sp_op(x3: torch.Tensor, k: torch.Tensor) -> torch.Tensor:
    x3, k = x3.contiguous(), k.contiguous()
    B, Cc, S = x3.shape
    t = torch.empty((B, k.shape[1], Cc), dtype=x3.dtype, device=x3.device)
    _execute_spatial_fwd(x3, k, t)
    return t


@_sp_op.register_fake
def _(x3, k):
    return x3.new_empty((x3.shape[0], k.shape[1], x3.shape[1]))


@torch.library.custom_op("tltorch::spatial_project_bwd", mutates_args=())
def _sp_bwd_op(x3: torch.Tensor, k: torch.Tensor, gt: torch.Tensor) -> List[torch.Tensor]:
    x3, k, gt = x3.contiguous(), k.contiguous(), gt.contiguous()
    dx = torch.empty_like(x3)
    dk = torch.zeros(k.shape, dtype=torch.float32, device=x3.device)
    _execute_spatial_bwd(x3, k, gt, dx, dk)
    return [dx, dk]


@_sp_bwd_op.register_fake
def _(x3, k, gt):
    return [torch.empty_like(x3), x3.new_empty(k.shape, dtype=torch.float32)]


def _sp_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _sp_backward(ctx, gt):
    x3, k = ctx.saved_tensors
    dx, dk = _sp_bwd_op(x3, k, gt)
    return (dx if ctx.needs_input_grad[0] else None), (dk.to(k.dtype) if ctx.needs_input_grad[1] else None)


_sp_op.register_autograd(_sp_backward, setup_context=_sp_setup)


def spatial_project(x3, k):
    """x3 (B, C, S) bf16 contiguous, k (S, J) -> (B, J, C); differentiable in both."""
    return _sp_op(x3, k)
