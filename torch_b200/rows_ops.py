"""Custom ops (with autograd) for dense k x k convolutions as row GEMMs -- BASELINE config 4's Tucker convolution.

  tltorch::nchw_to_rows / rows_to_nchw   (B, C, *S) <-> (B * prod(S), round8(C)) channels-last rows (each other's backward)
  tltorch::im2col_rows / col2im_rows     unfold / fold of rows in 16-byte channel chunks (each other's backward)

``dense_conv_rows`` chains them with ``tucker_ops.padded_linear`` (tcgen05 GEMM on zero-padded pitches):
  x -> rows -> [1x1: GEMM] -> unfold -> [k x k core: GEMM over (tap, channel)] -> [1x1: GEMM + bias] -> NCHW.
Reference: tucker_conv (tltorch/functional/convolution.py:140-173); the same building blocks serve tt_conv (:176-228) and
reconstructed kernels (:16-53).  CPU tensors raise NotImplementedError (no fallback); tests/_emulator.py swaps the funnels.
"""
import ctypes as C
from typing import List, Optional

import os

import torch
import torch.nn.functional as F

from . import _cuda
from . import ops
from .tucker_ops import padded_linear, _r8, _prod

__all__ = ["nchw_to_rows", "rows_to_nchw", "im2col_rows", "tucker_conv_rows", "rows_eligible", "skinny_linear", "expand_core"]


def _conv_desc(batch, ld, in_size, out_size, kernel, stride, padding, dilation):
    d = _cuda.ConvDesc()
    d.dtype = _cuda.BF16
    d.ndim = len(in_size)
    d.batch, d.channels = batch, ld
    for i in range(len(in_size)):
        d.in_size[i], d.out_size[i] = in_size[i], out_size[i]
        d.kernel[i], d.stride[i], d.padding[i], d.dilation[i] = kernel[i], stride[i], padding[i], dilation[i]
    return d


# =====================================================================================================================
# launch funnels
# =====================================================================================================================
def _execute_nchw_to_rows(x3, rows):
    _cuda.require_cuda(x3, rows)
    b, c, s = x3.shape
    _cuda.check(_cuda.lib().tlt_nchw_to_rows(_cuda.ptr(x3), b, c, s, _cuda.ptr(rows), rows.stride(0), _cuda.stream_ptr(x3)),
                "tlt_nchw_to_rows")


def _execute_rows_to_nchw(rows, y3):
    _cuda.require_cuda(rows, y3)
    b, c, s = y3.shape
    _cuda.check(_cuda.lib().tlt_rows_to_nchw(_cuda.ptr(rows), b, c, s, rows.stride(0), _cuda.ptr(y3), _cuda.stream_ptr(y3)),
                "tlt_rows_to_nchw")


def _execute_unfold_rows(z, cols, batch, in_size, out_size, kernel, stride, padding, dilation, fold):
    """fold=False: cols <- unfold(z);  fold=True: z <- fold(cols)."""
    _cuda.require_cuda(z, cols)
    d = _conv_desc(batch, z.shape[1], in_size, out_size, kernel, stride, padding, dilation)
    lib = _cuda.lib()
    if fold:
        _cuda.check(lib.tlt_col2im_rows(C.byref(d), _cuda.ptr(cols), _cuda.ptr(z), _cuda.stream_ptr(z)), "tlt_col2im_rows")
    else:
        _cuda.check(lib.tlt_im2col_rows(C.byref(d), _cuda.ptr(z), _cuda.ptr(cols), _cuda.stream_ptr(z)), "tlt_im2col_rows")


# =====================================================================================================================
# layout ops
# =====================================================================================================================
@torch.library.custom_op("tltorch::nchw_to_rows", mutates_args=())
def _to_rows_op(x: torch.Tensor) -> torch.Tensor:
    b, c = x.shape[0], x.shape[1]
    x3 = x.contiguous().reshape(b, c, -1)
    rows = torch.empty((b * x3.shape[2], _r8(c)), dtype=x.dtype, device=x.device)
    if rows.numel():
        _execute_nchw_to_rows(x3, rows)
    return rows


@_to_rows_op.register_fake
def _(x):
    return x.new_empty((x.shape[0] * _prod(x.shape[2:]), _r8(x.shape[1])))


@torch.library.custom_op("tltorch::rows_to_nchw", mutates_args=())
def _to_nchw_op(rows: torch.Tensor, batch: int, channels: int, spatial: List[int]) -> torch.Tensor:
    y = torch.empty((batch, channels, *spatial), dtype=rows.dtype, device=rows.device)
    if y.numel():
        _execute_rows_to_nchw(rows.contiguous(), y.reshape(batch, channels, -1))
    return y


@_to_nchw_op.register_fake
def _(rows, batch, channels, spatial):
    return rows.new_empty((batch, channels, *spatial))


def _to_rows_setup(ctx, inputs, output):
    ctx.shape = tuple(inputs[0].shape)


def _to_rows_backward(ctx, g):
    return _to_nchw_op(g, ctx.shape[0], ctx.shape[1], list(ctx.shape[2:]))


def _to_nchw_setup(ctx, inputs, output):
    pass


def _to_nchw_backward(ctx, gy):
    return _to_rows_op(gy), None, None, None


_to_rows_op.register_autograd(_to_rows_backward, setup_context=_to_rows_setup)
_to_nchw_op.register_autograd(_to_nchw_backward, setup_context=_to_nchw_setup)


def is_channels_last(x):
    """True when (B, C, *S) ``x`` is stored channels-last (torch.channels_last / channels_last_3d strides) with a channel count
    whose rows are 16-byte aligned: its memory IS the (B * prod(S), C) rows matrix -- no transposition pass is needed."""
    if x.dim() < 3 or x.shape[1] % 8 != 0 or x.shape[1] == 1:
        return False
    perm = [0] + list(range(2, x.dim())) + [1]
    return x.permute(perm).is_contiguous() and not x.is_contiguous()


def nchw_to_rows(x):
    """(B, C, *S) bf16 -> (B * prod(S), round8(C)); pad channels are zero.  A channels-last tensor is viewed, not copied."""
    if is_channels_last(x):
        perm = [0] + list(range(2, x.dim())) + [1]
        return x.permute(perm).reshape(-1, x.shape[1])
    return _to_rows_op(x)


def rows_to_nchw(rows, batch, channels, spatial, channels_last=False):
    """rows (B * prod(S), round8(C)) -> (B, C, *S).  channels_last=True (and C % 8 == 0) returns the rows themselves viewed as a
    channels-last (B, C, *S) tensor: the output memory format follows the input's, as for torch's own convolutions."""
    if channels_last and channels % 8 == 0 and rows.shape[1] == channels and rows.is_contiguous():
        nd = len(spatial)
        return rows.view(batch, *spatial, channels).permute(0, nd + 1, *range(1, nd + 1))
    return _to_nchw_op(rows, batch, channels, list(spatial))


@torch.library.custom_op("tltorch::im2col_rows", mutates_args=())
def _unfold_op(z: torch.Tensor, batch: int, in_size: List[int], kernel: List[int], stride: List[int], padding: List[int],
               dilation: List[int]) -> torch.Tensor:
    out_size = ops.conv_out_size(in_size, kernel, stride, padding, dilation)
    cols = torch.empty((batch * _prod(out_size), _prod(kernel) * z.shape[1]), dtype=z.dtype, device=z.device)
    if cols.numel():
        _execute_unfold_rows(z.contiguous(), cols, batch, in_size, out_size, kernel, stride, padding, dilation, False)
    return cols


@_unfold_op.register_fake
def _(z, batch, in_size, kernel, stride, padding, dilation):
    out_size = ops.conv_out_size(in_size, kernel, stride, padding, dilation)
    return z.new_empty((batch * _prod(out_size), _prod(kernel) * z.shape[1]))


@torch.library.custom_op("tltorch::col2im_rows", mutates_args=())
def _fold_op(cols: torch.Tensor, batch: int, channels: int, in_size: List[int], kernel: List[int], stride: List[int],
             padding: List[int], dilation: List[int]) -> torch.Tensor:
    out_size = ops.conv_out_size(in_size, kernel, stride, padding, dilation)
    z = torch.empty((batch * _prod(in_size), channels), dtype=cols.dtype, device=cols.device)
    if z.numel():
        _execute_unfold_rows(z, cols.contiguous(), batch, in_size, out_size, kernel, stride, padding, dilation, True)
    return z


@_fold_op.register_fake
def _(cols, batch, channels, in_size, kernel, stride, padding, dilation):
    return cols.new_empty((batch * _prod(in_size), channels))


def _unfold_setup(ctx, inputs, output):
    z, ctx.batch, ctx.in_size, ctx.kernel, ctx.stride, ctx.padding, ctx.dilation = inputs
    ctx.channels = z.shape[1]


def _unfold_backward(ctx, gcols):
    return (_fold_op(gcols, ctx.batch, ctx.channels, ctx.in_size, ctx.kernel, ctx.stride, ctx.padding, ctx.dilation),
            None, None, None, None, None, None)


def _fold_setup(ctx, inputs, output):
    _, ctx.batch, ctx.channels, ctx.in_size, ctx.kernel, ctx.stride, ctx.padding, ctx.dilation = inputs


def _fold_backward(ctx, gz):
    return (_unfold_op(gz, ctx.batch, ctx.in_size, ctx.kernel, ctx.stride, ctx.padding, ctx.dilation),
            None, None, None, None, None, None, None)


_unfold_op.register_autograd(_unfold_backward, setup_context=_unfold_setup)
_fold_op.register_autograd(_fold_backward, setup_context=_fold_setup)


def im2col_rows(z, batch, in_size, kernel, stride, padding, dilation):
    """z (B * prod(in_size), Cp) -> cols (B * prod(out_size), prod(kernel) * Cp), taps outer / channels inner."""
    return _unfold_op(z, batch, list(in_size), list(kernel), list(stride), list(padding), list(dilation))


# =====================================================================================================================
# Tucker convolution on rows
# =====================================================================================================================
MIN_CHANNELS = 32


def rows_eligible(x, tensors, order):
    if x.dtype != torch.bfloat16 or any(t is not None and t.dtype != torch.bfloat16 for t in tensors):
        return False
    return 1 <= order <= 3 and 1 <= x.shape[0] <= 65535 and x.numel() > 0


def tucker_conv_rows(x, w, f_in, f_out, bias, stride, padding, dilation):
    """x (B, C_in, *S) bf16; w (r0, r1, *k) the core already expanded over its kernel modes; f_in (C_in, r1); f_out (C_out, r0)."""
    b, spatial = x.shape[0], list(x.shape[2:])
    r0, r1 = w.shape[0], w.shape[1]
    kernel = list(w.shape[2:])
    taps = _prod(kernel)
    out_size = ops.conv_out_size(spatial, kernel, list(stride), list(padding), list(dilation))
    xt = nchw_to_rows(x)                                                       # (B S, round8(C_in))
    z = padded_linear(xt, f_in.t(), None, True)                                # (B S, round8(r1)), pad columns zero
    cols = im2col_rows(z, b, spatial, kernel, stride, padding, dilation)       # (B S', taps * round8(r1))
    wc = F.pad(w.reshape(r0, r1, taps).permute(0, 2, 1), (0, _r8(r1) - r1))    # (r0, taps, round8(r1)): K index = (tap, channel)
    h = padded_linear(cols, wc.reshape(r0, taps * _r8(r1)), None, True)        # (B S', round8(r0))
    yt = padded_linear(h, f_out, bias, True)                                   # (B S', round8(C_out))
    return rows_to_nchw(yt, b, f_out.shape[0], out_size, channels_last=is_channels_last(x))


# =====================================================================================================================
# tltorch::skinny_linear -- core expansion  w[r, :] = K a[r, :]  for tiny K (prod(kernel) x prod(kernel ranks))
# =====================================================================================================================
def _execute_skinny_fwd(a2, k2, y2):
    _cuda.require_cuda(a2, k2, y2)
    _cuda.check(_cuda.lib().tlt_skinny_fwd(_cuda.dtype_code(a2.dtype), a2.shape[0], a2.shape[1], k2.shape[0], _cuda.ptr(a2),
                                           _cuda.ptr(k2), _cuda.ptr(y2), _cuda.stream_ptr(a2)), "tlt_skinny_fwd")


def _execute_skinny_bwd(a2, g2, k2, da2, dk32):
    _cuda.require_cuda(a2, g2, k2, da2, dk32)
    _cuda.check(_cuda.lib().tlt_skinny_bwd(_cuda.dtype_code(g2.dtype), g2.shape[0], k2.shape[1], k2.shape[0], _cuda.ptr(a2),
                                           _cuda.ptr(g2), _cuda.ptr(k2), _cuda.ptr(da2), _cuda.ptr(dk32), _cuda.stream_ptr(g2)),
                "tlt_skinny_bwd")


@torch.library.custom_op("tltorch::skinny_linear", mutates_args=())
def _skinny_op(a2: torch.Tensor, k2: torch.Tensor) -> torch.Tensor:
    y2 = torch.empty((a2.shape[0], k2.shape[0]), dtype=a2.dtype, device=a2.device)
    if y2.numel():
        _execute_skinny_fwd(a2.contiguous(), k2.contiguous(), y2)
    return y2


@_skinny_op.register_fake
def _(a2, k2):
    return a2.new_empty((a2.shape[0], k2.shape[0]))


@torch.library.custom_op("tltorch::skinny_linear_bwd", mutates_args=())
def _skinny_bwd_op(a2: torch.Tensor, g2: torch.Tensor, k2: torch.Tensor, need_da: bool, need_dk: bool) -> List[torch.Tensor]:
    da = torch.empty_like(a2) if need_da else a2.new_empty(0)
    dk = torch.zeros(k2.shape, dtype=torch.float32, device=a2.device) if need_dk else a2.new_empty(0, dtype=torch.float32)
    if a2.shape[0] and (need_da or need_dk):
        _execute_skinny_bwd(a2.contiguous(), g2.contiguous(), k2.contiguous(), da if need_da else None, dk if need_dk else None)
    return [da, dk]


@_skinny_bwd_op.register_fake
def _(a2, g2, k2, need_da, need_dk):
    return [torch.empty_like(a2) if need_da else a2.new_empty(0),
            a2.new_empty(k2.shape, dtype=torch.float32) if need_dk else a2.new_empty(0, dtype=torch.float32)]


def _skinny_setup(ctx, inputs, output):
    ctx.save_for_backward(*inputs)


def _skinny_backward(ctx, g):
    a2, k2 = ctx.saved_tensors
    need_da, need_dk = ctx.needs_input_grad
    da, dk = _skinny_bwd_op(a2, g, k2, need_da, need_dk)
    return (da if need_da else None), (dk.to(k2.dtype) if need_dk else None)


_skinny_op.register_autograd(_skinny_backward, setup_context=_skinny_setup)

SKINNY_MAX = 64


def skinny_linear(a2, k2):
    """a2 (rows, I) x k2 (J, I)^T -> (rows, J) for I, J <= SKINNY_MAX; differentiable in both."""
    return _skinny_op(a2, k2)


def expand_core(core, kernel_factors):
    """core (r0, r1, q_1..q_d) x_2 F_1 ... x_{d+1} F_d with F_j (k_j, q_j)  ->  (r0, r1, k_1..k_d) in ONE launch:
    rows (r0 r1) times the Kronecker product of the factors (reference: tenalg.multi_mode_dot(core, factors[2:]),
    tltorch/functional/convolution.py:160).  None when the extents exceed the skinny kernel."""
    q = [f.shape[1] for f in kernel_factors]
    k = [f.shape[0] for f in kernel_factors]
    if _prod(q) > SKINNY_MAX or _prod(k) > SKINNY_MAX or core.dtype not in (torch.float32, torch.bfloat16):
        return None
    kron = kernel_factors[0]
    for f in kernel_factors[1:]:                                             # (K, Q) (x) (k, q) -> (K k, Q q), tiny
        if kron.dtype == torch.bfloat16 and f.dtype == torch.bfloat16:
            from .trl_ops import kron2                                      # two-kernel Kronecker op (csrc/trl_tail.cu)
            kron = kron2(kron, f)
        else:
            kron = ops.contract("ac,bd->abcd", kron, f).reshape(kron.shape[0] * f.shape[0], kron.shape[1] * f.shape[1])
    r0, r1 = core.shape[0], core.shape[1]
    return _skinny_op(core.reshape(r0 * r1, _prod(q)), kron).reshape(r0, r1, *k)


# =====================================================================================================================
# tltorch::scale_modes -- tensor dropout's mask application in fixed-shape form
# =====================================================================================================================
def _execute_scale_modes(x, masks, alpha, y):
    _cuda.require_cuda(x, y, *masks)
    nd = x.dim()
    dims = (C.c_int64 * nd)(*x.shape)
    ptrs = (C.c_void_p * nd)(*[None if m is None else m.data_ptr() for m in masks])
    _cuda.check(_cuda.lib().tlt_scale_modes(_cuda.dtype_code(x.dtype), nd, dims, _cuda.ptr(x), ptrs, float(alpha), _cuda.ptr(y),
                                            _cuda.stream_ptr(x)), "tlt_scale_modes")


@torch.library.custom_op("tltorch::scale_modes", mutates_args=())
def _scale_modes_op(x: torch.Tensor, masks: List[torch.Tensor], modes: List[int], alpha: float) -> torch.Tensor:
    x = x.contiguous()
    y = torch.empty_like(x)
    if x.numel():
        per_mode = [None] * x.dim()
        for m, mode in zip(masks, modes):
            per_mode[mode] = m.contiguous()
        _execute_scale_modes(x, per_mode, alpha, y)
    return y


@_scale_modes_op.register_fake
def _(x, masks, modes, alpha):
    return torch.empty_like(x)


def _scale_modes_setup(ctx, inputs, output):
    _, masks, ctx.modes, ctx.alpha = inputs
    ctx.save_for_backward(*masks)


def _scale_modes_backward(ctx, g):
    masks = list(ctx.saved_tensors)
    return _scale_modes_op(g, masks, ctx.modes, ctx.alpha), [None] * len(masks), None, None


_scale_modes_op.register_autograd(_scale_modes_backward, setup_context=_scale_modes_setup)


def scale_modes(x, masks, alpha=1.0):
    """x * alpha * prod_k masks[k][i_k] in one launch; ``masks`` has one entry per mode of x (None = untouched).  The masks
    are constants for autograd (dropout keep vectors)."""
    if x.dim() > 6 or x.dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("scale_modes: 1..6 modes, float32 / bfloat16")
    modes = [k for k, m in enumerate(masks) if m is not None]
    return _scale_modes_op(x, [masks[k].to(x.dtype) for k in modes], modes, float(alpha))


# =====================================================================================================================
# tltorch::dwconv_rows -- depthwise conv on channels-last rows; CP convolution on rows
# =====================================================================================================================
def _execute_dwconv_rows(kind, a, b, out, batch, in_size, out_size, kernel, stride, padding, dilation):
    """kind 'fwd': out <- conv(a = z, b = w);  'dgrad': out <- conv^T(a = gy, b = w);  'wgrad': out (fp32, +=) <- corr(a = z, b = gy)."""
    _cuda.require_cuda(a, b, out)
    ld = a.shape[1]
    d = _conv_desc(batch, ld, in_size, out_size, kernel, stride, padding, dilation)
    lib = _cuda.lib()
    fn = {"fwd": lib.tlt_dwconv_rows_fwd, "dgrad": lib.tlt_dwconv_rows_dgrad, "wgrad": lib.tlt_dwconv_rows_wgrad}[kind]
    _cuda.check(fn(C.byref(d), _cuda.ptr(a), _cuda.ptr(b), _cuda.ptr(out), _cuda.stream_ptr(a)), f"tlt_dwconv_rows_{kind}")


@torch.library.custom_op("tltorch::dwconv_rows", mutates_args=())
def _dwrows_op(z: torch.Tensor, w: torch.Tensor, batch: int, in_size: List[int], kernel: List[int], stride: List[int],
               padding: List[int], dilation: List[int]) -> torch.Tensor:
    out_size = ops.conv_out_size(in_size, kernel, stride, padding, dilation)
    y = torch.empty((batch * _prod(out_size), z.shape[1]), dtype=z.dtype, device=z.device)
    if y.numel():
        _execute_dwconv_rows("fwd", z.contiguous(), w.contiguous(), y, batch, in_size, out_size, kernel, stride, padding, dilation)
    return y


@_dwrows_op.register_fake
def _(z, w, batch, in_size, kernel, stride, padding, dilation):
    out_size = ops.conv_out_size(in_size, kernel, stride, padding, dilation)
    return z.new_empty((batch * _prod(out_size), z.shape[1]))


@torch.library.custom_op("tltorch::dwconv_rows_bwd", mutates_args=())
def _dwrows_bwd_op(gy: torch.Tensor, z: torch.Tensor, w: torch.Tensor, batch: int, in_size: List[int], kernel: List[int],
                   stride: List[int], padding: List[int], dilation: List[int], need_dz: bool, need_dw: bool) -> List[torch.Tensor]:
    out_size = ops.conv_out_size(in_size, kernel, stride, padding, dilation)
    gy, z, w = gy.contiguous(), z.contiguous(), w.contiguous()
    dz = torch.empty_like(z) if need_dz else z.new_empty(0)
    dw = torch.zeros(w.shape, dtype=torch.float32, device=z.device) if need_dw else z.new_empty(0, dtype=torch.float32)
    if z.numel() and gy.numel():
        if need_dz:
            _execute_dwconv_rows("dgrad", gy, w, dz, batch, in_size, out_size, kernel, stride, padding, dilation)
        if need_dw:
            _execute_dwconv_rows("wgrad", z, gy, dw, batch, in_size, out_size, kernel, stride, padding, dilation)
    elif need_dz:
        dz.zero_()
    return [dz, dw]


@_dwrows_bwd_op.register_fake
def _(gy, z, w, batch, in_size, kernel, stride, padding, dilation, need_dz, need_dw):
    return [torch.empty_like(z) if need_dz else z.new_empty(0),
            z.new_empty(w.shape, dtype=torch.float32) if need_dw else z.new_empty(0, dtype=torch.float32)]


def _dwrows_setup(ctx, inputs, output):
    z, w, ctx.batch, ctx.in_size, ctx.kernel, ctx.stride, ctx.padding, ctx.dilation = inputs
    ctx.save_for_backward(z, w)


def _dwrows_backward(ctx, gy):
    z, w = ctx.saved_tensors
    need_dz, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    dz, dw = _dwrows_bwd_op(gy, z, w, ctx.batch, ctx.in_size, ctx.kernel, ctx.stride, ctx.padding, ctx.dilation, need_dz, need_dw)
    return (dz if need_dz else None), (dw.to(w.dtype) if need_dw else None), None, None, None, None, None, None


_dwrows_op.register_autograd(_dwrows_backward, setup_context=_dwrows_setup)

DW_ROWS_MAX_TAPS = 9


# =====================================================================================================================
# separable 3x3 depthwise on rows (tlt_dwsep_rows_*): both passes in one sweep, the CP taps go in unmerged
# =====================================================================================================================
def _dwsep_desc(batch, height, width, ld, stride):
    d = _cuda.DwSepDesc()
    d.batch, d.height, d.width, d.ld, d.stride = batch, height, width, ld, stride
    return d


def _execute_dwsep_taps(kh, kw, lam, ld, flip, taps):
    _cuda.require_cuda(kh, kw, lam, taps)
    _cuda.check(_cuda.lib().tlt_dwsep_rows_taps(_cuda.ptr(kh), _cuda.ptr(kw), _cuda.ptr(lam), kh.shape[1], ld, 1 if flip else 0,
                                                _cuda.ptr(taps), _cuda.stream_ptr(taps)), "tlt_dwsep_rows_taps")


def _execute_dwsep(kind, a, b, out, batch, height, width, stride):
    """kind 'fwd': out = sweep(a; taps b); 'dgrad': out = sweep^T(a; taps b); 'wgrad': out (9, ld) fp32 = sum a(+tap) * b."""
    _cuda.require_cuda(a, b, out)
    d = _dwsep_desc(batch, height, width, a.shape[1], stride)
    fn = {"fwd": "tlt_dwsep_rows_fwd", "dgrad": "tlt_dwsep_rows_dgrad", "wgrad": "tlt_dwsep_rows_wgrad"}[kind]
    _cuda.check(getattr(_cuda.lib(), fn)(C.byref(d), _cuda.ptr(a), _cuda.ptr(b), _cuda.ptr(out), _cuda.stream_ptr(a)), fn)


def _dwsep_parts():
    """Number of per-block partial-sum slabs the weight-gradient kernel writes (tlt_dwsep_rows_wgrad_parts)."""
    return int(_cuda.lib().tlt_dwsep_rows_wgrad_parts())


def _execute_dwsep_finalize(acc, kh, kw, lam, g_kh, g_kw, g_lam):
    _cuda.require_cuda(acc, kh, kw, lam, g_kh, g_kw, g_lam)
    _cuda.check(_cuda.lib().tlt_dwsep_rows_finalize(_cuda.ptr(acc), acc.shape[0], _cuda.ptr(kh), _cuda.ptr(kw), _cuda.ptr(lam),
                                                    kh.shape[1], acc.shape[2], _cuda.ptr(g_kh), _cuda.ptr(g_kw), _cuda.ptr(g_lam),
                                                    _cuda.stream_ptr(acc)), "tlt_dwsep_rows_finalize")


@torch.library.custom_op("tltorch::dwsep_rows", mutates_args=())
def _dwsep_op(z: torch.Tensor, kh: torch.Tensor, kw: torch.Tensor, lam: torch.Tensor, batch: int, height: int, width: int,
              stride: int) -> torch.Tensor:
    z, kh, kw, lam = z.contiguous(), kh.contiguous(), kw.contiguous(), lam.contiguous()
    taps = torch.empty((6, z.shape[1]), dtype=torch.float32, device=z.device)
    _execute_dwsep_taps(kh, kw, lam, z.shape[1], False, taps)
    y = torch.empty((batch * (height // stride) * (width // stride), z.shape[1]), dtype=z.dtype, device=z.device)
    _execute_dwsep("fwd", z, taps, y, batch, height, width, stride)
    return y


@_dwsep_op.register_fake
def _(z, kh, kw, lam, batch, height, width, stride):
    return z.new_empty((batch * (height // stride) * (width // stride), z.shape[1]))


@torch.library.custom_op("tltorch::dwsep_rows_bwd", mutates_args=())
def _dwsep_bwd_op(gy: torch.Tensor, z: torch.Tensor, kh: torch.Tensor, kw: torch.Tensor, lam: torch.Tensor, batch: int, height: int,
                  width: int, stride: int) -> List[torch.Tensor]:
    gy, z, kh, kw, lam = gy.contiguous(), z.contiguous(), kh.contiguous(), kw.contiguous(), lam.contiguous()
    ld = z.shape[1]
    taps = torch.empty((6, ld), dtype=torch.float32, device=z.device)
    _execute_dwsep_taps(kh, kw, lam, ld, stride == 1, taps)      # stride 1: the same sweep with reversed taps
    dz = torch.empty_like(z)
    _execute_dwsep("dgrad", gy, taps, dz, batch, height, width, stride)
    acc = torch.empty((_dwsep_parts(), 9, ld), dtype=torch.float32, device=z.device)
    _execute_dwsep("wgrad", z, gy, acc, batch, height, width, stride)
    g_kh, g_kw, g_lam = torch.empty_like(kh), torch.empty_like(kw), torch.empty_like(lam)
    _execute_dwsep_finalize(acc, kh, kw, lam, g_kh, g_kw, g_lam)
    return [dz, g_kh, g_kw, g_lam]


@_dwsep_bwd_op.register_fake
def _(gy, z, kh, kw, lam, batch, height, width, stride):
    return [torch.empty_like(z), torch.empty_like(kh), torch.empty_like(kw), torch.empty_like(lam)]


def _dwsep_setup(ctx, inputs, output):
    z, kh, kw, lam, ctx.batch, ctx.height, ctx.width, ctx.stride = inputs
    ctx.save_for_backward(z, kh, kw, lam)


def _dwsep_backward(ctx, gy):
    z, kh, kw, lam = ctx.saved_tensors
    dz, g_kh, g_kw, g_lam = _dwsep_bwd_op(gy, z, kh, kw, lam, ctx.batch, ctx.height, ctx.width, ctx.stride)
    return dz, g_kh, g_kw, g_lam, None, None, None, None


_dwsep_op.register_autograd(_dwsep_backward, setup_context=_dwsep_setup)


def dwsep_eligible(factors, kernel, stride, padding, dilation, spatial):
    """2-D 3x3, padding 1, dilation 1, equal strides 1 or 2 (even maps), rank small enough for the kernel's shared accumulators."""
    if len(factors) != 4 or list(kernel) != [3, 3] or list(padding) != [1, 1] or list(dilation) != [1, 1]:
        return False
    if stride[0] != stride[1] or stride[0] not in (1, 2) or spatial[1] < 2 or _r8(factors[2].shape[1]) * 36 > 48 * 1024:
        return False
    return stride[0] == 1 or (spatial[0] % 2 == 0 and spatial[1] % 2 == 0)


_BRANCHES = os.environ.get("TLT_CHAIN_BRANCHES", "1") != "0"
def _side_stream(dev):
    """The side stream of the weight-gradient branch of the chain's backward op (torch_b200/streams.py)."""
    from . import streams
    return streams.side_stream(dev)


def _execute_chain_prep(c_in, c_out, rank, flip_d, f_in, f_out, kh, kw, lam, wp_in, wp_out, taps_f, taps_d):
    _cuda.require_cuda(f_in, f_out, kh, kw, lam, wp_in, wp_out, taps_f, taps_d)
    _cuda.check(_cuda.lib().tlt_cpchain_prep(c_in, c_out, rank, 1 if flip_d else 0, _cuda.ptr(f_in), _cuda.ptr(f_out), _cuda.ptr(kh),
                                             _cuda.ptr(kw), _cuda.ptr(lam), _cuda.ptr(wp_in), _cuda.ptr(wp_out), _cuda.ptr(taps_f),
                                             _cuda.ptr(taps_d), _cuda.stream_ptr(wp_in)), "tlt_cpchain_prep")


def _execute_chain_post(c_in, c_out, rank, dw_out, dw_in, acc, kh, kw, lam, g_fout, g_fin, g_kh, g_kw, g_lam):
    _cuda.require_cuda(dw_out, dw_in, acc, kh, kw, lam, g_fout, g_fin, g_kh, g_kw, g_lam)
    _cuda.check(_cuda.lib().tlt_cpchain_post(c_in, c_out, rank, _cuda.ptr(dw_out), _cuda.ptr(dw_in), _cuda.ptr(acc), acc.shape[0],
                                             _cuda.ptr(kh), _cuda.ptr(kw), _cuda.ptr(lam), _cuda.ptr(g_fout), _cuda.ptr(g_fin),
                                             _cuda.ptr(g_kh), _cuda.ptr(g_kw), _cuda.ptr(g_lam), _cuda.stream_ptr(acc)), "tlt_cpchain_post")


@torch.library.custom_op("tltorch::cpconv_chain_fwd", mutates_args=())
def _chain_fwd_op(xr: torch.Tensor, lam: torch.Tensor, f_out: torch.Tensor, f_in: torch.Tensor, kh: torch.Tensor, kw: torch.Tensor,
                  bias: Optional[torch.Tensor], batch: int, height: int, width: int, stride: int) -> List[torch.Tensor]:
    """xr (B H W, round8(C_in)) -> [y (B H' W', round8(C_out)), z1, z2, wp_in, wp_out, taps_d]: the unfused chain as ONE op --
    prep launch, GEMM, separable sweep, GEMM (4 launches); everything the backward op needs comes back as outputs."""
    from .blocktt_ops import _execute_gemm
    xr = xr.contiguous()
    (c_in, r), c_out = f_in.shape, f_out.shape[0]
    cin_p, cout_p, ld = _r8(c_in), _r8(c_out), _r8(r)
    dev, bf = xr.device, torch.bfloat16
    wp_in, wp_out = torch.empty((ld, cin_p), dtype=bf, device=dev), torch.empty((cout_p, ld), dtype=bf, device=dev)
    taps_f, taps_d = torch.empty((6, ld), dtype=torch.float32, device=dev), torch.empty((6, ld), dtype=torch.float32, device=dev)
    _execute_chain_prep(c_in, c_out, r, stride == 1, f_in.contiguous(), f_out.contiguous(), kh.contiguous(), kw.contiguous(),
                        lam.contiguous(), wp_in, wp_out, taps_f, taps_d)
    p_in, p_out = xr.shape[0], batch * (height // stride) * (width // stride)
    z1 = torch.empty((p_in, ld), dtype=bf, device=dev)
    _execute_gemm(xr, wp_in, None, z1, M=p_in, N=ld, K=cin_p, lda=xr.stride(0), ldb=cin_p, a_major=0, b_major=0)
    z2 = torch.empty((p_out, ld), dtype=bf, device=dev)
    _execute_dwsep("fwd", z1, taps_f, z2, batch, height, width, stride)
    y = torch.empty((p_out, cout_p), dtype=bf, device=dev)
    if bias is not None and cout_p != c_out:
        bias = torch.cat([bias, bias.new_zeros(cout_p - c_out)])
    _execute_gemm(z2, wp_out, bias, y, M=p_out, N=cout_p, K=ld, lda=ld, ldb=ld, a_major=0, b_major=0)
    return [y, z1, z2, wp_in, wp_out, taps_d]


@_chain_fwd_op.register_fake
def _(xr, lam, f_out, f_in, kh, kw, bias, batch, height, width, stride):
    (c_in, r), c_out = f_in.shape, f_out.shape[0]
    cin_p, cout_p, ld = _r8(c_in), _r8(c_out), _r8(r)
    p_out = batch * (height // stride) * (width // stride)
    return [xr.new_empty((p_out, cout_p)), xr.new_empty((xr.shape[0], ld)), xr.new_empty((p_out, ld)), xr.new_empty((ld, cin_p)),
            xr.new_empty((cout_p, ld)), xr.new_empty((6, ld), dtype=torch.float32)]


@torch.library.custom_op("tltorch::cpconv_chain_bwd", mutates_args=())
def _chain_bwd_op(gy: torch.Tensor, xr: torch.Tensor, z1: torch.Tensor, z2: torch.Tensor, wp_in: torch.Tensor, wp_out: torch.Tensor,
                  taps_d: torch.Tensor, lam: torch.Tensor, kh: torch.Tensor, kw: torch.Tensor, c_in: int, c_out: int, batch: int,
                  height: int, width: int, stride: int) -> List[torch.Tensor]:
    """-> [dx, g_lambda, g_f_out, g_f_in, g_kh, g_kw]: 2 data-gradient GEMMs + transposed sweep, sweep + 2 GEMMs for the weight
    gradients, one finalisation launch (7 launches)."""
    from .blocktt_ops import _execute_gemm
    gy, xr = gy.contiguous(), xr.contiguous()
    r = lam.shape[0]
    cin_p, cout_p, ld = wp_in.shape[1], wp_out.shape[0], wp_in.shape[0]
    dev, bf, f32 = xr.device, torch.bfloat16, torch.float32
    p_in, p_out = xr.shape[0], gy.shape[0]
    dz2 = torch.empty((p_out, ld), dtype=bf, device=dev)                 # dz2[b, r] = sum_o gy[b, o] F_out[o, r]
    dz1 = torch.empty((p_in, ld), dtype=bf, device=dev)
    dx = torch.empty((p_in, cin_p), dtype=bf, device=dev)                # dx[b, c] = sum_r dz1[b, r] F_in[c, r]
    acc = torch.empty((_dwsep_parts(), 9, ld), dtype=f32, device=dev)
    dw_out = torch.empty((c_out, ld), dtype=f32, device=dev)             # dW_out[o, r] = sum_b gy[b, o] z2[b, r]   (split-K, fp32)
    dw_in = torch.empty((c_in, ld), dtype=f32, device=dev)               # dW_in[c, r] = sum_b x[b, c] dz1[b, r]   (F_in's own layout)
    _execute_gemm(gy, wp_out, None, dz2, M=p_out, N=ld, K=cout_p, lda=gy.stride(0), ldb=ld, a_major=0, b_major=1)
    # two branches from here: [transposed sweep -> dx GEMM -> dW_in GEMM] on the current stream, [weight-gradient sweep -> dW_out GEMM]
    # on a side stream (every buffer is allocated above, on the current stream, and outlives the join).  On the small maps
    # (14^2, 7^2) each of these launches is a few waves of work: the second branch fills the tails of the first.  Captured into a
    # CUDA graph the fork / join are plain graph edges.
    side = _side_stream(dev) if _BRANCHES and dev.type == "cuda" else None
    if side is not None:
        main = torch.cuda.current_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _execute_dwsep("wgrad", z1, dz2, acc, batch, height, width, stride)
            _execute_gemm(gy, z2, None, dw_out, M=c_out, N=ld, K=p_out, lda=gy.stride(0), ldb=ld, a_major=1, b_major=1)
    _execute_dwsep("dgrad", dz2, taps_d, dz1, batch, height, width, stride)
    _execute_gemm(dz1, wp_in, None, dx, M=p_in, N=cin_p, K=ld, lda=ld, ldb=cin_p, a_major=0, b_major=1)
    if side is None:
        _execute_dwsep("wgrad", z1, dz2, acc, batch, height, width, stride)
        _execute_gemm(gy, z2, None, dw_out, M=c_out, N=ld, K=p_out, lda=gy.stride(0), ldb=ld, a_major=1, b_major=1)
    _execute_gemm(xr, dz1, None, dw_in, M=c_in, N=ld, K=p_in, lda=xr.stride(0), ldb=ld, a_major=1, b_major=1)
    if side is not None:
        main.wait_stream(side)
    g_lam, g_kh, g_kw = torch.empty_like(lam), torch.empty_like(kh), torch.empty_like(kw)
    g_fout, g_fin = torch.empty((c_out, r), dtype=lam.dtype, device=dev), torch.empty((c_in, r), dtype=lam.dtype, device=dev)
    _execute_chain_post(c_in, c_out, r, dw_out, dw_in, acc, kh.contiguous(), kw.contiguous(), lam.contiguous(), g_fout, g_fin, g_kh,
                        g_kw, g_lam)
    return [dx, g_lam, g_fout, g_fin, g_kh, g_kw]


@_chain_bwd_op.register_fake
def _(gy, xr, z1, z2, wp_in, wp_out, taps_d, lam, kh, kw, c_in, c_out, batch, height, width, stride):
    r = lam.shape[0]
    return [torch.empty_like(xr), torch.empty_like(lam), lam.new_empty((c_out, r)), lam.new_empty((c_in, r)), torch.empty_like(kh),
            torch.empty_like(kw)]


def _chain_setup(ctx, inputs, output):
    xr, lam, f_out, f_in, kh, kw, bias, ctx.batch, ctx.height, ctx.width, ctx.stride = inputs
    ctx.c_in, ctx.c_out = f_in.shape[0], f_out.shape[0]
    ctx.bias_dtype = None if bias is None else bias.dtype
    ctx.save_for_backward(xr, output[1], output[2], output[3], output[4], output[5], lam, kh, kw)
    ctx.set_materialize_grads(False)       # the auxiliary outputs never get gradients: no zero tensors for them


def _chain_backward(ctx, grads):
    if grads[0] is None:
        return (None,) * 11
    gy = grads[0].contiguous()
    xr, z1, z2, wp_in, wp_out, taps_d, lam, kh, kw = ctx.saved_tensors
    dx, g_lam, g_fout, g_fin, g_kh, g_kw = _chain_bwd_op(gy, xr, z1, z2, wp_in, wp_out, taps_d, lam, kh, kw, ctx.c_in, ctx.c_out,
                                                         ctx.batch, ctx.height, ctx.width, ctx.stride)
    g_bias = None
    if ctx.bias_dtype is not None and ctx.needs_input_grad[6]:
        g_bias = ops.reduce_to(gy, "bo", "o")[:ctx.c_out].to(ctx.bias_dtype)
    return dx, g_lam, g_fout, g_fin, g_kh, g_kw, g_bias, None, None, None, None


_chain_fwd_op.register_autograd(_chain_backward, setup_context=_chain_setup)


def cp_conv_rows_sep(x, weights, factors, bias, stride):
    """CP convolution on channels-last rows with the separable depthwise sweep: x (B, C_in, H, W) bf16; factors = [F_out (C_out, R),
    F_in (C_in, R), kh (3, R), kw (3, R)], weights = lambda (R).  x -> rows -> [C_in -> R: GEMM] -> separable depthwise (lambda in
    the W taps) -> [R -> C_out: GEMM + bias] -> NCHW, as one op: 4 launches forward, 7 backward, no Khatri-Rao product of the taps,
    no re-pitching / casting / slicing kernels around the GEMMs (csrc/cpchain.cu)."""
    b, h, w = x.shape[0], x.shape[2], x.shape[3]
    xt = nchw_to_rows(x)                                                       # (B S, round8(C_in))
    yt = _chain_fwd_op(xt, weights, factors[0], factors[1], factors[2], factors[3], bias, b, h, w, stride)[0]
    return rows_to_nchw(yt, b, factors[0].shape[0], [h // stride, w // stride], channels_last=is_channels_last(x))


def dwconv_rows(z, w_taps, batch, in_size, kernel, stride, padding, dilation):
    """z (B * prod(in_size), Cp) bf16, w_taps (prod(kernel), Cp) -> (B * prod(out_size), Cp)."""
    return _dwrows_op(z, w_taps, batch, list(in_size), list(kernel), list(stride), list(padding), list(dilation))


def cp_conv_rows(x, taps_t, f_in, f_out_scaled, bias, kernel, stride, padding, dilation):
    """CP convolution on channels-last rows, generic kernel shapes (1-D / 3-D, dilation, other paddings).  x (B, C_in, *S) bf16;
    taps_t (prod(kernel), R) the merged depthwise kernel, taps-major; f_in (C_in, R); f_out_scaled (C_out, R); lambda is folded
    into the taps or into f_out_scaled by the caller.
    x -> rows -> [C_in -> R: GEMM] -> depthwise on rows -> [R -> C_out: GEMM + bias] -> NCHW."""
    b, spatial = x.shape[0], list(x.shape[2:])
    r = f_in.shape[1]
    out_size = ops.conv_out_size(spatial, list(kernel), list(stride), list(padding), list(dilation))
    xt = nchw_to_rows(x)                                                       # (B S, round8(C_in))
    z = padded_linear(xt, f_in.t(), None, True)                                # (B S, round8(R)), pad columns zero
    w_taps = F.pad(taps_t, (0, _r8(r) - r))                                    # (taps, round8(R))
    z = dwconv_rows(z, w_taps, b, spatial, kernel, stride, padding, dilation)  # (B S', round8(R))
    yt = padded_linear(z, f_out_scaled, bias, True)                            # (B S', round8(C_out))
    return rows_to_nchw(yt, b, f_out_scaled.shape[0], out_size, channels_last=is_channels_last(x))
