"""Custom op (with autograd) for a fused chain of mode products on the trailing modes of a contiguous tensor.

  tltorch::multi_mode_dot(x (batch, d0, d1, d2), m0 | None, m1 | None, m2 | None, transpose flags) -> (batch, r0, r1, r2)

One launch of tlt_multi_mode_dot keeps each sample on chip (x[b] staged once, the up-to-three small contractions
ping-pong through shared memory, only y[b] is written).  Reference call sites: tenalg.multi_mode_dot in TCL.forward
(tltorch/factorized_layers/tensor_contraction_layers.py:75), tucker_trl's input projection
(tltorch/functional/tensor_regression.py:40) and the project / expand stages of linear_tucker
(tltorch/functional/factorized_tensordot.py:10-66).  Backward: the data gradient is the same op with the transposed
matrices; each matrix gradient is the fused op with that mode left untouched followed by one long-reduction
contraction.  ``plan`` decides eligibility; everything else stays on the per-mode ``ops.contract`` chain.
"""
import ctypes as C
from typing import List, Optional

import torch

from . import _cuda
from . import ops

__all__ = ["plan", "fused_multi_mode_dot"]

_SMEM_LIMIT = 200 * 1024
MIN_BATCH = 64
MIN_SAMPLE_ELEMS = 512


def _r4(v):
    return (v + 3) & ~3


def _fits(d, r, skip):
    p2 = d[2] if skip[2] else (d[2] | 1)
    off = sum(_r4(r[k] * _r4(d[k])) for k in range(3) if not skip[k])
    e0 = d[0] if skip[0] else r[0]
    e1 = d[1] if skip[1] else r[1]
    buf = _r4(max(d[0] * d[1] * p2, e0 * d[1] * p2, e0 * e1 * p2))
    return (off + 2 * buf) * 4 <= _SMEM_LIMIT


def plan(tensor, matrices, modes, transpose):
    """Can ``multi_mode_dot(tensor, matrices, modes)`` run as ONE fused launch?  Returns (lead, d, r, mats, flags) or None.

    Conditions: every entry is a 2-D matrix, the touched modes lie among the last three dims, the tensor is contiguous
    floating point, the leading dims give enough samples to fill the machine, and one sample fits shared memory."""
    nd = tensor.dim()
    if nd < 2 or not tensor.is_contiguous() or tensor.dtype not in (torch.float32, torch.bfloat16, torch.float64):
        return None
    modes = [m + nd if m < 0 else m for m in modes]
    if len(set(modes)) != len(modes) or not matrices:
        return None
    first = min(modes)
    if first < max(nd - 3, 1):            # mode 0 is never fused: the leading dims are the samples
        return None
    n_tail = nd - first
    shape = list(tensor.shape)
    d = [1] * (3 - n_tail) + shape[first:]
    r = list(d)
    mats: List[Optional[torch.Tensor]] = [None, None, None]
    for m, mode in zip(matrices, modes):
        if m.dim() != 2 or m.dtype != tensor.dtype:
            return None
        k = 3 - n_tail + (mode - first)
        rows, cols = (m.shape[1], m.shape[0]) if transpose else (m.shape[0], m.shape[1])
        if cols != shape[mode]:
            return None
        mats[k] = m
        r[k] = rows
    skip = [m is None for m in mats]
    batch = 1
    for s in shape[:first]:
        batch *= s
    per_sample = d[0] * d[1] * d[2]
    if batch < MIN_BATCH or per_sample < MIN_SAMPLE_ELEMS or max(d + r) > 4096 or not _fits(d, r, skip):
        return None             # tiny samples (core tensors, 3x3 taps) cannot fill a 256-thread block: per-mode chain
    return shape[:first], d, r, mats, [bool(transpose)] * 3


# =====================================================================================================================
# launch funnel
# =====================================================================================================================
def _execute_mmd(x4, mats, flags, addend, y4):
    _cuda.require_cuda(x4, y4, addend, *mats)
    d = _cuda.MmdDesc()
    d.dtype = _cuda.dtype_code(x4.dtype)
    d.batch = x4.shape[0]
    for k in range(3):
        d.in_dims[k] = x4.shape[1 + k]
        d.out_dims[k] = y4.shape[1 + k]
        d.transpose[k] = 1 if flags[k] else 0
    _cuda.check(_cuda.lib().tlt_multi_mode_dot(C.byref(d), _cuda.ptr(x4), _cuda.ptr(mats[0]), _cuda.ptr(mats[1]),
                                               _cuda.ptr(mats[2]), _cuda.ptr(addend), _cuda.ptr(y4), _cuda.stream_ptr(x4)),
                "tlt_multi_mode_dot")


def _out_dims(x4, mats, flags):
    dims = []
    for k in range(3):
        if mats[k] is None:
            dims.append(x4.shape[1 + k])
        else:
            dims.append(mats[k].shape[1] if flags[k] else mats[k].shape[0])
    return dims


@torch.library.custom_op("tltorch::multi_mode_dot", mutates_args=())
def _mmd_op(x4: torch.Tensor, m0: Optional[torch.Tensor], m1: Optional[torch.Tensor], m2: Optional[torch.Tensor],
            flags: List[bool], addend: Optional[torch.Tensor]) -> torch.Tensor:
    mats = [None if m is None else m.contiguous() for m in (m0, m1, m2)]
    x4 = x4.contiguous()
    y4 = torch.empty((x4.shape[0], *_out_dims(x4, mats, flags)), dtype=x4.dtype, device=x4.device)
    if y4.numel():
        _execute_mmd(x4, mats, flags, None if addend is None else addend.contiguous(), y4)
    return y4


@_mmd_op.register_fake
def _(x4, m0, m1, m2, flags, addend):
    return x4.new_empty((x4.shape[0], *_out_dims(x4, [m0, m1, m2], flags)))


def _mmd_setup(ctx, inputs, output):
    x4, m0, m1, m2, flags, addend = inputs
    ctx.has_addend = addend is not None
    ctx.save_for_backward(x4, *[m for m in (m0, m1, m2) if m is not None])
    ctx.present = [m is not None for m in (m0, m1, m2)]
    ctx.flags = list(flags)


def _mmd_backward(ctx, gy):
    saved = list(ctx.saved_tensors)
    x4 = saved.pop(0)
    mats = [saved.pop(0) if present else None for present in ctx.present]
    flags = ctx.flags
    gx = None
    if ctx.needs_input_grad[0]:
        gx = _mmd_op(gy, mats[0], mats[1], mats[2], [not f for f in flags], None)
    gm = [None, None, None]
    letters = "pqr"
    for k in range(3):
        if mats[k] is None or not ctx.needs_input_grad[1 + k]:
            continue
        others = [m if j != k else None for j, m in enumerate(mats)]
        z = _mmd_op(x4, others[0], others[1], others[2], flags, None) if any(m is not None for m in others) else x4
        g_spec = "b" + letters.replace(letters[k], "j")
        z_spec = "b" + letters.replace(letters[k], "i")
        out = "ij" if flags[k] else "ji"           # stored (d, r) when transposed, (r, d) otherwise
        gm[k] = ops.contract(f"{g_spec},{z_spec}->{out}", gy, z)
    gb = ops.reduce_to(gy, "bpqr", "pqr") if (ctx.has_addend and ctx.needs_input_grad[5]) else None
    return gx, gm[0], gm[1], gm[2], None, gb


_mmd_op.register_autograd(_mmd_backward, setup_context=_mmd_setup)


def fused_multi_mode_dot(tensor, lead, d, r, mats, flags, addend=None):
    """Apply the plan returned by ``plan``; ``addend`` (shaped like the trailing output modes, e.g. a bias) is fused in."""
    x4 = tensor.reshape(-1, *d)
    if addend is not None:
        addend = addend.reshape(r)
    y4 = _mmd_op(x4, mats[0], mats[1], mats[2], flags, addend)
    n_tail = tensor.dim() - len(lead)
    return y4.reshape(*lead, *r[3 - n_tail:])
