"""Custom ops (with autograd) for the fused Tucker-linear path -- BASELINE config 5 (Tucker, tensor-core-bound regime).

  tltorch::mmd16           y[b] = x[b] x_0 M0 x_1 M1 x_2 M2 (+ addend) for bf16 samples whose three modes are all <= 16:
                           ONE launch of tlt_mmd16_fwd (a warp keeps a sample on chip as a 16^3 tile, HMMA mode products).
                           Backward is ONE launch of tlt_mmd16_bwd: data gradient AND all three matrix gradients.
  tltorch::padded_linear   y = x . W^T for a compact (N, K) bf16 weight whose K / N break TMA's 16-byte pitch rule (3375 =
                           15^3): activations live at a pitch of round8(K) / round8(N), the weight is re-pitched once per
                           step (tlt_pack_kmajor) and the three GEMMs (fwd, dX, dW) run on tcgen05 with the true extents
                           (TMA zero-fills / clips the ragged edge).

``tucker3_linear`` chains  mmd16 (project the in-modes) -> padded_linear (core) -> mmd16 (expand the out-modes, + bias).
Reference: linear_tucker / tensor_dot_tucker (tltorch/functional/factorized_linear.py:7-21, factorized_tensordot.py:10-66),
whose single n-ary einsum is evaluated here in the FLOP-sane order of SURVEY.md 8(a) row a3.
CPU tensors raise NotImplementedError (no fallback); the ``_execute_*`` funnels are what tests/_emulator.py swaps out.
"""
import ctypes as C
from typing import List, Optional, Tuple

import torch

from . import _cuda
from . import ops
from . import blocktt_ops as _bo

__all__ = ["mmd16_eligible", "mmd16", "padded_linear", "tucker3_linear", "tucker3_eligible"]

MIN_BATCH = 64


def _r8(v):
    return (v + 7) & ~7


def _prod(v):
    out = 1
    for s in v:
        out *= int(s)
    return out


def mmd16_eligible(dims_in, dims_out):
    """All extents in 1..16 and the zero-padding to 16^3 wastes at most 4x on either side."""
    if len(dims_in) != 3 or len(dims_out) != 3:
        return False
    if max(*dims_in, *dims_out) > 16 or min(*dims_in, *dims_out) < 1:
        return False
    return _prod(dims_in) * 4 >= 4096 and _prod(dims_out) * 4 >= 4096


# =====================================================================================================================
# launch funnels
# =====================================================================================================================
def _desc16(batch, dims_in, dims_out, flags, ld_in, ld_out):
    d = _cuda.Mmd16Desc()
    d.batch = batch
    for k in range(3):
        d.in_dims[k] = dims_in[k]
        d.out_dims[k] = dims_out[k]
        d.transpose[k] = 1 if flags[k] else 0
    d.ld_in, d.ld_out = ld_in, ld_out
    return d


def _execute_mmd16_fwd(x2, dims_in, dims_out, mats, flags, addend, y2):
    _cuda.require_cuda(x2, y2, addend, *mats)
    d = _desc16(x2.shape[0], dims_in, dims_out, flags, x2.stride(0), y2.stride(0))
    _cuda.check(_cuda.lib().tlt_mmd16_fwd(C.byref(d), _cuda.ptr(x2), _cuda.ptr(mats[0]), _cuda.ptr(mats[1]), _cuda.ptr(mats[2]),
                                          _cuda.ptr(addend), _cuda.ptr(y2), _cuda.stream_ptr(x2)), "tlt_mmd16_fwd")


def _execute_mmd16_bwd(u2, g2, dims_in, dims_out, mats, flags, du2, dM):
    """u2 (B, ld_in) forward input, g2 (B, ld_out) output gradient -> du2 (B, ld_in) and dM fp32 (3,16,16) (+=)."""
    _cuda.require_cuda(u2, g2, du2, dM, *mats)
    d = _desc16(g2.shape[0], dims_in, dims_out, flags, (u2 if u2 is not None else du2).stride(0), g2.stride(0))
    _cuda.check(_cuda.lib().tlt_mmd16_bwd(C.byref(d), _cuda.ptr(u2), _cuda.ptr(g2), _cuda.ptr(mats[0]), _cuda.ptr(mats[1]),
                                          _cuda.ptr(mats[2]), _cuda.ptr(du2), _cuda.ptr(dM), _cuda.stream_ptr(g2)),
                "tlt_mmd16_bwd")


def _execute_pack_kmajor(src2, dst):
    """dst (rows, ld) bf16 <- src2 (rows, cols) (any strides), zero-padded columns."""
    _cuda.require_cuda(src2, dst)
    _cuda.check(_cuda.lib().tlt_pack_kmajor(_cuda.ptr(src2), _cuda.dtype_code(src2.dtype), src2.shape[0], src2.shape[1],
                                            src2.stride(0), src2.stride(1), _cuda.ptr(dst), dst.stride(0),
                                            _cuda.stream_ptr(dst)), "tlt_pack_kmajor")


# =====================================================================================================================
# tltorch::mmd16
# =====================================================================================================================
@torch.library.custom_op("tltorch::mmd16", mutates_args=())
def _mmd16_op(x2: torch.Tensor, dims_in: List[int], dims_out: List[int], m0: Optional[torch.Tensor],
              m1: Optional[torch.Tensor], m2: Optional[torch.Tensor], flags: List[bool],
              addend: Optional[torch.Tensor]) -> torch.Tensor:
    y2 = torch.empty((x2.shape[0], _r8(_prod(dims_out))), dtype=x2.dtype, device=x2.device)
    if x2.shape[0]:
        mats = [None if m is None else m.contiguous() for m in (m0, m1, m2)]
        _execute_mmd16_fwd(x2, dims_in, dims_out, mats, flags, None if addend is None else addend.contiguous(), y2)
    return y2


@_mmd16_op.register_fake
def _(x2, dims_in, dims_out, m0, m1, m2, flags, addend):
    return x2.new_empty((x2.shape[0], _r8(_prod(dims_out))))


@torch.library.custom_op("tltorch::mmd16_bwd", mutates_args=())
def _mmd16_bwd_op(u2: torch.Tensor, g2: torch.Tensor, dims_in: List[int], dims_out: List[int], m0: Optional[torch.Tensor],
                  m1: Optional[torch.Tensor], m2: Optional[torch.Tensor], flags: List[bool], need_du: bool,
                  need_dm: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    du2 = torch.empty_like(u2) if need_du else u2.new_empty(0)
    dM = torch.zeros((3, 16, 16), dtype=torch.float32, device=u2.device) if need_dm else u2.new_empty(0, dtype=torch.float32)
    if g2.shape[0] and (need_du or need_dm):
        mats = [None if m is None else m.contiguous() for m in (m0, m1, m2)]
        _execute_mmd16_bwd(u2, g2, dims_in, dims_out, mats, flags, du2 if need_du else None, dM if need_dm else None)
    return du2, dM


@_mmd16_bwd_op.register_fake
def _(u2, g2, dims_in, dims_out, m0, m1, m2, flags, need_du, need_dm):
    return (torch.empty_like(u2) if need_du else u2.new_empty(0),
            u2.new_empty((3, 16, 16), dtype=torch.float32) if need_dm else u2.new_empty(0, dtype=torch.float32))


def _mmd16_setup(ctx, inputs, output):
    x2, dims_in, dims_out, m0, m1, m2, flags, addend = inputs
    ctx.save_for_backward(x2, *[m for m in (m0, m1, m2) if m is not None])
    ctx.present = [m is not None for m in (m0, m1, m2)]
    ctx.dims_in, ctx.dims_out, ctx.flags = list(dims_in), list(dims_out), list(flags)
    ctx.has_addend = addend is not None


def _mmd16_backward(ctx, gy):
    saved = list(ctx.saved_tensors)
    x2 = saved.pop(0)
    mats = [saved.pop(0) if present else None for present in ctx.present]
    need_du = ctx.needs_input_grad[0]
    need_dm = any(ctx.needs_input_grad[3 + k] for k in range(3) if mats[k] is not None)
    gy = gy.contiguous()
    du, dM = _mmd16_bwd_op(x2, gy, ctx.dims_in, ctx.dims_out, mats[0], mats[1], mats[2], ctx.flags, need_du, need_dm)
    gm = [None, None, None]
    for k in range(3):
        if mats[k] is None or not ctx.needs_input_grad[3 + k]:
            continue
        blk = dM[k, :ctx.dims_out[k], :ctx.dims_in[k]]            # [j = out index][i = in index]
        gm[k] = (blk.t() if ctx.flags[k] else blk).to(mats[k].dtype)
    gb = None
    if ctx.has_addend and ctx.needs_input_grad[7]:
        n = _prod(ctx.dims_out)
        gb = ops.reduce_to(gy[:, :n] if gy.shape[1] != n else gy, "bo", "o")
    return (du if need_du else None), None, None, gm[0], gm[1], gm[2], None, gb


_mmd16_op.register_autograd(_mmd16_backward, setup_context=_mmd16_setup)


def mmd16(x2, dims_in, dims_out, mats, flags, addend=None):
    """x2 (B, >= prod(dims_in)) at a pitch that is a multiple of 8 -> (B, round8(prod(dims_out))); pad columns are zero."""
    return _mmd16_op(x2, list(dims_in), list(dims_out), mats[0], mats[1], mats[2], list(flags), addend)


# =====================================================================================================================
# tltorch::padded_linear
# =====================================================================================================================
@torch.library.custom_op("tltorch::padded_linear", mutates_args=())
def _plinear_op(xp: torch.Tensor, w2: torch.Tensor, bias: Optional[torch.Tensor], full_pad: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """xp (B, round8(K)) bf16, w2 (N, K) bf16 (any strides) -> yp (B, round8(N)), wp the re-pitched weight.

    full_pad=False: wp is (N, round8(K)); the GEMM runs with the true extents (TMA zero-fills / clips the ragged edge) and
    the pad columns of yp are left unwritten.  full_pad=True: wp is (round8(N), round8(K)) with zero pad rows / columns
    and the GEMM runs on the padded extents, so the pad columns of yp are exact zeros (xp's pad columns must be finite)."""
    N, K = w2.shape
    B = xp.shape[0]
    Np, Kp = _r8(N), _r8(K)
    wp = torch.empty((Np if full_pad else N, Kp), dtype=torch.bfloat16, device=xp.device)
    _execute_pack_kmajor(w2, wp[:N])
    if full_pad and Np != N:
        wp[N:].zero_()
    yp = torch.empty((B, Np), dtype=torch.bfloat16, device=xp.device)
    if bias is not None and full_pad and Np != N:
        bias = torch.cat([bias, bias.new_zeros(Np - N)])
    if B:
        n, k = (Np, Kp) if full_pad else (N, K)
        _bo._execute_gemm(xp, wp, bias, yp, M=B, N=n, K=k, lda=xp.stride(0), ldb=wp.stride(0), a_major=0, b_major=0)
    return yp, wp


@_plinear_op.register_fake
def _(xp, w2, bias, full_pad):
    N, K = w2.shape
    return xp.new_empty((xp.shape[0], _r8(N))), xp.new_empty((_r8(N) if full_pad else N, _r8(K)))


@torch.library.custom_op("tltorch::padded_linear_bwd", mutates_args=())
def _plinear_bwd_op(gyp: torch.Tensor, xp: torch.Tensor, wp: torch.Tensor, N: int, K: int, full_pad: bool, need_dx: bool,
                    need_dw: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    B = xp.shape[0]
    n, k = (wp.shape[0], wp.shape[1]) if full_pad else (N, K)
    dxp = torch.empty_like(xp) if need_dx else xp.new_empty(0)
    if need_dx and B:
        # dx[b, i] = sum_o gy[b, o] W[o, i]:  B operand (n = i, k = o) at wp[o * ld + i] -> MN-major
        _bo._execute_gemm(gyp, wp, None, dxp, M=B, N=k, K=n, lda=gyp.stride(0), ldb=wp.stride(0), a_major=0, b_major=1)
    if need_dw:
        # few output tiles (skinny weights): fp32 output so that the GEMM may split K across all SM pairs (TMA reduce-add)
        few_tiles = ((N + 255) // 256) * ((K + 255) // 256) < 74
        dwp = torch.empty((N, _r8(K)), dtype=torch.float32 if few_tiles else torch.bfloat16, device=xp.device)
        if B:
            # dW[o, i] = sum_b gy[b, o] x[b, i]: both operands MN-major
            _bo._execute_gemm(gyp, xp, None, dwp, M=N, N=K, K=B, lda=gyp.stride(0), ldb=xp.stride(0), a_major=1, b_major=1)
        else:
            dwp.zero_()
    else:
        dwp = xp.new_empty(0)
    return dxp, dwp


@_plinear_bwd_op.register_fake
def _(gyp, xp, wp, N, K, full_pad, need_dx, need_dw):
    return (torch.empty_like(xp) if need_dx else xp.new_empty(0),
            xp.new_empty((N, _r8(K))) if need_dw else xp.new_empty(0))


def _plinear_setup(ctx, inputs, output):
    xp, w2, bias, full_pad = inputs
    _, wp = output
    ctx.save_for_backward(xp, wp)
    ctx.N, ctx.K = w2.shape
    ctx.w_dtype = w2.dtype
    ctx.full_pad = full_pad
    ctx.has_bias = bias is not None
    ctx.set_materialize_grads(False)


def _plinear_backward(ctx, gyp, gwp_unused):
    if gyp is None:
        return None, None, None, None
    xp, wp = ctx.saved_tensors
    need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
    gyp = gyp.contiguous()
    dxp, dwp = _plinear_bwd_op(gyp, xp, wp, ctx.N, ctx.K, ctx.full_pad, need_dx, need_dw)
    dw = dwp[:, :ctx.K].to(ctx.w_dtype) if need_dw else None
    db = None
    if ctx.has_bias and ctx.needs_input_grad[2]:
        db = ops.reduce_to(gyp, "bo", "o")[:ctx.N]
    return (dxp if need_dx else None), dw, db, None


_plinear_op.register_autograd(_plinear_backward, setup_context=_plinear_setup)


def padded_linear(xp, w2, bias=None, full_pad=False):
    """yp[:, :N] = xp[:, :K] . w2^T (+ bias) with every operand on the tcgen05 GEMM; see ``_plinear_op`` for the pitches."""
    yp, _ = _plinear_op(xp, w2, bias, full_pad)
    return yp


# =====================================================================================================================
# Tucker linear (3 in-modes, 3 out-modes)
# =====================================================================================================================
def tucker3_eligible(x2, core, factors, bias, in_shape, out_shape):
    if len(in_shape) != 3 or len(out_shape) != 3 or core.dim() != 6 or len(factors) != 6:
        return False
    if x2.dtype != torch.bfloat16 or core.dtype != torch.bfloat16 or any(f.dtype != torch.bfloat16 for f in factors):
        return False
    if bias is not None and bias.dtype != torch.bfloat16:
        return False
    if x2.shape[0] < MIN_BATCH or _prod(in_shape) % 8 or _prod(out_shape) % 8:
        return False
    r_out, r_in = list(core.shape[:3]), list(core.shape[3:])
    return mmd16_eligible(in_shape, r_in) and mmd16_eligible(r_out, out_shape)


def tucker3_linear(x2, core, factors, bias, in_shape, out_shape):
    """x2 (B, prod(in_shape)) bf16 -> (B, prod(out_shape)).  core (r_out..., r_in...); factors = out factors (o_k, r_k) then
    in factors (i_k, r_k) -- the storage order of TuckerTensorized for transpose=True."""
    r_out, r_in = list(core.shape[:3]), list(core.shape[3:])
    f_out, f_in = list(factors[:3]), list(factors[3:])
    xp = mmd16(x2.contiguous(), in_shape, r_in, f_in, [True] * 3)                     # (B, round8(prod r_in))
    yp = padded_linear(xp, core.reshape(_prod(r_out), _prod(r_in)))                   # (B, round8(prod r_out))
    return mmd16(yp, r_out, out_shape, f_out, [False] * 3, bias)                      # (B, prod(out_shape))
