"""Factorized N-D convolutions: B200 counterparts of tltorch/functional/convolution.py.

Every 1x1 channel-mixing stage is an ``ops.contract`` launch on the NCHW tensor where it lies (weights as the row
operand so stores are coalesced along the spatial dim), depthwise stages run in the dedicated depthwise kernels, and
dense k x k stages (Tucker core, TT cores, reconstructed kernels) run as unfold + contraction.
"""
import math
import os

import torch

from .. import ops, tenalg, pointwise_ops, rows_ops, kr_ops, cprows_ops
from ..ops import contract


def _tuple(v, order):
    return (v,) * order if isinstance(v, int) else tuple(int(e) for e in v)


def _pointwise(x, w_spec, w, bias=None):
    """1x1 conv: x (B, C, *S), w indexed by ``w_spec`` in {'oc', 'co'} -> (B, O, *S)  (reference: F.conv1d k=1 call sites,
    convolution.py:154,168,207,223,262,278,311,340)."""
    b, c = x.shape[0], x.shape[1]
    x3 = x.reshape(b, c, -1)
    w_nk = w if w_spec == "oc" else w.t()
    if pointwise_ops.eligible(x3, w_nk, bias):
        # tcgen05 path: the NCHW tile is the MN-major MMA operand (3-D TMA), output transposed back in the epilogue
        y = pointwise_ops.pointwise_tc(x3, w_nk, bias)
    elif bias is not None:
        y = contract(f"{w_spec},bcs->bos", w, x3, bias, "o")
    else:
        y = contract(f"{w_spec},bcs->bos", w, x3)
    return y.reshape(b, y.shape[1], *x.shape[2:])


def _dense_conv(x, w_kmajor, w_spec, kernel, bias, stride, padding, dilation):
    """Dense conv through unfold: cols[(b,s),(c,taps)] then one contraction writing NCHW directly.
    ``w_kmajor`` is indexed by ``w_spec`` whose letters are 'q' (out channel), 'c' (in channel) and 't' (flattened taps)."""
    order = len(kernel)
    b, c = x.shape[0], x.shape[1]
    out_size = ops.conv_out_size(list(x.shape[2:]), kernel, stride, padding, dilation)
    cols = ops.im2col(x, kernel, stride, padding, dilation).reshape(b, math.prod(out_size), c, math.prod(kernel))
    if bias is not None:
        y = contract(f"{w_spec},bsct->bqs", w_kmajor, cols, bias, "q")
    else:
        y = contract(f"{w_spec},bsct->bqs", w_kmajor, cols)
    return y.reshape(b, y.shape[1], *out_size)


def convolve(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    """Convolution of any order 1..3 with a dense or factorized (reconstructed on the fly) kernel.
    reference: functional/convolution.py:16-53."""
    from ..factorized_tensors import TTTensor
    if not torch.is_tensor(weight):
        if isinstance(weight, TTTensor):
            weight = torch.movedim(weight.to_tensor(), -1, 0)
        else:
            weight = weight.to_tensor()
    order = weight.dim() - 2
    if order not in (1, 2, 3):
        raise ValueError(f"Got tensor of order={weight.dim()} but pytorch only supports up to 3rd order (3D) Convs.")
    stride, padding, dilation = _tuple(stride, order), _tuple(padding, order), _tuple(dilation, order)
    kernel = list(weight.shape[2:])
    if groups == 1:
        w = weight.reshape(weight.shape[0], weight.shape[1], -1)
        return _dense_conv(x, w, "qct", kernel, bias, stride, padding, dilation)
    if groups == x.shape[1] and weight.shape[0] == groups and weight.shape[1] == 1:
        y = ops.dwconv(x, weight.reshape(groups, *kernel), stride, padding, dilation)
        if bias is not None:
            one = torch.ones((), dtype=y.dtype, device=y.device)
            y3 = y.reshape(y.shape[0], y.shape[1], -1)
            y = contract("bcs,->bcs", y3, one, bias, "c").reshape(y.shape)
        return y
    raise NotImplementedError("grouped convolutions other than depthwise are not on the factorized-layer hot path")


def general_conv1d(x, kernel, mode, bias=None, stride=1, padding=0, groups=1, dilation=1, verbose=False):
    """1-D convolution along dimension ``mode`` of x (B, C, S1..SN); kernel (out, in/groups, k).
    reference: functional/convolution.py:99-137."""
    if verbose:
        print(f"Convolving {tuple(x.shape)} with {tuple(kernel.shape)} along mode {mode}, "
              f"stride={stride}, padding={padding}, groups={groups}")
    order = x.dim() - 2
    axis = mode - 2
    pick = lambda v, other: [v if i == axis else other for i in range(order)]   # noqa: E731
    kshape = pick(kernel.shape[-1], 1)
    w = kernel.reshape(kernel.shape[0], kernel.shape[1], *kshape)
    return convolve(x, w, bias=bias, stride=pick(stride, 1), padding=pick(padding, 0), dilation=pick(dilation, 1),
                    groups=groups)


def _cp_rows_route(x, cp_tensor, bias, kernel, stride):
    """Rows route for the CP convolution: bf16, enough channels for the GEMMs, and a map the NCHW tensor-core 1x1 kernel
    cannot take on BOTH sides of the depthwise stage (spatial extents not multiples of 8)."""
    factors = list(cp_tensor.factors)
    order = len(factors) - 2
    if not rows_ops.rows_eligible(x, [cp_tensor.weights, bias] + factors, order):
        return False
    if math.prod(kernel) > rows_ops.DW_ROWS_MAX_TAPS or min(x.shape[1], factors[0].shape[0], factors[0].shape[1]) < 64:
        return False
    mode = os.environ.get("TLT_CP_ROWS", "auto")          # development switch: all | off | auto
    if mode in ("all", "off"):
        return mode == "all"
    if rows_ops.is_channels_last(x):
        return True        # the activation already IS the rows matrix: no transposition on either side of the chain
    # strided depthwise passes and maps the NCHW tensor-core 1x1 kernel cannot take (spatial extent not a multiple of 8
    # on either side) go through rows; measured per layer in gpurun_out/s44 (config 3: 64->128 stride 2 3.45 -> 0.98 ms)
    s_in = math.prod(x.shape[2:])
    s_out = math.prod(-(-e // st) for e, st in zip(x.shape[2:], stride))
    return s_in % 8 != 0 or s_out % 8 != 0 or any(st > 1 for st in stride)


def _cp_conv_impl(x, cp_tensor, bias, stride, padding, dilation):
    factors = list(cp_tensor.factors)
    order = len(factors) - 2
    stride, padding, dilation = _tuple(stride, order), _tuple(padding, order), _tuple(dilation, order)
    # depthwise over all spatial modes at once: taps[(k1..kd), r] = lambda_r prod_j F_{2+j}[k_j, r] -- the weighted
    # Khatri-Rao product of the kernel-mode factors, ONE element-wise launch (kr_ops); lambda rides in the taps, so the
    # output factor is used as it is (no 'or,r->or' scaling pass and no gradient kernel for it)
    kernel = [f.shape[0] for f in factors[2:]]
    if cprows_ops.eligible(x, cp_tensor.weights, factors, bias, kernel, stride, padding, dilation):
        # HBM-bound 3x3 layers on channels-last maps (config 3's 56^2 layers): the whole chain -- and, going back, every gradient --
        # in one launch per direction, the R-channel intermediates in TMEM / shared memory (cprows_ops, csrc/cpconv_rows.cu)
        return cprows_ops.cp_conv_rows_fused(x, cp_tensor.weights, factors, bias, stride[0])
    if order == 2 and rows_ops.dwsep_eligible(factors, kernel, stride, padding, dilation, list(x.shape[2:])) and \
            _cp_rows_route(x, cp_tensor, bias, kernel, stride):
        # channels-last rows with the separable depthwise sweep: the taps go in unmerged (rows_ops.cp_conv_rows_sep)
        return rows_ops.cp_conv_rows_sep(x, cp_tensor.weights, factors, bias, stride[0])
    if kr_ops.eligible(factors[2:], cp_tensor.weights):
        taps_t = kr_ops.khatri_rao(factors[2:], cp_tensor.weights)                 # (prod k, R), taps-major
        f0 = factors[0]
    else:
        taps = factors[2].t()
        for f in factors[3:]:
            taps = tenalg.batched_outer(taps, f.t())
        taps_t = taps.reshape(taps.shape[0], -1).t()
        f0 = contract("or,r->or", factors[0], cp_tensor.weights)                   # lambda folded into the output factor
    if _cp_rows_route(x, cp_tensor, bias, kernel, stride):
        # small odd-sized or strided maps: channels-last rows -- both 1x1 stages are K-major tcgen05 GEMMs and the
        # depthwise stage moves 16-byte channel chunks (rows_ops.cp_conv_rows)
        return rows_ops.cp_conv_rows(x, taps_t, factors[1], f0, bias, kernel, stride, padding, dilation)
    # C_in -> R : z[b,r,s] = sum_c F1[c,r] x[b,c,s]
    z = _pointwise(x, "co", factors[1])
    z = ops.dwconv(z, taps_t.t().reshape(taps_t.shape[1], *kernel), stride, padding, dilation)
    return _pointwise(z, "oc", f0, bias)


def cp_conv(x, cp_tensor, bias=None, stride=1, padding=0, dilation=1):
    """Factorized CP convolution (reference: functional/convolution.py:231-283).

    The reference chains d depthwise 1-D convs; the product of those separable passes equals one depthwise N-D conv
    with the rank-1 kernel prod_j F_{2+j}[k_j, r], which is what is launched (one pass over the R-channel tensor
    instead of d)."""
    return _cp_conv_impl(x, cp_tensor, bias, stride, padding, dilation)


def cp_conv_mobilenet(x, cp_tensor, bias=None, stride=1, padding=0, dilation=1):
    """MobileNet-style CP convolution (reference: functional/convolution.py:286-345)."""
    return _cp_conv_impl(x, cp_tensor, bias, stride, padding, dilation)


def tucker_conv(x, tucker_tensor, bias=None, stride=1, padding=0, dilation=1):
    """Factorized Tucker convolution (reference: functional/convolution.py:140-173)."""
    core, factors = tucker_tensor.core, list(tucker_tensor.factors)
    order = core.dim() - 2
    stride, padding, dilation = _tuple(stride, order), _tuple(padding, order), _tuple(dilation, order)
    w = rows_ops.expand_core(core, factors[2:])                                    # (r0, r1, k...) in one launch
    if w is None:
        w = tenalg.multi_mode_dot(core, factors[2:], modes=list(range(2, core.dim())))  # (r0, r1, k...)
    if rows_ops.rows_eligible(x, [core, bias] + factors, order) and min(x.shape[1], factors[0].shape[0], *w.shape[:2]) >= rows_ops.MIN_CHANNELS:
        # channels-last rows: 1x1 -> unfold -> core GEMM over (tap, channel) -> 1x1 (+ bias), all on the tcgen05 GEMM
        return rows_ops.tucker_conv_rows(x, w, factors[1], factors[0], bias, stride, padding, dilation)
    z = _pointwise(x, "co", factors[1])                                           # C_in -> r1
    kernel = list(w.shape[2:])
    z = _dense_conv(z, w.reshape(w.shape[0], w.shape[1], -1), "qct", kernel, None, stride, padding, dilation)
    return _pointwise(z, "oc", factors[0], bias)                                   # r0 -> C_out (+ bias)


def tt_conv(x, tt_tensor, bias=None, stride=1, padding=0, dilation=1):
    """Factorized TT convolution (reference: functional/convolution.py:176-228); cores follow (C_in, k_1..k_d, C_out)."""
    cores = list(tt_tensor.factors)
    order = len(cores) - 2
    stride, padding, dilation = _tuple(stride, order), _tuple(padding, order), _tuple(dilation, order)
    z = _pointwise(x, "co", cores[0][0])                                           # (C_in, r1)
    for j in range(order):
        core = cores[j + 1]                                                        # (r_in, k_j, r_out)
        pick = lambda v, other: [v if i == j else other for i in range(order)]    # noqa: E731
        z = _dense_conv(z, core, "ctq", pick(core.shape[1], 1), None, pick(stride[j], 1), pick(padding[j], 0),
                        pick(dilation[j], 1))
    return _pointwise(z, "co", cores[-1][:, :, 0], bias)                           # (r, C_out)


def _get_factorized_conv(factorization, implementation="factorized"):
    from ..factorized_tensors import CPTensor, TTTensor, TuckerTensor
    if implementation == "reconstructed" or factorization == "Dense":
        return convolve
    if isinstance(factorization, CPTensor):
        if implementation == "factorized":
            return cp_conv
        if implementation == "mobilenet":
            return cp_conv_mobilenet
    elif isinstance(factorization, TuckerTensor):
        return tucker_conv
    elif isinstance(factorization, TTTensor):
        return tt_conv
    raise ValueError(f"Got unknown type {factorization}")


def convNd(x, weight, bias=None, stride=1, padding=0, dilation=1, implementation="factorized"):
    """reference: functional/convolution.py:363-381."""
    from ..factorized_tensors import CPTensor, TTTensor, TuckerTensor, DenseTensor
    if implementation == "reconstructed":
        weight = weight.to_tensor()
    if isinstance(weight, DenseTensor):
        return convolve(x, weight.tensor, bias=bias, stride=stride, padding=padding, dilation=dilation)
    if torch.is_tensor(weight):
        return convolve(x, weight, bias=bias, stride=stride, padding=padding, dilation=dilation)
    if isinstance(weight, CPTensor):
        if implementation == "factorized":
            return cp_conv(x, weight, bias=bias, stride=stride, padding=padding, dilation=dilation)
        if implementation == "mobilenet":
            return cp_conv_mobilenet(x, weight, bias=bias, stride=stride, padding=padding, dilation=dilation)
    elif isinstance(weight, TuckerTensor):
        return tucker_conv(x, weight, bias=bias, stride=stride, padding=padding, dilation=dilation)
    elif isinstance(weight, TTTensor):
        return tt_conv(x, weight, bias=bias, stride=stride, padding=padding, dilation=dilation)
