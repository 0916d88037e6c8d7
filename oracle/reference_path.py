"""CPU restatement of the reference's factorized-layer hot path, on plain torch tensors.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product (torch_b200/) never imports it.

Every function names the reference file:line it follows (paths relative to /root/reference).  Functions take
the raw parameter tensors (weights / core / factors) rather than the reference's nn.Module containers so that
the same call works on fixtures loaded from tests/golden/ without the reference being importable.

``order='reference'`` executes the identical n-ary ``torch.einsum`` the reference issues (left-to-right
without opt_einsum -- infeasible at BASELINE sizes for CP/Tucker linear);
``order='sane'`` executes the same contraction with in-mode factors first (what opt_einsum would pick).
"""
import math
from typing import List, Sequence

import torch
import torch.nn.functional as F

from . import tensorly_semantics as tls

_SYMS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
_CONV = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}


# ======================================================================================
# Reconstruction  
# ======================================================================================
def cp_to_tensor(weights, factors):
    """CPTensor.to_tensor -- tltorch/factorized_tensors/factorized_tensors.py:129-130."""
    return tls.cp_to_tensor((weights, list(factors)))


def tucker_to_tensor(core, factors):
    """TuckerTensor.to_tensor -- tltorch/factorized_tensors/factorized_tensors.py:307-308."""
    return tls.tucker_to_tensor((core, list(factors)))


def tt_to_tensor(cores):
    """TTTensor.to_tensor -- tltorch/factorized_tensors/factorized_tensors.py:414-415."""
    return tls.tt_to_tensor(list(cores))


def blocktt_to_tensor(cores, tensorized_shape):
    """BlockTT.to_tensor -- tltorch/factorized_tensors/tensorized_matrices.py:354-384.

    Cores are (r_k, d0_k, d1_k, ..., r_{k+1}); tensorized (tuple) modes are multiplied out core by core, plain
    int modes are shared (batched).  Result has shape ``sum(tensorized_shape)`` flattened.
    """
    n_modes = len(tensorized_shape)
    lhs, rhs, out = "", "", ""
    for i, s in enumerate(tensorized_shape):
        a = _SYMS[3 + i]
        lhs += a
        if isinstance(s, int):
            rhs += a
            out += a
        else:
            b = _SYMS[3 + n_modes + i]
            rhs += b
            out += a + b
    eq = f"a{lhs}b,b{rhs}c->a{out}c"
    res = cores[0]
    for core in list(cores)[1:]:
        new_shape = list(res.shape)
        for i, s in enumerate(tensorized_shape):
            if not isinstance(s, int):
                new_shape[i + 1] *= core.shape[i + 1]
        new_shape[-1] = core.shape[-1]
        res = torch.einsum(eq, res, core).reshape(new_shape)
    flat = sum([(e,) if isinstance(e, int) else tuple(e) for e in tensorized_shape], ())
    return res.squeeze(0).squeeze(-1).reshape(flat)


def to_matrix(full_tensor, matrix_shape):
    """TensorizedTensor.to_matrix -- tltorch/factorized_tensors/core.py:536-543."""
    return full_tensor.reshape(tuple(int(s) for s in matrix_shape))


# ======================================================================================
# Factorized linear  
# ======================================================================================
def blocktt_linear_equation(n):
    """Equation built by linear_blocktt(transpose=True) -- tltorch/functional/factorized_linear.py:46-58."""
    x = "a" + "".join(_SYMS[1 + i] for i in range(n))
    cores = []
    for i in range(n):
        cores.append(_SYMS[1 + 2 * n + i] + _SYMS[1 + n + i] + _SYMS[1 + i] + _SYMS[1 + 2 * n + i + 1])
    out = "a" + "".join(_SYMS[1 + n + i] for i in range(n))
    return x + "," + ",".join(cores) + "->" + out


def blocktt_linear(x, cores, in_shape):
    """linear_blocktt(tensor, tt_matrix, transpose=True) -- tltorch/functional/factorized_linear.py:39-60.

    cores[k] has shape (r_k, o_k, i_k, r_{k+1}); returns (B, prod(o))."""
    n = len(in_shape)
    xt = x.reshape(-1, *in_shape)
    res = torch.einsum(blocktt_linear_equation(n), xt, *cores)
    return res.reshape(res.shape[0], -1)


def cp_linear_equation(n_out, n_in):
    """Equation built by linear_cp -> tensor_dot_cp -- tltorch/functional/factorized_tensordot.py:83-101."""
    x = "".join(_SYMS[1 + i] for i in range(n_in + 1))          # 'b' batch, then in-modes
    out_syms = "".join(_SYMS[1 + n_in + 1 + i] for i in range(n_out))
    factors = [f"{s}a" for s in out_syms] + [f"{s}a" for s in x[1:]]
    return x + ",a," + ",".join(factors) + "->" + x[0] + out_syms


def cp_linear(x, weights, factors, in_shape, order="reference"):
    """linear_cp(transpose=True) -- tltorch/functional/factorized_linear.py:23-36 + factorized_tensordot.py:70-103.

    factors = out-mode factors (o_k, R) first, then in-mode factors (i_l, R).  Returns (B, *out_shape)."""
    n_in = len(in_shape)
    n_out = len(factors) - n_in
    xt = x.reshape(-1, *in_shape)
    if order == "reference":
        return torch.einsum(cp_linear_equation(n_out, n_in), xt, weights, *factors)
    # same contraction, in-factors first
    out_f, in_f = list(factors[:n_out]), list(factors[n_out:])
    z = xt.reshape(xt.shape[0], -1) @ tls.khatri_rao(in_f)          # (B, R)
    z = z * weights
    y = z @ tls.khatri_rao(out_f).t()                               # (B, O)
    return y.reshape(xt.shape[0], *[f.shape[0] for f in out_f])


def tucker_linear_equation(n_out, n_in):
    """Equation built by linear_tucker -> tensor_dot_tucker -- tltorch/functional/factorized_tensordot.py:37-64."""
    w = n_out + n_in
    rank = _SYMS[:w]
    tsym = _SYMS[w:2 * w]
    x = _SYMS[2 * w] + tsym[n_out:]
    eq = x + "," + rank + "," + ",".join(f"{s}{r}" for s, r in zip(tsym, rank))
    return eq + "->" + _SYMS[2 * w] + tsym[:n_out]


def tucker_linear(x, core, factors, in_shape, order="reference"):
    """linear_tucker(transpose=True) -- tltorch/functional/factorized_linear.py:7-21 + factorized_tensordot.py:10-66."""
    n_in = len(in_shape)
    n_out = len(factors) - n_in
    xt = x.reshape(-1, *in_shape)
    if order == "reference":
        return torch.einsum(tucker_linear_equation(n_out, n_in), xt, core, *factors)
    out_f, in_f = list(factors[:n_out]), list(factors[n_out:])
    t = tls.multi_mode_dot(xt, in_f, modes=list(range(1, n_in + 1)), transpose=True)     # (B, r_in...)
    r_in = int(math.prod(t.shape[1:]))
    g = core.reshape(-1, r_in)                                                          # (prod r_out, prod r_in)
    t = (t.reshape(t.shape[0], r_in) @ g.t()).reshape(t.shape[0], *core.shape[:n_out])
    return tls.multi_mode_dot(t, out_f, modes=list(range(1, n_out + 1)))


def factorized_linear(x, kind, params, tensorized_shape, bias=None, implementation="factorized", order="reference"):
    """factorized_linear -- tltorch/functional/linear.py:15-50.

    kind in {'cp','tucker','blocktt','dense'}; params: cp (weights, factors) / tucker (core, factors) /
    blocktt cores / dense tensor.  tensorized_shape = (out_shape, in_shape)."""
    assert implementation in {"factorized", "reconstructed"}, \
        f"Expect implementation from [factorized, reconstructed], but got {implementation}"
    out_shape, in_shape = tuple(tensorized_shape[0]), tuple(tensorized_shape[1])
    in_features = int(math.prod(in_shape))
    if implementation == "factorized" and kind in ("cp", "tucker", "blocktt"):
        lead = x.shape[:-1]
        if kind == "cp":
            y = cp_linear(x.reshape(*lead, *in_shape), params[0], params[1], in_shape, order=order)
        elif kind == "tucker":
            y = tucker_linear(x.reshape(*lead, *in_shape), params[0], params[1], in_shape, order=order)
        else:
            y = blocktt_linear(x.reshape(*lead, *in_shape), params, in_shape)
        y = y.reshape(*lead, -1)
        return y if bias is None else y + bias
    w = reconstruct(kind, params, tensorized_shape)
    return F.linear(x, w.reshape(-1, in_features), bias)


def reconstruct(kind, params, tensorized_shape=None):
    if kind == "cp":
        return cp_to_tensor(params[0], params[1])
    if kind == "tucker":
        return tucker_to_tensor(params[0], params[1])
    if kind == "tt":
        return tt_to_tensor(params)
    if kind == "blocktt":
        return blocktt_to_tensor(params, tensorized_shape)
    if kind == "dense":
        return params
    raise ValueError(kind)


# ======================================================================================
# Factorized convolutions  
# ======================================================================================
def _tuple(v, order):
    return (v,) * order if isinstance(v, int) else tuple(v)


def convolve(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    """convolve -- tltorch/functional/convolution.py:16-53 (dense weight)."""
    try:
        fn = _CONV[weight.dim() - 2]
    except KeyError:
        raise ValueError(f"Got tensor of order={weight.dim()} but pytorch only supports up to 3rd order (3D) Convs.")
    return fn(x, weight, bias=bias, stride=stride, padding=padding, dilation=dilation, groups=groups)


def general_conv1d(x, kernel, mode, bias=None, stride=1, padding=0, groups=1, dilation=1):
    """general_conv1d -- tltorch/functional/convolution.py:99-137: 1-D conv along ``mode`` through an N-D conv
    whose kernel is size-1 (stride/dilation 1, padding 0) on every other spatial mode."""
    order = x.dim() - 2
    for i in range(2, x.dim()):
        if i != mode:
            kernel = kernel.unsqueeze(i)
    pick = lambda v, other: tuple(v if i == (mode - 2) else other for i in range(order))  # noqa: E731
    return _CONV[order](x, kernel, bias=bias, stride=pick(stride, 1), padding=pick(padding, 0),
                        dilation=pick(dilation, 1), groups=groups)


def cp_conv(x, weights, factors, bias=None, stride=1, padding=0, dilation=1):
    """cp_conv -- tltorch/functional/convolution.py:231-283.  factors: [F0 (C_out,R), F1 (C_in,R), F2.. (k_j,R)]."""
    rank = factors[0].shape[1]
    order = len(factors) - 2
    stride, padding, dilation = _tuple(stride, order), _tuple(padding, order), _tuple(dilation, order)
    b = x.shape[0]
    shp = list(x.shape)
    x = x.reshape(b, shp[1], -1).contiguous()
    x = F.conv1d(x, factors[1].t().unsqueeze(2))                                   # :262
    shp[1] = rank
    x = x.reshape(shp)
    for i in range(order):                                                         # :266-271
        kernel = factors[i + 2].t().unsqueeze(1)
        x = general_conv1d(x.contiguous(), kernel, i + 2, stride=stride[i], padding=padding[i],
                           dilation=dilation[i], groups=rank)
    shp = list(x.shape)
    x = x.reshape(b, shp[1], -1)
    x = F.conv1d(x * weights.unsqueeze(1).unsqueeze(0), factors[0].unsqueeze(2), bias=bias)   # :278
    shp[1] = x.shape[1]
    return x.reshape(shp)


def cp_conv_mobilenet(x, weights, factors, bias=None, stride=1, padding=0, dilation=1):
    """cp_conv_mobilenet -- tltorch/functional/convolution.py:286-345."""
    rank = factors[0].shape[1]
    order = len(factors) - 2
    b = x.shape[0]
    shp = list(x.shape)
    x = x.reshape(b, shp[1], -1).contiguous()
    x = F.conv1d(x, factors[1].t().unsqueeze(2))                                   # :311
    shp[1] = rank
    x = x.reshape(shp)
    if order == 1:
        w = factors[2].t().unsqueeze(1)
    elif order == 2:
        w = tls.batched_outer(factors[2].t(), factors[3].t()).unsqueeze(1)           # :323-326
    elif order == 3:
        w = tls.batched_outer(factors[2].t(), tls.batched_outer(factors[3].t(), factors[4].t())).unsqueeze(1)
    else:
        raise ValueError("order > 3")
    x = _CONV[order](x.contiguous(), w, stride=stride, padding=padding, dilation=dilation, groups=rank)
    shp = list(x.shape)
    x = x.reshape(b, shp[1], -1)
    x = F.conv1d(x * weights.unsqueeze(1).unsqueeze(0), factors[0].unsqueeze(2), bias=bias)   # :340
    shp[1] = x.shape[1]
    return x.reshape(shp)


def tucker_conv(x, core, factors, bias=None, stride=1, padding=0, dilation=1):
    """tucker_conv -- tltorch/functional/convolution.py:140-173.  factors: [F0 (C_out,r0), F1 (C_in,r1), F2.. (k_j,r_j)]."""
    b = x.shape[0]
    n_dim = x.dim()
    shp = list(x.shape)
    x = x.reshape(b, shp[1], -1).contiguous()
    x = F.conv1d(x, factors[1].t().unsqueeze(2))                                   # :154
    shp[1] = core.shape[1]
    x = x.reshape(shp)
    modes = list(range(2, n_dim))
    w = tls.multi_mode_dot(core, list(factors[2:]), modes=modes)                   # :160
    x = convolve(x, w, bias=None, stride=stride, padding=padding, dilation=dilation)
    shp = list(x.shape)
    x = x.reshape(b, shp[1], -1)
    x = F.conv1d(x, factors[0].unsqueeze(2), bias=bias)                            # :168
    shp[1] = x.shape[1]
    return x.reshape(shp)


def tt_conv(x, cores, bias=None, stride=1, padding=0, dilation=1):
    """tt_conv -- tltorch/functional/convolution.py:176-228.  cores follow the factorization shape
    (C_in, k_1..k_d, C_out): cores[0] (1,C_in,r1), cores[j+1] (r,k_j,r'), cores[-1] (r,C_out,1)."""
    order = len(cores) - 2
    stride, padding, dilation = _tuple(stride, order), _tuple(padding, order), _tuple(dilation, order)
    b = x.shape[0]
    shp = list(x.shape)
    x = x.reshape(b, shp[1], -1).contiguous()
    x = F.conv1d(x, cores[0].permute(2, 1, 0))                                     # :207
    shp[1] = x.shape[1]
    x = x.reshape(shp)
    for i in range(order):                                                         # :211-216
        kernel = cores[i + 1].permute(2, 0, 1)
        x = general_conv1d(x.contiguous(), kernel, i + 2, stride=stride[i], padding=padding[i],
                           dilation=dilation[i])
    shp = list(x.shape)
    x = x.reshape(b, shp[1], -1)
    x = F.conv1d(x, cores[-1].permute(1, 0, 2), bias=bias)                         # :223
    shp[1] = x.shape[1]
    return x.reshape(shp)


def reconstructed_conv(x, kind, params, bias=None, stride=1, padding=0, dilation=1):
    """convolve() with a factorized weight -- convolution.py:45-50 (TT: moveaxis(-1, 0))."""
    w = reconstruct(kind, params)
    if kind == "tt":
        w = torch.movedim(w, -1, 0)
    return convolve(x, w, bias=bias, stride=stride, padding=padding, dilation=dilation)


# ======================================================================================
# Tensor regression / contraction layers  
# ======================================================================================
def tucker_trl(x, core, factors, project_input=False, bias=None):
    """tucker_trl -- tltorch/functional/tensor_regression.py:37-49."""
    n_input = x.dim() - 1
    if project_input:
        x = tls.multi_mode_dot(x, list(factors[:n_input]), modes=range(1, n_input + 1), transpose=True)
        w = tls.multi_mode_dot(core, list(factors[n_input:]), modes=range(n_input, core.dim()))
    else:
        w = tucker_to_tensor(core, factors)
    y = tls.inner(x, w, n_modes=x.dim() - 1)
    return y if bias is None else y + bias


def trl(x, kind, params, bias=None, project_input=False):
    """trl -- tltorch/functional/tensor_regression.py:11-34."""
    if kind == "tucker":
        return tucker_trl(x, params[0], params[1], project_input=project_input, bias=bias)
    y = tls.inner(x, reconstruct(kind, params), n_modes=x.dim() - 1)
    return y if bias is None else y + bias


def tcl(x, factors, bias=None):
    """TCL.forward -- tltorch/factorized_layers/tensor_contraction_layers.py:73-81; factors[k] is (rank_k, input_k)."""
    y = tls.multi_mode_dot(x, list(factors), modes=list(range(1, len(factors) + 1)))
    return y if bias is None else y + bias


# ======================================================================================
# Tensor dropout (mask APPLICATION).  Index sampling is separated from application so that the
# product and the oracle can be fed identical keep-index vectors.
# ======================================================================================
def sample_keep_indices(rank, proba, min_dim, min_values=1, device="cpu"):
    """One rank mode of the sampling loop -- tltorch/tensor_hooks/_tensor_dropout.py:64-73 (Tucker),
    :93-98 (CP), :116-127 (TT).  Same torch RNG calls in the same order."""
    idx = torch.arange(rank, device=device, dtype=torch.int64)
    if rank > min_dim:
        keep = torch.bernoulli(torch.ones(rank, device=device) * (1 - proba),
                               out=torch.empty(rank, device=device, dtype=torch.bool))
        idx = idx[keep]
        if len(idx) == 0:
            idx = torch.randint(0, rank, size=(min_values,), device=device, dtype=torch.int64)
    return idx


def tucker_dropout_apply(core, factors, indices, proba, training=True):
    """TuckerDropout._apply_tensor_dropout -- _tensor_dropout.py:75-82 (note the 1/(1-p)^core.ndim scale)."""
    sub = core[torch.meshgrid(*indices, indexing="ij")]
    if training:
        sub = sub * (1 / ((1 - proba) ** core.dim()))
    return sub, [f[:, idx] for f, idx in zip(factors, indices)]


def cp_dropout_apply(weights, factors, indices, proba, training=True):
    """CPDropout._apply_tensor_dropout -- _tensor_dropout.py:100-106."""
    w = weights[indices]
    if training:
        w = w * (1 / (1 - proba))
    return w, [f[:, indices] for f in factors]


def tt_dropout_apply(cores, indices, proba, training=True):
    """TTDropout._apply_tensor_dropout -- _tensor_dropout.py:129-142; indices[i] subsamples rank[i+1]."""
    scaling = 1 / (1 - proba) if training else 1
    out = []
    n = len(cores)
    for i, f in enumerate(cores):
        if i == 0:
            out.append(f[..., indices[i]] * scaling)
        elif i == n - 1:
            out.append(f[indices[i - 1], ...])
        else:
            out.append(f[indices[i - 1], ...][..., indices[i]] * scaling)
    return out


# ======================================================================================
# fwd + bwd helper used by parity tests and the CPU baseline
# ======================================================================================
def fwd_bwd(fn, inputs: Sequence[torch.Tensor], grad_out=None):
    """Run fn(*leaves) and backprop grad_out; returns (y, [grad per input])."""
    leaves = [t.detach().clone().requires_grad_(True) for t in inputs]
    y = fn(*leaves)
    if grad_out is None:
        grad_out = torch.ones_like(y)
    grads: List[torch.Tensor] = list(torch.autograd.grad(y, leaves, grad_out, allow_unused=True))
    return y.detach(), grads
