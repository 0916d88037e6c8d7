"""Restatement of the ``tensorly`` (pytorch backend) functions used by the reference hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``tensorly`` is a third-party dependency of the reference that is absent from /root/reference
(requirements.txt:5 lists it un-pinned; CI clones git HEAD, .github/workflows/test.yml:26-35) and
cannot be installed offline.  Each function below restates the *published* semantics of the tensorly
>= 0.8 function of the same name; the reference call sites that rely on it are cited.  Every
function has a unique closed-form definition, so an fp64 evaluation is exact up to summation order.
"""
import math
import warnings

import numpy as np
import torch


# --------------------------------------------------------------------------------------
# backend helpers (tensorly/backend/pytorch_backend.py)
# --------------------------------------------------------------------------------------
def shape(t):
    return tuple(t.shape)


def ndim(t):
    return t.dim()


def transpose(t, axes=None):
    """tl.transpose: permute; axes=None reverses all dims (convolution.py:154,207,215,223,262,270)."""
    if axes is None:
        axes = list(range(t.dim()))[::-1]
    return t.permute(*axes)


def moveaxis(t, source, destination):
    return torch.movedim(t, source, destination)


# --------------------------------------------------------------------------------------
# tenalg (tensorly/tenalg/core_tenalg)
# --------------------------------------------------------------------------------------
def unfold(t, mode):
    return torch.movedim(t, mode, 0).reshape(t.shape[mode], -1)


def fold(unfolded, mode, full_shape):
    full_shape = list(full_shape)
    mode_dim = full_shape.pop(mode)
    full_shape.insert(0, mode_dim)
    return torch.movedim(unfolded.reshape(full_shape), 0, mode)


def mode_dot(tensor, matrix_or_vector, mode, transpose=False):
    """n-mode product (reference call sites: factorized_tensors.py:327,432,448,454,468).

    matrix (J, I_mode): out = fold(M @ unfold(T, mode)); vector (I_mode,): contracts and drops the mode.
    transpose=True uses the (conjugate) transpose of the matrix.
    """
    if mode < 0:
        mode += tensor.dim()
    new_shape = list(tensor.shape)
    if matrix_or_vector.dim() == 2:
        m = matrix_or_vector.conj().t() if transpose else matrix_or_vector
        if m.shape[1] != tensor.shape[mode]:
            raise ValueError(
                f"shapes {tuple(tensor.shape)} and {tuple(m.shape)} not aligned in mode-{mode} multiplication")
        new_shape[mode] = m.shape[0]
        return fold(m @ unfold(tensor, mode), mode, new_shape)
    elif matrix_or_vector.dim() == 1:
        if matrix_or_vector.shape[0] != tensor.shape[mode]:
            raise ValueError("vector size does not match the contracted mode")
        res = matrix_or_vector @ unfold(tensor, mode)
        new_shape.pop(mode)
        return res.reshape(new_shape) if new_shape else res.reshape(())
    raise ValueError("Can only take n_mode_product with a vector or a matrix.")


def multi_mode_dot(tensor, matrix_or_vec_list, modes=None, skip=None, transpose=False):
    """Sequential mode_dot (tensor_regression.py:40-42, convolution.py:160, tensor_contraction_layers.py:75).

    Each vector contraction removes a mode, so later mode indices are shifted down by one.
    """
    if modes is None:
        modes = range(len(matrix_or_vec_list))
    decrement = 0
    res = tensor
    for i, (m, mode) in enumerate(zip(matrix_or_vec_list, modes)):
        if skip is not None and i == skip:
            continue
        res = mode_dot(res, m, mode - decrement, transpose=transpose)
        if m.dim() == 1:
            decrement += 1
    return res


def inner(a, b, n_modes=None):
    """Generalised inner product (tensor_regression.py:32,34,47,49)."""
    if n_modes is None:
        if a.shape != b.shape:
            raise ValueError("Taking a generalised product between two tensors without specifying "
                             "common modes is equivalent to taking inner product. Shapes must match.")
        return torch.sum(a * b)
    if tuple(a.shape[-n_modes:]) != tuple(b.shape[:n_modes]):
        raise ValueError(f"Incorrect shapes for inner product along {n_modes} common modes: "
                         f"{tuple(a.shape)}, {tuple(b.shape)}")
    common = int(np.prod(a.shape[-n_modes:]))
    out_shape = tuple(a.shape[:-n_modes]) + tuple(b.shape[n_modes:])
    return (a.reshape(-1, common) @ b.reshape(common, -1)).reshape(out_shape)


def batched_outer(a, b):
    """tenalg.tensordot(a, b, modes=(), batched_modes=0) (convolution.py:323-331):
    out[r, *a_rest, *b_rest] = a[r, *a_rest] * b[r, *b_rest]."""
    r = a.shape[0]
    a_rest, b_rest = a.shape[1:], b.shape[1:]
    av = a.reshape(r, *a_rest, *([1] * len(b_rest)))
    bv = b.reshape(r, *([1] * len(a_rest)), *b_rest)
    return av * bv


def tensordot(a, b, modes, batched_modes=()):
    if (modes == () or modes == []) and batched_modes in (0, (0,), [0]):
        return batched_outer(a, b)
    raise NotImplementedError("only the call pattern used by the reference (convolution.py:323-331) is restated")


def validate_contraction_modes(shape1, shape2, modes, batched_modes=False):
    """tensorly.tenalg.tenalg_utils._validate_contraction_modes (factorized_tensordot.py:4,25-31,81)."""
    if isinstance(modes, int):
        modes = [modes]
    modes = list(modes) if not isinstance(modes, range) else list(modes)
    if len(modes) == 2 and not isinstance(modes[0], (int, np.integer)) and not isinstance(modes[1], (int, np.integer)):
        modes1, modes2 = list(modes[0]), list(modes[1])
    elif len(modes) == 0:
        modes1, modes2 = [], []
    else:
        modes1 = list(modes)
        modes2 = list(modes)
    if len(modes1) != len(modes2):
        raise ValueError("Can only contract two tensors along the same number of modes")
    for i in range(len(modes1)):
        if shape1[modes1[i]] != shape2[modes2[i]]:
            raise ValueError(f"Contraction dimensions must have the same dimensions: "
                             f"{shape1[modes1[i]]} != {shape2[modes2[i]]}")
        if modes1[i] < 0:
            modes1[i] += len(shape1)
        if modes2[i] < 0:
            modes2[i] += len(shape2)
    return modes1, modes2


# --------------------------------------------------------------------------------------
# reconstruction (tensorly/{cp,tucker,tt}_tensor.py)
# --------------------------------------------------------------------------------------
def khatri_rao(matrices):
    """Column-wise Kronecker; rows ordered with the first matrix slowest (row-major)."""
    res = matrices[0]
    for m in matrices[1:]:
        res = (res[:, None, :] * m[None, :, :]).reshape(-1, res.shape[1])
    return res


def cp_to_tensor(cp_tensor):
    """T[i0..] = sum_r w_r prod_k F_k[i_k, r] (factorized_tensors.py:130)."""
    weights, factors = cp_tensor
    factors = list(factors)
    full_shape = [f.shape[0] for f in factors]
    if len(factors) == 1:
        f0 = factors[0] if weights is None else factors[0] * weights
        return f0.sum(dim=1)
    f0 = factors[0] if weights is None else factors[0] * weights
    return (f0 @ khatri_rao(factors[1:]).t()).reshape(full_shape)


def tucker_to_tensor(tucker_tensor):
    """core x_0 F_0 x_1 F_1 ... (factorized_tensors.py:308)."""
    core, factors = tucker_tensor
    return multi_mode_dot(core, list(factors))


def tt_to_tensor(factors):
    """T[i1..id] = F1[:, i1, :] F2[:, i2, :] ... (factorized_tensors.py:415)."""
    factors = list(factors)
    full_shape = [f.shape[1] for f in factors]
    res = factors[0].reshape(full_shape[0], -1)
    for f in factors[1:]:
        rank_prev = f.shape[0]
        res = res.reshape(-1, rank_prev) @ f.reshape(rank_prev, -1)
    return res.reshape(full_shape)


# --------------------------------------------------------------------------------------
# shape / rank validation
# --------------------------------------------------------------------------------------
def validate_cp_tensor(cp_tensor):
    weights, factors = cp_tensor
    factors = list(factors)
    rank = int(factors[0].shape[1])
    shp = []
    for i, f in enumerate(factors):
        if f.dim() != 2:
            raise ValueError(f"A CP tensor should be composed of matrices; factor {i} has {f.dim()} dims")
        if f.shape[1] != rank:
            raise ValueError("All the factors of a CP tensor should have the same number of columns")
        shp.append(int(f.shape[0]))
    if weights is not None and tuple(weights.shape) != (rank,):
        raise ValueError(f"weights should have shape ({rank},) but got {tuple(weights.shape)}")
    return tuple(shp), rank


def validate_tucker_tensor(tucker_tensor):
    core, factors = tucker_tensor
    factors = list(factors)
    if len(factors) < 2:
        raise ValueError("A Tucker tensor should be composed of at least two factors and a core")
    if len(factors) != core.dim():
        raise ValueError("Tucker decomposition needs one factor per core mode")
    shp, rank = [], []
    for i, f in enumerate(factors):
        cur_shape, cur_rank = f.shape
        if core.shape[i] != cur_rank:
            raise ValueError(f"Factor {i} has rank {cur_rank} but core.shape[{i}]={core.shape[i]}")
        shp.append(int(cur_shape))
        rank.append(int(cur_rank))
    return tuple(shp), tuple(rank)


def validate_tt_tensor(factors):
    factors = list(factors)
    shp, rank = [], []
    for i, f in enumerate(factors):
        if f.dim() != 3:
            raise ValueError("TT cores must be third order")
        r0, s, r1 = f.shape
        if i and factors[i - 1].shape[2] != r0:
            raise ValueError("consecutive TT cores must have matching ranks")
        shp.append(int(s))
        rank.append(int(r0))
    if factors[0].shape[0] != 1 or factors[-1].shape[2] != 1:
        raise ValueError("TT boundary ranks must be 1")
    rank.append(int(factors[-1].shape[2]))
    return tuple(shp), tuple(rank)


def _rounding(rounding):
    if rounding == "ceil":
        return math.ceil
    if rounding == "floor":
        return math.floor
    if rounding == "round":
        return round
    raise ValueError(f"Rounding should be round, floor or ceil, but got {rounding}")


def validate_cp_rank(tensor_shape, rank="same", rounding="round"):
    """float f -> int(round(prod(shape) * f / sum(shape))) (factorized_tensors.py:97,109)."""
    rfun = _rounding(rounding)
    if rank == "same":
        rank = float(1)
    if isinstance(rank, float):
        rank = int(rfun(float(np.prod(tensor_shape)) * rank / float(np.sum(tensor_shape))))
    return rank


# SURVEY.md App. A.5 states the Tucker float-rank polynomial as  N x^n + (sum s^2) x + (sum_fixed s^2) x - f N.
# tensorly's source is not available offline to confirm the leading coefficient, so it is kept switchable.
TUCKER_RANK_LEADING_COEFF_IS_FRACTION_SCALED = False


def validate_tucker_rank(tensor_shape, rank="same", rounding="round", fixed_modes=None):
    from scipy.optimize import brentq
    rfun = _rounding(rounding)
    if rank == "same":
        rank = float(1)
    if isinstance(rank, float):
        tensor_shape = list(tensor_shape)
        n_modes_compressed = len(tensor_shape)
        n_full = float(np.prod(tensor_shape))
        n_param_tensor = n_full * rank
        if fixed_modes is not None:
            fixed = [(mode, tensor_shape.pop(mode)) for mode in sorted(fixed_modes, reverse=True)][::-1]
            n_fixed_params = float(sum(s ** 2 for _, s in fixed))
            n_modes_compressed -= len(fixed)
        else:
            fixed = None
            n_fixed_params = 0.0
        squared_dims = float(sum(s ** 2 for s in tensor_shape))
        lead = n_param_tensor if TUCKER_RANK_LEADING_COEFF_IS_FRACTION_SCALED else n_full

        def fun(x):
            return lead * x ** n_modes_compressed + squared_dims * x + n_fixed_params * x - n_param_tensor

        fraction_param = brentq(fun, 0.0, max(rank, 1.0))
        rank = [max(int(rfun(s * fraction_param)), 1) for s in tensor_shape]
        if fixed is not None:
            for mode, size in fixed:
                rank.insert(mode, size)
    elif isinstance(rank, int):
        n_modes = len(tensor_shape)
        warnings.warn(f"Given only one int for 'rank' for decomposition a tensor of order {n_modes}. "
                      f"Using this rank for all modes.")
        if fixed_modes is None:
            rank = [rank] * n_modes
        else:
            rank = [rank if i not in fixed_modes else s for i, s in enumerate(tensor_shape)]
    return tuple(int(r) for r in rank)


def validate_tt_rank(tensor_shape, rank="same", constant_rank=False, rounding="round",
                     allow_overparametrization=True):
    rfun = _rounding(rounding)
    if rank == "same":
        rank = float(1)
    tensor_shape = [int(s) for s in tensor_shape]
    if isinstance(rank, float) and constant_rank:
        n_param_tensor = float(np.prod(tensor_shape)) * rank
        order = len(tensor_shape)
        if order == 2:
            rank = (1, max(int(rfun(n_param_tensor / (tensor_shape[0] + tensor_shape[1]))), 1), 1)
            warnings.warn("Determining the tt-rank for the trivial case of a matrix (order 2 tensor).")
        else:
            a = float(np.sum(tensor_shape[1:-1]))
            b = float(tensor_shape[0] + tensor_shape[-1])
            c = -n_param_tensor
            delta = math.sqrt(b ** 2 - 4 * a * c)
            solution = int(rfun((-b + delta) / (2 * a)))
            rank = (1,) + (max(solution, 1),) * (order - 1) + (1,)
        return tuple(rank)
    if isinstance(rank, float):
        order = len(tensor_shape)
        avg_dim = [(tensor_shape[i] + tensor_shape[i + 1]) / 2 for i in range(order - 1)]
        if len(avg_dim) > 1:
            a = sum(avg_dim[i - 1] * tensor_shape[i] * avg_dim[i] for i in range(1, order - 1))
        else:
            warnings.warn("Determining the tt-rank for the trivial case of a matrix (order 2 tensor).")
            a = avg_dim[0] ** 2 * tensor_shape[0]
        b = tensor_shape[0] * avg_dim[0] + tensor_shape[-1] * avg_dim[-1]
        c = -float(np.prod(tensor_shape)) * rank
        delta = math.sqrt(b ** 2 - 4 * a * c)
        fraction_param = (-b + delta) / (2 * a)
        rank = tuple(max(int(rfun(d * fraction_param)), 1) for d in avg_dim)
        return (1,) + rank + (1,)
    n_dim = len(tensor_shape)
    if isinstance(rank, int):
        rank = [1] + [rank] * (n_dim - 1) + [1]
    elif n_dim + 1 != len(rank):
        raise ValueError(f"Provided incorrect number of ranks. Should verify len(rank) == tl.ndim(tensor)+1, "
                         f"but len(rank) = {len(rank)} while tl.ndim(tensor) + 1 = {n_dim + 1}")
    if rank[0] != 1:
        raise ValueError(f"Provided rank[0] == {rank[0]} but boundary conditions dictate rank[0] == rank[-1] == 1.")
    if rank[-1] != 1:
        raise ValueError(f"Provided rank[-1] == {rank[-1]} but boundary conditions dictate rank[0] == rank[-1] == 1.")
    if allow_overparametrization:
        return tuple(int(r) for r in rank)
    validated = [1]
    for i, s in enumerate(tensor_shape[:-1]):
        n_row = int(validated[-1] * s)
        n_col = int(np.prod(tensor_shape[i + 1:]))
        validated.append(min(n_row, n_col, rank[i + 1]))
    validated.append(1)
    return tuple(validated)


def validate_tt_matrix_rank(tensorized_shape, rank="same", rounding="round"):
    """TT rank of the merged shape (o_k * i_k)_k (tensorized_matrices.py:518,527)."""
    n_dim = len(tensorized_shape) // 2
    merged = tuple(int(tensorized_shape[i] * tensorized_shape[i + n_dim]) for i in range(n_dim))
    return validate_tt_rank(merged, rank=rank, rounding=rounding)
