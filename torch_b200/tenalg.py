"""Tensor algebra of the hot path, expressed as chains of ``ops.contract`` launches (CUDA only).

These are the B200-side counterparts of the tensorly functions the reference calls: mode products,
generalised inner product, Khatri-Rao, and the CP / Tucker / TT / BlockTT reconstructions.  No function here permutes
or copies an operand: the contraction kernel reads every tensor through its strides and writes the output directly in
the requested index order.  Rank validation (host-side integer logic) lives here too.
"""
import math
import warnings

import numpy as np
import torch

from .ops import contract

_L = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


# ---------------------------------------------------------------------------------------------------------------------
# mode products (tensorly.tenalg.mode_dot / multi_mode_dot; reference call sites: functional/tensor_regression.py:40-42,
# functional/convolution.py:160, factorized_layers/tensor_contraction_layers.py:75)
# ---------------------------------------------------------------------------------------------------------------------
def mode_dot(tensor, matrix_or_vector, mode, transpose=False):
    nd = tensor.dim()
    if mode < 0:
        mode += nd
    t = _L[:nd]
    if matrix_or_vector.dim() == 2:
        new = _L[nd]
        m = (t[mode] + new) if transpose else (new + t[mode])
        contracted = matrix_or_vector.shape[0 if transpose else 1]
        if contracted != tensor.shape[mode]:
            raise ValueError(f"shapes {tuple(tensor.shape)} and {tuple(matrix_or_vector.shape)} not aligned in "
                             f"mode-{mode} multiplication")
        return contract(f"{t},{m}->{t[:mode] + new + t[mode + 1:]}", tensor, matrix_or_vector)
    if matrix_or_vector.dim() == 1:
        if matrix_or_vector.shape[0] != tensor.shape[mode]:
            raise ValueError("vector size does not match the contracted mode")
        return contract(f"{t},{t[mode]}->{t[:mode] + t[mode + 1:]}", tensor, matrix_or_vector)
    raise ValueError("Can only take n_mode_product with a vector or a matrix.")


def multi_mode_dot(tensor, matrix_or_vec_list, modes=None, skip=None, transpose=False, addend=None):
    """tensorly.tenalg.multi_mode_dot.  When the touched modes are the trailing (<= 3) dims of a contiguous tensor with
    enough leading samples, the whole chain runs as ONE fused launch with the intermediates in shared memory
    (modedot_ops / csrc/modedot.cu); otherwise mode by mode.  ``addend`` (shaped like the trailing output modes, e.g. a
    bias) is added in the last launch."""
    if modes is None:
        modes = range(len(matrix_or_vec_list))
    modes = list(modes)
    pairs = [(m, mode) for i, (m, mode) in enumerate(zip(matrix_or_vec_list, modes)) if skip is None or i != skip]
    if pairs and all(m.dim() == 2 for m, _ in pairs):
        from . import modedot_ops
        p = modedot_ops.plan(tensor, [m for m, _ in pairs], [mode for _, mode in pairs], transpose)
        if p is not None and (addend is None or _addend_matches(tensor, p, addend)):
            return modedot_ops.fused_multi_mode_dot(tensor, *p, addend=addend)
        if p is None and addend is None:
            res = _trailing_modes_first(tensor, pairs, transpose)
            if res is not None:
                return res
    if addend is not None:
        # unfused chain: all but the last product, then the last one with the addend in its epilogue
        res = multi_mode_dot(tensor, [m for m, _ in pairs[:-1]], [mode for _, mode in pairs[:-1]], transpose=transpose) if len(pairs) > 1 else tensor
        m, mode = pairs[-1]
        nd = res.dim()
        mode = mode + nd if mode < 0 else mode
        t = _L[:nd]
        new = _L[nd]
        out = t[:mode] + new + t[mode + 1:]
        mspec = (t[mode] + new) if transpose else (new + t[mode])
        letters = out[nd - addend.dim():]
        return contract(f"{t},{mspec}->{out}", res, m, addend, letters)
    removed = 0
    res = tensor
    for i, (m, mode) in enumerate(zip(matrix_or_vec_list, modes)):
        if skip is not None and i == skip:
            continue
        res = mode_dot(res, m, mode - removed, transpose=transpose)
        if m.dim() == 1:
            removed += 1
    return res


def _trailing_modes_first(tensor, pairs, transpose):
    """(batch, C, s_1..s_n) x_1 M_0 x_2 M_1 ... with a LARGE channel mode and a small odd-sized map -- tucker_trl's input
    projection on (B, 2048, 7, 7) feature maps (functional/tensor_regression.py:40), TCL on the same maps.

    The channel mode is the expensive one (it reads the whole activation) and its natural GEMM needs rows of contiguous
    channels: the map is transposed once to channels-last rows (rows_ops), the channel mode runs as ONE K-major tcgen05
    GEMM (batch * spatial, C) x (C, r_0), and the spatial modes then act on the r_0-channel result, which is C / r_0
    times smaller.  Returns None when the shapes do not fit this plan."""
    import os
    import torch
    from . import rows_ops, spatial_ops
    from .tucker_ops import padded_linear
    nd = tensor.dim()
    if nd < 3 or not tensor.is_contiguous() or tensor.dtype != torch.bfloat16:
        return None
    if [mode + nd if mode < 0 else mode for _, mode in pairs] != list(range(1, nd)):
        return None
    if any(m.dtype != tensor.dtype or m.dim() != 2 for m, _ in pairs):
        return None
    mats = [m for m, _ in pairs]
    b, c, spatial = tensor.shape[0], tensor.shape[1], list(tensor.shape[2:])
    m0 = mats[0].t() if transpose else mats[0]                                   # (r_0, C)
    r0 = m0.shape[0]
    if c < 256 or r0 * 4 > c or b * math.prod(spatial) < 1024 or b > 65535:
        return None
    S = math.prod(spatial)
    oriented = [m if transpose else m.t() for m in mats[1:]]                     # (d_i, q_i)
    J = math.prod(m.shape[1] for m in oriented)
    if os.environ.get("TLT_SPATIAL_FIRST", "on") != "off" and len(oriented) in (1, 2, 3) and spatial_ops.eligible(tensor, S, J):
        # spatial modes first: ONE pass over the activation where it lies (no transposition): t[b, (q..), c], then the channel
        # mode as a K-major tcgen05 GEMM on the S / J times smaller, channels-last t
        k = oriented[0]
        for m in oriented[1:]:
            from .trl_ops import kron2                                           # bf16 here: the two-kernel Kronecker op
            k = kron2(k, m)
        t = spatial_ops.spatial_project(tensor.reshape(b, c, S), k)              # (B, J, C)
        z = padded_linear(t.reshape(b * J, c), m0, None, True)                   # (B J, round8(r_0)), pad columns zero
        z = z.reshape(b, *[m.shape[1] for m in oriented], z.shape[1])
        return z[..., :r0].movedim(-1, 1)
    rows = rows_ops.nchw_to_rows(tensor)                                         # (B S, round8(C))
    z = padded_linear(rows, m0, None, True)                                      # (B S, round8(r_0)), pad columns zero
    z = z.reshape(b, *spatial, z.shape[1])
    z = multi_mode_dot(z, mats[1:], modes=list(range(1, nd - 1)), transpose=transpose)     # small: (B, q_1.., round8(r_0))
    return z[..., :r0].movedim(-1, 1)                                            # (B, r_0, q_1, ...), a strided view


def _addend_matches(tensor, plan, addend):
    lead, d, r, mats, flags = plan
    n_tail = tensor.dim() - len(lead)
    return list(addend.shape) == list(r[3 - n_tail:])


def inner(a, b, n_modes=None, bias=None):
    """Generalised inner product over the last ``n_modes`` of ``a`` and first ``n_modes`` of ``b`` (tenalg.inner,
    functional/tensor_regression.py:32-49); an optional bias shaped like b's trailing modes is fused into the epilogue."""
    if n_modes is None:
        if a.shape != b.shape:
            raise ValueError("Taking a generalised product between two tensors without specifying common modes is "
                             "equivalent to taking inner product. Shapes must match.")
        n_modes = a.dim()
    if tuple(a.shape[a.dim() - n_modes:]) != tuple(b.shape[:n_modes]):
        raise ValueError(f"Incorrect shapes for inner product along {n_modes} common modes: "
                         f"{tuple(a.shape)}, {tuple(b.shape)}")
    na, nb = a.dim(), b.dim()
    lead = _L[:na - n_modes]
    common = _L[na - n_modes:na]
    tail = _L[na:na + nb - n_modes]
    spec = f"{lead + common},{common + tail}->{lead + tail}"
    if bias is not None:
        return contract(spec, a, b, bias, tail[len(tail) - bias.dim():])
    return contract(spec, a, b)


def batched_outer(a, b):
    """out[r, *a_rest, *b_rest] = a[r, ...] * b[r, ...]  (tenalg.tensordot(modes=(), batched_modes=0), convolution.py:323-331)."""
    la = _L[1:a.dim()]
    lb = _L[a.dim():a.dim() + b.dim() - 1]
    return contract(f"a{la},a{lb}->a{la}{lb}", a, b)


def khatri_rao_tensor(factors, weights=None):
    """Column-wise Kronecker of (d_k, R) matrices kept as a (d_1, ..., d_n, R) tensor (row-major = first factor slowest),
    optionally scaled by ``weights`` (R).  One element-wise launch (kr_ops) for up to 4 factors."""
    from . import kr_ops
    factors = list(factors)
    if kr_ops.eligible(factors, weights):
        return kr_ops.khatri_rao(factors, weights).reshape(*[f.shape[0] for f in factors], factors[0].shape[1])
    if weights is not None:
        factors = [contract("ar,r->ar", factors[0], weights)] + factors[1:]
    res = factors[0]
    for i, f in enumerate(factors[1:], start=1):
        dims = _L[:i]
        res = contract(f"{dims}z,{_L[i]}z->{dims}{_L[i]}z", res, f)
    return res


# ---------------------------------------------------------------------------------------------------------------------
# reconstructions
# ---------------------------------------------------------------------------------------------------------------------
def cp_to_tensor(weights, factors):
    """T[i0..] = sum_r w_r prod_k F_k[i_k, r]   (CPTensor.to_tensor, factorized_tensors/factorized_tensors.py:129-130)."""
    factors = list(factors)
    if len(factors) == 1:
        f0 = factors[0] if weights is None else contract("ar,r->ar", factors[0], weights)
        ones = torch.ones((), dtype=f0.dtype, device=f0.device).expand(f0.shape[1])
        return contract("ar,r->a", f0, ones)
    kr = khatri_rao_tensor(factors[1:], weights)                 # lambda rides in the element-wise Khatri-Rao launch
    rest = _L[1:len(factors)]
    return contract(f"az,{rest}z->a{rest}", factors[0], kr)


def tucker_to_tensor(core, factors):
    """core x_0 F_0 x_1 F_1 ...   (TuckerTensor.to_tensor, factorized_tensors.py:307-308)."""
    return multi_mode_dot(core, list(factors))


def tt_to_tensor(cores):
    """T[i1..id] = F1[:, i1, :] ... Fd[:, id, :]   (TTTensor.to_tensor, factorized_tensors.py:414-415)."""
    cores = list(cores)
    res = cores[0]                                      # (1, i1, r1)
    for k, f in enumerate(cores[1:], start=1):
        lead = _L[:k + 1]                               # boundary rank + k modes
        res = contract(f"{lead}y,y{_L[k + 1]}z->{lead}{_L[k + 1]}z", res, f)
    return res.reshape([c.shape[1] for c in cores])


def blocktt_to_tensor(cores, tensorized_shape):
    """BlockTT.to_tensor (factorized_tensors/tensorized_matrices.py:354-384).

    cores[k]: (r_k, d0_k, d1_k, ..., r_{k+1}).  Plain-int entries of ``tensorized_shape`` are modes shared by all
    cores (batched); tuple entries are multiplied out core by core.  Each step writes its result directly in the final
    (entry-major, core-minor) index order, so no reshuffling pass over the reconstructed weight is ever needed."""
    cores = list(cores)
    n_entries = len(tensorized_shape)
    if (len(cores) == 3 and n_entries == 2 and all(c.dim() == 4 for c in cores)
            and not any(isinstance(s, int) for s in tensorized_shape) and cores[0].shape[0] == 1 and cores[2].shape[3] == 1
            and not cores[0].is_complex()):
        from . import blocktt_ops
        rows, cols = [c.shape[1] for c in cores], [c.shape[2] for c in cores]
        if blocktt_ops.blocktt3_fits(rows, cols, cores[0].shape[3], cores[1].shape[3], True):
            # one fused launch: the rank-indexed intermediate stays in shared memory (csrc/blocktt.cu)
            return blocktt_ops.blocktt3_to_matrix(*cores).reshape(rows + cols)
    shared = {j: _L[j] for j, s in enumerate(tensorized_shape) if isinstance(s, int)}
    nxt = n_entries

    def fresh():
        nonlocal nxt
        nxt += 1
        return _L[nxt - 1]

    groups = [[shared[j]] if j in shared else [] for j in range(n_entries)]
    res = None
    for k, core in enumerate(cores):
        mine = [shared[j] if j in shared else fresh() for j in range(n_entries)]
        core_spec = "Y" + "".join(mine) + "Z"
        if res is None:
            for j in range(n_entries):
                if j not in shared:
                    groups[j].append(mine[j])
            res = core                                   # (r0=1, d..., r1): letters X? handled through spec below
            res_spec = "X" + "".join(mine) + "Y"
            continue
        for j in range(n_entries):
            if j not in shared:
                groups[j].append(mine[j])
        out_spec = "X" + "".join("".join(g) for g in groups) + "Z"
        res = contract(f"{res_spec},{core_spec}->{out_spec}", res, core)
        res_spec = out_spec[:-1] + "Y"
    flat = sum([(e,) if isinstance(e, int) else tuple(e) for e in tensorized_shape], ())
    return res.reshape(flat)


# ---------------------------------------------------------------------------------------------------------------------
# rank validation (host-side integer logic; semantics per SURVEY.md App. A.5)
# ---------------------------------------------------------------------------------------------------------------------
def _round_fn(rounding):
    try:
        return {"ceil": math.ceil, "floor": math.floor, "round": round}[rounding]
    except KeyError:
        raise ValueError(f"Rounding should be round, floor or ceil, but got {rounding}")


def validate_cp_rank(tensor_shape, rank="same", rounding="round"):
    if rank == "same":
        rank = 1.0
    if isinstance(rank, float):
        n_full = float(np.prod(tensor_shape))
        rank = int(_round_fn(rounding)(n_full * rank / float(np.sum(tensor_shape))))
    return rank


def validate_tucker_rank(tensor_shape, rank="same", rounding="round", fixed_modes=None):
    if rank == "same":
        rank = 1.0
    if isinstance(rank, float):
        from scipy.optimize import brentq
        rf = _round_fn(rounding)
        free = list(tensor_shape)
        budget = float(np.prod(free)) * rank
        n_full = float(np.prod(free))
        pinned = []
        if fixed_modes is not None:
            for mode in sorted(fixed_modes, reverse=True):
                pinned.insert(0, (mode, free.pop(mode)))
        sq_free = float(sum(s * s for s in free))
        sq_pinned = float(sum(s * s for _, s in pinned))
        n_free = len(free)
        frac = brentq(lambda x: n_full * x ** n_free + (sq_free + sq_pinned) * x - budget, 0.0, max(rank, 1.0))
        out = [max(int(rf(s * frac)), 1) for s in free]
        for mode, s in pinned:
            out.insert(mode, s)
        return tuple(out)
    if isinstance(rank, int):
        warnings.warn(f"Given only one int for 'rank' for decomposition a tensor of order {len(tensor_shape)}. "
                      "Using this rank for all modes.")
        if fixed_modes is None:
            return (rank,) * len(tensor_shape)
        return tuple(s if i in fixed_modes else rank for i, s in enumerate(tensor_shape))
    return tuple(int(r) for r in rank)


def validate_tt_rank(tensor_shape, rank="same", constant_rank=False, rounding="round", allow_overparametrization=True):
    rf = _round_fn(rounding)
    shape = [int(s) for s in tensor_shape]
    order = len(shape)
    if rank == "same":
        rank = 1.0
    if isinstance(rank, float):
        budget = float(np.prod(shape)) * rank
        if constant_rank:
            if order == 2:
                return (1, max(int(rf(budget / (shape[0] + shape[1]))), 1), 1)
            a, b = float(sum(shape[1:-1])), float(shape[0] + shape[-1])
            r = max(int(rf((-b + math.sqrt(b * b + 4 * a * budget)) / (2 * a))), 1)
            return (1,) + (r,) * (order - 1) + (1,)
        avg = [(shape[i] + shape[i + 1]) / 2 for i in range(order - 1)]
        if len(avg) > 1:
            a = sum(avg[i - 1] * shape[i] * avg[i] for i in range(1, order - 1))
        else:
            warnings.warn("Determining the tt-rank for the trivial case of a matrix (order 2 tensor).")
            a = avg[0] ** 2 * shape[0]
        b = shape[0] * avg[0] + shape[-1] * avg[-1]
        frac = (-b + math.sqrt(b * b + 4 * a * budget)) / (2 * a)
        return (1,) + tuple(max(int(rf(d * frac)), 1) for d in avg) + (1,)
    if isinstance(rank, int):
        rank = [1] + [rank] * (order - 1) + [1]
    elif order + 1 != len(rank):
        raise ValueError(f"Provided incorrect number of ranks. Should verify len(rank) == tl.ndim(tensor)+1, "
                         f"but len(rank) = {len(rank)} while tl.ndim(tensor) + 1 = {order + 1}")
    if rank[0] != 1:
        raise ValueError(f"Provided rank[0] == {rank[0]} but boundary conditions dictate rank[0] == rank[-1] == 1.")
    if rank[-1] != 1:
        raise ValueError(f"Provided rank[-1] == {rank[-1]} but boundary conditions dictate rank[0] == rank[-1] == 1.")
    if allow_overparametrization:
        return tuple(int(r) for r in rank)
    out = [1]
    for i, s in enumerate(shape[:-1]):
        out.append(min(int(out[-1] * s), int(np.prod(shape[i + 1:])), rank[i + 1]))
    out.append(1)
    return tuple(out)


def cp_shape_rank(weights, factors):
    factors = list(factors)
    rank = int(factors[0].shape[1])
    for i, f in enumerate(factors):
        if f.dim() != 2 or f.shape[1] != rank:
            raise ValueError(f"CP factor {i} must be a matrix with {rank} columns, got {tuple(f.shape)}")
    if weights is not None and tuple(weights.shape) != (rank,):
        raise ValueError(f"CP weights must have shape ({rank},), got {tuple(weights.shape)}")
    return tuple(int(f.shape[0]) for f in factors), rank


def tucker_shape_rank(core, factors):
    factors = list(factors)
    if len(factors) != core.dim():
        raise ValueError("Tucker decomposition needs one factor per core mode")
    for i, f in enumerate(factors):
        if f.dim() != 2 or f.shape[1] != core.shape[i]:
            raise ValueError(f"Tucker factor {i} of shape {tuple(f.shape)} does not match core mode of size {core.shape[i]}")
    return tuple(int(f.shape[0]) for f in factors), tuple(int(s) for s in core.shape)


def tt_shape_rank(cores):
    cores = list(cores)
    for i, f in enumerate(cores):
        if f.dim() != 3:
            raise ValueError("TT cores must be third-order tensors")
        if i and cores[i - 1].shape[2] != f.shape[0]:
            raise ValueError("consecutive TT cores must have matching ranks")
    if cores[0].shape[0] != 1 or cores[-1].shape[2] != 1:
        raise ValueError("TT boundary ranks must be 1")
    return tuple(int(f.shape[1]) for f in cores), tuple(int(f.shape[0]) for f in cores) + (1,)
