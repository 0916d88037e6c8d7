"""Init-time decompositions of a dense tensor into CP / Tucker / TT / TT-matrix form.

The reference gets these from its third-party dependency TensorLy (``tensorly.decomposition.parafac / tucker /
tensor_train / tensor_train_matrix``; requirements.txt lists ``tensorly`` unpinned, the API used is the 0.8 line), which is in
neither this image nor the offline wheelhouse.  What is here restates TensorLy's published algorithms with the defaults the
reference's call sites rely on:

  parafac              CP-ALS, SVD initialisation, ``l2_reg`` ridge on the normal equations, stop when the relative
                       reconstruction error moves by less than ``tol``         <- factorized_tensors.py:107-123 (CPTensor)
  tucker               HOSVD initialisation + HOOI sweeps                       <- factorized_tensors.py:246-301 (TuckerTensor)
  tensor_train         TT-SVD, left to right                                    <- factorized_tensors.py:391-408 (TTTensor)
  tensor_train_matrix  TT-SVD of the row/column-interleaved tensor              <- tensorized_matrices.py:517-534 (BlockTT)

This is host-side initialisation code that runs once per layer, outside the per-step hot path; it runs on whatever device the
dense tensor lives on through ``torch.linalg`` (fp64 SVDs of small unfoldings), not through the CUDA extension.  A
decomposition is unique only up to sign / permutation / scaling, so parity with the reference is stated on what its own tests
check (factorized_layers/tests/test_factorized_linear.py:31-58, test_factorized_convolution.py:66-74): the reconstruction
reproduces the dense tensor to the tested number of decimals.
"""
import math
import warnings

import torch

from . import tenalg

__all__ = ["svd_interface", "parafac", "tucker", "partial_tucker", "tensor_train", "tensor_train_matrix", "unfold"]


def unfold(tensor, mode):
    """Mode-``mode`` unfolding with the remaining modes in their original (C) order."""
    return torch.movedim(tensor, mode, 0).reshape(tensor.shape[mode], -1)


def _svd_flip(U, V):
    """Deterministic signs: the entry of largest magnitude in every column of U is positive (TensorLy's ``svd_flip``, u-based)."""
    n = min(U.shape[1], V.shape[0])                  # with full_matrices only the first min(rows, cols) vectors are paired
    idx = torch.argmax(U[:, :n].abs(), dim=0)
    signs = torch.sign(U[idx, torch.arange(n, device=U.device)])
    signs = torch.where(signs == 0, torch.ones_like(signs), signs)
    U, V = U.clone(), V.clone()
    U[:, :n] *= signs
    V[:n] *= signs.unsqueeze(1)
    return U, V


def svd_interface(matrix, n_eigenvecs=None):
    """Truncated SVD ``matrix ~ U[:, :k] diag(S[:k]) V[:k]`` (TensorLy ``svd_interface(method='truncated_svd')``).  When ``k``
    exceeds min(matrix.shape) the left basis is completed to ``k`` orthonormal columns (full_matrices), as TensorLy does."""
    rows, cols = matrix.shape
    min_dim, max_dim = min(rows, cols), max(rows, cols)
    k = max_dim if n_eigenvecs is None else int(n_eigenvecs)
    if k > max_dim:
        warnings.warn(f"Trying to compute SVD with n_eigenvecs={k}, which is larger than max(matrix.shape)={max_dim}. "
                      f"Setting n_eigenvecs to {max_dim}")
        k = max_dim
    U, S, V = torch.linalg.svd(matrix, full_matrices=k > min_dim)
    U, V = _svd_flip(U, V)
    return U[:, :k], S[:k], V[:k]


# =====================================================================================================================
# CP
# =====================================================================================================================
def _initialize_cp(tensor, rank, init, generator):
    if init == "random":
        return [torch.rand((s, rank), dtype=tensor.dtype, device=tensor.device, generator=generator) for s in tensor.shape]
    if init != "svd":
        raise ValueError(f'Initialization method "{init}" not recognized')
    factors = []
    for mode in range(tensor.dim()):
        U, S, _ = svd_interface(unfold(tensor, mode), n_eigenvecs=rank)
        if mode == 0:                                   # carry the tensor's scale in the first factor
            idx = min(rank, S.shape[0])
            U = U.clone()
            U[:, :idx] = U[:, :idx] * S[:idx]
        if U.shape[1] < rank:                           # more components than this mode has directions: random completion
            extra = torch.rand((U.shape[0], rank - U.shape[1]), dtype=tensor.dtype, device=tensor.device, generator=generator)
            U = torch.cat([U, extra], dim=1)
        factors.append(U[:, :rank].contiguous())
    return factors


def _mttkrp(tensor, weights, factors, mode):
    """Matricised tensor times Khatri-Rao product of all factors but ``mode``: (shape[mode], R), without forming the product."""
    order = tensor.dim()
    letters = "abcdefghijklmnopqstuvwxyz"[:order]
    operands, subs = [tensor], [letters]
    for k, f in enumerate(factors):
        if k != mode:
            operands.append(f)
            subs.append(letters[k] + "r")
    out = torch.einsum(",".join(subs) + "->" + letters[mode] + "r", *operands)
    return out * weights.unsqueeze(0)


def _fp64(tensor):
    """Decompositions iterate in fp64 whatever the parameter dtype (the reconstruction-error estimate ||T||^2 + ||rec||^2 - 2<T, rec>
    cancels catastrophically in fp32 and stops ALS early); results are cast back by the callers below."""
    wide = torch.complex128 if tensor.is_complex() else torch.float64
    return tensor.to(wide), tensor.dtype


def parafac(tensor, rank, n_iter_max=100, init="svd", tol=1e-8, l2_reg=0.0, normalize_factors=False, random_state=None,
            fixed_modes=None, return_errors=False, verbose=False, **unused):
    """CP decomposition by alternating least squares -> (weights (R), [factors (d_k, R)])."""
    tensor, out_dtype = _fp64(tensor)
    rank = tenalg.validate_cp_rank(tuple(tensor.shape), rank)
    gen = None
    if random_state is not None:
        gen = torch.Generator(device=tensor.device)
        gen.manual_seed(int(random_state))
    factors = _initialize_cp(tensor, rank, init, gen)
    weights = torch.ones(rank, dtype=tensor.dtype, device=tensor.device)
    ridge = torch.eye(rank, dtype=tensor.dtype, device=tensor.device) * l2_reg
    norm_tensor = torch.linalg.norm(tensor)
    modes = [m for m in range(tensor.dim()) if not fixed_modes or m not in fixed_modes]
    errors = []
    for iteration in range(n_iter_max):
        mttkrp = None
        for mode in modes:
            gram = torch.ones((rank, rank), dtype=tensor.dtype, device=tensor.device)
            for k, f in enumerate(factors):
                if k != mode:
                    gram = gram * (f.conj().t() @ f)
            gram = gram + ridge
            gram = weights.reshape(-1, 1) * gram * weights.reshape(1, -1)
            mttkrp = _mttkrp(tensor, weights, factors, mode)
            try:
                factor = torch.linalg.solve(gram.conj().t(), mttkrp.t()).t()
            except RuntimeError:                        # singular normal equations (rank-deficient sweep): least squares
                factor = torch.linalg.lstsq(gram.conj().t(), mttkrp.t()).solution.t()
            if normalize_factors:
                scales = torch.linalg.norm(factor, dim=0)
                scales = torch.where(scales == 0, torch.ones_like(scales), scales)
                weights = scales
                factor = factor / scales
            factors[mode] = factor.contiguous()
        if tol:
            gram_all = torch.ones((rank, rank), dtype=tensor.dtype, device=tensor.device)
            for f in factors:
                gram_all = gram_all * (f.conj().t() @ f)
            factors_norm_sq = (weights.reshape(-1, 1) * gram_all * weights.reshape(1, -1)).sum().abs()
            iprod = (mttkrp * factors[modes[-1]].conj()).sum()
            err = torch.sqrt(torch.abs(norm_tensor ** 2 + factors_norm_sq - 2 * iprod)) / norm_tensor
            errors.append(float(err))
            if verbose:
                print(f"parafac iteration {iteration}: reconstruction error {errors[-1]:.3e}")
            if iteration >= 1 and abs(errors[-2] - errors[-1]) < tol:
                break
    weights, factors = weights.to(out_dtype), [f.to(out_dtype) for f in factors]
    if return_errors:
        return (weights, factors), errors
    return weights, factors


# =====================================================================================================================
# Tucker
# =====================================================================================================================
def _multi_mode_dot_t(tensor, factors, modes, skip=None):
    """tensor x_m factors[i]^T for every listed mode (projection onto the factor bases)."""
    out = tensor
    for i, (f, m) in enumerate(zip(factors, modes)):
        if skip is not None and i == skip:
            continue
        out = torch.movedim(torch.tensordot(f.conj().t(), out, dims=([1], [m])), 0, m)
    return out


def partial_tucker(tensor, rank, modes=None, n_iter_max=100, init="svd", tol=1e-4, random_state=None, verbose=False, **unused):
    """Tucker decomposition along ``modes`` (HOSVD initialisation, then HOOI) -> ((core, factors), errors)."""
    if modes is None:
        modes = list(range(tensor.dim()))
    if isinstance(rank, int):
        rank = [rank] * len(modes)
    rank = [int(r) for r in rank]
    if init == "svd":
        factors = [svd_interface(unfold(tensor, m), n_eigenvecs=rank[i])[0] for i, m in enumerate(modes)]
    elif init == "random":
        gen = None
        if random_state is not None:
            gen = torch.Generator(device=tensor.device)
            gen.manual_seed(int(random_state))
        factors = [torch.rand((tensor.shape[m], rank[i]), dtype=tensor.dtype, device=tensor.device, generator=gen)
                   for i, m in enumerate(modes)]
    else:                                   # an initial (core, factors) pair
        factors = [f.clone() for f in init[1]]
    norm_tensor = torch.linalg.norm(tensor)
    errors = []
    core = _multi_mode_dot_t(tensor, factors, modes)
    for iteration in range(n_iter_max):
        for i, m in enumerate(modes):
            approx = _multi_mode_dot_t(tensor, factors, modes, skip=i)
            factors[i] = svd_interface(unfold(approx, m), n_eigenvecs=rank[i])[0]
        core = _multi_mode_dot_t(tensor, factors, modes)
        # orthonormal factors: ||T - rec||^2 = ||T||^2 - ||core||^2
        err = torch.sqrt(torch.abs(norm_tensor ** 2 - torch.linalg.norm(core) ** 2)) / norm_tensor
        errors.append(float(err))
        if verbose:
            print(f"tucker iteration {iteration}: reconstruction error {errors[-1]:.3e}")
        if iteration > 1 and tol and abs(errors[-2] - errors[-1]) < tol:
            break
    return (core, [f.contiguous() for f in factors]), errors


def tucker(tensor, rank, fixed_factors=None, n_iter_max=100, init="svd", tol=1e-4, random_state=None, verbose=False,
           return_errors=False, **unused):
    """Tucker decomposition -> (core (r_0..), [factors (d_k, r_k)])."""
    if fixed_factors:
        raise NotImplementedError("tucker(fixed_factors=...) is not used by the factorized layers")
    tensor, out_dtype = _fp64(tensor)
    rank = tenalg.validate_tucker_rank(tuple(tensor.shape), rank)
    (core, factors), errors = partial_tucker(tensor, rank, None, n_iter_max, init, tol, random_state, verbose)
    core, factors = core.to(out_dtype), [f.to(out_dtype) for f in factors]
    if return_errors:
        return (core, factors), errors
    return core, factors


# =====================================================================================================================
# TT
# =====================================================================================================================
def tensor_train(tensor, rank, verbose=False, **unused):
    """TT-SVD -> [cores (r_k, d_k, r_{k+1})].  Ranks are clipped to what each unfolding supports."""
    tensor, out_dtype = _fp64(tensor)
    shape = tuple(tensor.shape)
    rank = list(tenalg.validate_tt_rank(shape, rank))
    order = len(shape)
    unfolding = tensor
    cores = [None] * order
    for k in range(order - 1):
        unfolding = unfolding.reshape(int(rank[k] * shape[k]), -1)
        n_row, n_col = unfolding.shape
        current = min(n_row, n_col, rank[k + 1])
        U, S, V = svd_interface(unfolding, n_eigenvecs=current)
        rank[k + 1] = current
        cores[k] = U.reshape(rank[k], shape[k], rank[k + 1]).contiguous()     # (cast back to the caller's dtype at the end)
        if verbose:
            print(f"TT factor {k} computed with shape {tuple(cores[k].shape)}")
        unfolding = S.reshape(-1, 1) * V
    prev_rank, last_dim = unfolding.shape
    cores[-1] = unfolding.reshape(prev_rank, last_dim, 1).contiguous()
    return [c.to(out_dtype) for c in cores]


def tensor_train_matrix(tensor, rank, verbose=False, **unused):
    """TT-matrix (block-TT) decomposition of a tensor of shape (rows_1..rows_d, cols_1..cols_d) -> [cores (r_k, rows_k, cols_k,
    r_{k+1})]: modes interleaved to (rows_1 cols_1, rows_2 cols_2, ...), TT-SVD, cores split back."""
    order = tensor.dim()
    half = order // 2
    if half * 2 != order:
        raise ValueError(f"The tensor should have as many dimensions for inputs and outputs, but got a tensor of order {order}")
    rows, cols = tuple(tensor.shape[:half]), tuple(tensor.shape[half:])
    if half == 1:
        return [tensor.reshape(1, rows[0], cols[0], 1).contiguous()]
    perm = [i for pair in zip(range(half), range(half, order)) for i in pair]
    merged = [a * b for a, b in zip(rows, cols)]
    cores = tensor_train(tensor.permute(perm).reshape(merged), rank, verbose=verbose)
    return [c.reshape(c.shape[0], rows[k], cols[k], -1).contiguous() for k, c in enumerate(cores)]
