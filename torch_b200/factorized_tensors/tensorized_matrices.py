"""Tensorized-matrix containers: CPTensorized / TuckerTensorized / BlockTT (TT-matrix) / DenseTensorized.

``tensorized_shape`` is a tuple whose entries are either ints (plain / batch modes, e.g. the n_layers mode) or tuples
(tensorized modes); ``shape`` is the matrix shape obtained by multiplying each tuple out.  BlockTT stores one 4-way
(more generally (2 + n_entries)-way) core per tensorized position: factors[k] (r_k, rows_k, cols_k, r_{k+1}).
API of tltorch/factorized_tensors/tensorized_matrices.py, re-authored.
"""
import math
import warnings
from collections.abc import Iterable

import numpy as np
import torch
from torch import nn

from .. import tenalg, embedding_ops
from ..ops import contract
from ..utils import FactorList
from .core import TensorizedTensor, flatten_tensorized_shape
from .. import decomposition
from .factorized_tensors import CPTensor, TuckerTensor, DenseTensor, _is_int


def is_tensorized_shape(shape):
    return not all(isinstance(s, (int, np.integer)) for s in shape)


def tensorized_shape_to_shape(tensorized_shape):
    return [int(s) if isinstance(s, (int, np.integer)) else int(np.prod(s)) for s in tensorized_shape]


def _leading_int_index(indices, tensorized_shape):
    """Normalise ``indices`` to a tuple of ints that address leading plain (int) modes; None if anything else."""
    if _is_int(indices):
        indices = (indices,)
    if not isinstance(indices, Iterable):
        return None
    indices = tuple(indices)
    if len(indices) > len(tensorized_shape):
        return None
    for index, s in zip(indices, tensorized_shape):
        if not _is_int(index) or not isinstance(s, (int, np.integer)):
            return None
    return indices


class DenseTensorized(DenseTensor, TensorizedTensor, name="Dense"):
    def __init__(self, tensor, tensorized_shape, rank=None):
        tensor_shape = flatten_tensorized_shape(tensorized_shape)
        matrix_shape = tensorized_shape_to_shape(tensorized_shape)
        super().__init__(tensor.reshape(matrix_shape) if not isinstance(tensor, nn.Parameter)
                         else nn.Parameter(tensor.data.reshape(matrix_shape)), tensor_shape, rank)
        self.shape = matrix_shape
        self.order = len(tensor_shape)
        self.tensorized_shape = tensorized_shape

    @classmethod
    def new(cls, tensorized_shape, rank, device=None, dtype=None, **kwargs):
        flat = flatten_tensorized_shape(tensorized_shape)
        return cls(nn.Parameter(torch.empty(flat, device=device, dtype=dtype)), tensorized_shape, rank=rank)

    @classmethod
    def from_tensor(cls, tensor, tensorized_shape, rank="same", **kwargs):
        return cls(nn.Parameter(tensor.clone()), tensorized_shape, rank=rank)

    def to_tensor(self):
        return self.tensor

    def __getitem__(self, indices):
        lead = _leading_int_index(indices, self.tensorized_shape)
        if lead is None:
            raise NotImplementedError("DenseTensorized indexing supports integer indices on leading plain modes")
        return self.__class__(self.tensor[lead], tensorized_shape=tuple(self.tensorized_shape[len(lead):]))


class CPTensorized(CPTensor, TensorizedTensor, name="CP"):
    def __init__(self, weights, factors, tensorized_shape, rank=None):
        tensor_shape = flatten_tensorized_shape(tensorized_shape)
        super().__init__(weights, factors, tensor_shape, rank)
        self.shape = tensorized_shape_to_shape(tensorized_shape)
        self.order = len(tensor_shape)
        self.tensorized_shape = tensorized_shape

    @classmethod
    def new(cls, tensorized_shape, rank, device=None, dtype=None, **kwargs):
        flat = flatten_tensorized_shape(tensorized_shape)
        rank = tenalg.validate_cp_rank(flat, rank)
        weights = nn.Parameter(torch.empty(rank, device=device, dtype=dtype))
        factors = [nn.Parameter(torch.empty((s, rank), device=device, dtype=dtype)) for s in flat]
        return cls(weights, factors, tensorized_shape, rank=rank)

    @classmethod
    def from_tensor(cls, tensor, tensorized_shape, rank="same", **kwargs):
        """CP-ALS (fp64) of the tensorized matrix (reference: tensorized_matrices.py:124-134)."""
        rank = tenalg.validate_cp_rank(tuple(tensor.shape), rank)
        dtype = tensor.dtype
        with torch.no_grad():
            weights, factors = decomposition.parafac(tensor.to(torch.float64), rank, **kwargs)
        return cls(nn.Parameter(weights.to(dtype).contiguous()), [nn.Parameter(f.to(dtype).contiguous()) for f in factors],
                   tensorized_shape, rank=rank)

    def __getitem__(self, indices):
        """Integer indices on leading plain modes (the n_layers case, reference: tensorized_matrices.py:136-193):
        the selected factor rows are folded into ``weights``."""
        gather = embedding_ops.parse_row_gather(indices, self.tensorized_shape)
        if gather is not None:
            # embedding lookup ``w[(layer,) rows, :]``: the selected rows of the row-mode factors collapse into ONE (P, R)
            # factor; the result stays factorized (reference: tensorized_matrices.py:171-185) -- to_matrix() gives (P, D)
            lead, rows = gather
            factors, weights = list(self.factors), self.weights
            for pos, index in enumerate(lead):
                weights = contract("r,r->r", weights, factors[pos][index, :])
            factors = factors[len(lead):]
            row_shape = tuple(self.tensorized_shape[len(lead)])
            rows = embedding_ops._as_index(rows, weights.device)
            picked = embedding_ops.cp_row_factor(weights, factors[:len(row_shape)], rows, row_shape)
            return self.__class__(weights, [picked] + factors[len(row_shape):],
                                  tensorized_shape=(int(rows.numel()), tuple(self.tensorized_shape[len(lead) + 1])))
        lead = _leading_int_index(indices, self.tensorized_shape)
        if lead is None:
            raise NotImplementedError("CPTensorized indexing supports integer indices on leading plain modes and row gathers")
        factors = list(self.factors)
        weights = self.weights
        for pos, index in enumerate(lead):
            weights = contract("r,r->r", weights, factors[pos][index, :])
        return self.__class__(weights, factors[len(lead):], tensorized_shape=tuple(self.tensorized_shape[len(lead):]))


class TuckerTensorized(TensorizedTensor, TuckerTensor, name="Tucker"):
    def __init__(self, core, factors, tensorized_shape, rank=None):
        tensor_shape = flatten_tensorized_shape(tensorized_shape)
        TuckerTensor.__init__(self, core, factors, tensor_shape, rank)
        self.shape = tensorized_shape_to_shape(tensorized_shape)
        self.tensorized_shape = tensorized_shape

    @classmethod
    def new(cls, tensorized_shape, rank, device=None, dtype=None, **kwargs):
        tensor_shape = flatten_tensorized_shape(tensorized_shape)
        rank = tenalg.validate_tucker_rank(tensor_shape, rank)
        core = nn.Parameter(torch.empty(rank, device=device, dtype=dtype))
        factors = [nn.Parameter(torch.empty((s, r), device=device, dtype=dtype)) for s, r in zip(tensor_shape, rank)]
        return cls(core, factors, tensorized_shape, rank=rank)

    @classmethod
    def from_tensor(cls, tensor, tensorized_shape, rank="same", fixed_rank_modes=None, **kwargs):
        """HOSVD + HOOI of the tensorized matrix (reference: tensorized_matrices.py:219-228)."""
        rank = tenalg.validate_tucker_rank(tuple(tensor.shape), rank, fixed_modes=fixed_rank_modes)
        with torch.no_grad():
            core, factors = decomposition.tucker(tensor, rank, **kwargs)
        return cls(nn.Parameter(core.contiguous()), [nn.Parameter(f.contiguous()) for f in factors], tensorized_shape, rank=rank)

    def __getitem__(self, indices):
        """Integer indices on leading plain modes (reference: tensorized_matrices.py:230-311)."""
        gather = embedding_ops.parse_row_gather(indices, self.tensorized_shape)
        if gather is not None:
            # embedding lookup: dense (P, *column modes) tensor (the reference returns a partial-Tucker object whose only use,
            # FactorizedEmbedding.forward, reshapes it to the dense rows at once -- factorized_embedding.py:111-115)
            lead, rows = gather
            factors, core = list(self.factors), self.core
            for pos, index in enumerate(lead):
                core = tenalg.mode_dot(core, factors[pos][index, :], 0)
            factors = factors[len(lead):]
            row_shape = tuple(self.tensorized_shape[len(lead)])
            col_shape = tuple(self.tensorized_shape[len(lead) + 1])
            rows = embedding_ops._as_index(rows, core.device)
            out = embedding_ops.tucker_rows(core, factors[:len(row_shape)], factors[len(row_shape):], rows, row_shape)
            return out.reshape(out.shape[0], *col_shape)
        lead = _leading_int_index(indices, self.tensorized_shape)
        if lead is None:
            raise NotImplementedError("TuckerTensorized indexing supports integer indices on leading plain modes and row gathers")
        factors = list(self.factors)
        core = self.core
        for pos, index in enumerate(lead):
            core = tenalg.mode_dot(core, factors[pos][index, :], 0)
        return self.__class__(core, factors[len(lead):], tensorized_shape=tuple(self.tensorized_shape[len(lead):]))


def validate_block_tt_rank(tensorized_shape, rank):
    """TT rank of the per-core merged shape prod_over_entries(d_k)  (reference: tensorized_matrices.py:313-318)."""
    ndim = max(1 if isinstance(s, (int, np.integer)) else len(s) for s in tensorized_shape)
    per_entry = [(int(s),) * ndim if isinstance(s, (int, np.integer)) else tuple(s) for s in tensorized_shape]
    merged = [math.prod(e) for e in zip(*per_entry)]
    return tenalg.validate_tt_rank(merged, rank)


class BlockTT(TensorizedTensor, name="BlockTT"):
    """Block tensor-train == TT-matrix / matrix-product-operator form of a tensorized matrix."""

    def __init__(self, factors, tensorized_shape=None, rank=None):
        super().__init__()
        self.shape = tensorized_shape_to_shape(tensorized_shape)
        self.tensorized_shape = tensorized_shape
        self.rank = rank
        self.order = len(self.shape)
        self.factors = FactorList(factors)

    @classmethod
    def new(cls, tensorized_shape, rank, device=None, dtype=None, **kwargs):
        if all(isinstance(s, (int, np.integer)) for s in tensorized_shape):
            warnings.warn(f'Given a "flat" shape {tensorized_shape}. This will be considered as the shape of a '
                          "tensorized vector. If you just want a 1D tensor, use a regular Tensor-Train. ")
            ndim = 1
            per_entry = [tuple(tensorized_shape)]
            tensorized_shape = (tuple(tensorized_shape),)
        else:
            ndim = max(1 if isinstance(s, (int, np.integer)) else len(s) for s in tensorized_shape)
            per_entry = [(int(s),) * ndim if isinstance(s, (int, np.integer)) else tuple(s) for s in tensorized_shape]
        rank = validate_block_tt_rank(tensorized_shape, rank)
        core_shapes = list(zip(rank[:-1], *per_entry, rank[1:]))
        factors = [nn.Parameter(torch.empty(s, device=device, dtype=dtype)) for s in core_shapes]
        return cls(factors, tensorized_shape=tensorized_shape, rank=rank)

    @property
    def decomposition(self):
        return self.factors

    def to_tensor(self):
        return tenalg.blocktt_to_tensor(list(self.factors), self.tensorized_shape)

    def __getitem__(self, indices):
        """Integer indices on leading plain modes select that slice of every core (reference:
        tensorized_matrices.py:386-489, the ``contraction_op`` free path)."""
        gather = embedding_ops.parse_row_gather(indices, self.tensorized_shape)
        if gather is not None:
            # embedding lookup: dense (P, D) rows, the ``contract_factors`` path of the reference (tensorized_matrices.py:470-486)
            lead, rows = gather
            sel = (slice(None),) + tuple(lead)
            cores = [f[sel] for f in self.factors]
            rows = embedding_ops._as_index(rows, cores[0].device)
            return embedding_ops.blocktt_rows(cores, rows, tuple(self.tensorized_shape[len(lead)]))
        lead = _leading_int_index(indices, self.tensorized_shape)
        if lead is None:
            raise NotImplementedError("BlockTT indexing supports integer indices on leading plain modes and row gathers")
        sel = (slice(None),) + tuple(lead)
        factors = [f[sel] for f in self.factors]
        return self.__class__(factors, tuple(self.tensorized_shape[len(lead):]), self.rank)

    def normal_(self, mean=0, std=1):
        if mean != 0:
            raise ValueError(f"Currently only mean=0 is supported, but got mean={mean}")
        r = float(np.prod(self.rank))
        std_factors = (std / r) ** (1 / self.order)      # order = len(shape) = 2 for a matrix (reference quirk kept)
        with torch.no_grad():
            for factor in self.factors:
                factor.data.normal_(0, std_factors)
        return self

    @classmethod
    def from_tensor(cls, tensor, tensorized_shape, rank, **kwargs):
        """TT-matrix SVD of a tensor of shape (rows_1..rows_d, cols_1..cols_d) (reference: tensorized_matrices.py:516-524)."""
        rank = validate_block_tt_rank(tensorized_shape, rank)
        with torch.no_grad():
            factors = decomposition.tensor_train_matrix(tensor, rank, **kwargs)
        factors = [nn.Parameter(f.contiguous()) for f in factors]
        return cls(factors, tensorized_shape, tuple([f.shape[0] for f in factors] + [1]))

    def init_from_tensor(self, tensor, **kwargs):
        """Re-initialise the cores from a dense tensor (reference: tensorized_matrices.py:526-535)."""
        with torch.no_grad():
            factors = decomposition.tensor_train_matrix(tensor, self.rank, **kwargs)
        self.factors = FactorList([nn.Parameter(f.contiguous()) for f in factors])
        self.rank = tuple([f.shape[0] for f in factors] + [1])
        return self


# north_star names a "TT-matrix"/"TTMatrix" weight; the reference has no such class:
# a TT-matrix is exactly a BlockTT whose tensorized_shape entries are all tuples.  Alias provided as an extension.
TTMatrix = BlockTT
TensorizedTensor._factorizations["ttm"] = BlockTT
TensorizedTensor._factorizations["ttmatrix"] = BlockTT
