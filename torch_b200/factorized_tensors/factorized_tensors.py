"""CP / Tucker / TT / Dense parameter containers (API of tltorch/factorized_tensors/factorized_tensors.py, re-authored).

Storage layout in HBM (all row-major contiguous, fp32 by default):
  CPTensor      weights (R,), factors[k] (d_k, R)
  TuckerTensor  core (r_0..r_{n-1}), factors[k] (d_k, r_k)
  TTTensor      factors[k] (r_k, d_k, r_{k+1}), r_0 = r_n = 1
Reconstruction (``to_tensor``) runs on the GPU through ``torch_b200.tenalg`` (contraction kernels).
"""
import math

import numpy as np
import torch
from torch import nn

from .. import tenalg
from ..ops import contract
from ..utils import FactorList
from .core import FactorizedTensor
from .. import decomposition


def _register(module, name, value):
    if isinstance(value, nn.Parameter):
        module.register_parameter(name, value)
    else:
        module.register_buffer(name, value)


def _is_int(i):
    return isinstance(i, (int, np.integer)) and not isinstance(i, bool)


class DenseTensor(FactorizedTensor, name="Dense"):
    """Plain dense tensor behind the factorized-tensor interface."""

    def __init__(self, tensor, shape=None, rank=None):
        super().__init__()
        if shape is not None and rank is not None:
            self.shape, self.rank = shape, rank
        else:
            self.shape, self.rank = tensor.shape, None
        self.order = len(self.shape)
        _register(self, "tensor", tensor)

    @classmethod
    def new(cls, shape, rank=None, device=None, dtype=None, **kwargs):
        return cls(nn.Parameter(torch.empty(shape, device=device, dtype=dtype)))

    @classmethod
    def from_tensor(cls, tensor, rank="same", **kwargs):
        return cls(nn.Parameter(tensor.clone()))

    def init_from_tensor(self, tensor, l2_reg=1e-5, **kwargs):
        with torch.no_grad():
            self.tensor = nn.Parameter(tensor.clone())
        return self

    @property
    def decomposition(self):
        return self.tensor

    def to_tensor(self):
        return self.tensor

    def normal_(self, mean=0, std=1):
        with torch.no_grad():
            self.tensor.data.normal_(mean, std)
        return self

    def __getitem__(self, indices):
        return self.__class__(self.tensor[indices])


class CPTensor(FactorizedTensor, name="CP"):
    """CP (Kruskal) factorization: T[i_0..] = sum_r weights[r] prod_k factors[k][i_k, r]."""

    def __init__(self, weights, factors, shape=None, rank=None):
        super().__init__()
        if shape is not None and rank is not None:
            self.shape, self.rank = shape, rank
        else:
            self.shape, self.rank = tenalg.cp_shape_rank(weights, factors)
        self.order = len(self.shape)
        _register(self, "weights", weights)
        self.factors = FactorList(factors)

    @classmethod
    def new(cls, shape, rank, device=None, dtype=None, **kwargs):
        rank = tenalg.validate_cp_rank(shape, rank)
        weights = nn.Parameter(torch.empty(rank, device=device, dtype=dtype))
        factors = [nn.Parameter(torch.empty((s, rank), device=device, dtype=dtype)) for s in shape]
        return cls(weights, factors)

    @classmethod
    def from_tensor(cls, tensor, rank="same", **kwargs):
        """CP-ALS of a dense tensor in fp64, cast back (reference: factorized_tensors.py:107-116)."""
        rank = tenalg.validate_cp_rank(tuple(tensor.shape), rank)
        dtype = tensor.dtype
        with torch.no_grad():
            weights, factors = decomposition.parafac(tensor.to(torch.float64), rank, **kwargs)
        return cls(nn.Parameter(weights.to(dtype).contiguous()), [nn.Parameter(f.to(dtype).contiguous()) for f in factors])

    def init_from_tensor(self, tensor, l2_reg=1e-5, **kwargs):
        """Re-initialise the factors from a dense tensor (reference: factorized_tensors.py:118-124)."""
        with torch.no_grad():
            weights, factors = decomposition.parafac(tensor, self.rank, l2_reg=l2_reg, **kwargs)
        self.weights = nn.Parameter(weights.contiguous())
        self.factors = FactorList([nn.Parameter(f.contiguous()) for f in factors])
        return self

    @property
    def decomposition(self):
        return self.weights, self.factors

    def to_tensor(self):
        return tenalg.cp_to_tensor(self.weights, list(self.factors))

    def normal_(self, mean=0, std=1):
        super().normal_(mean, std)
        std_factors = (std / math.sqrt(self.rank)) ** (1 / self.order)
        with torch.no_grad():
            self.weights.fill_(1)
            for factor in self.factors:
                factor.data.normal_(0, std_factors)
        return self

    def __getitem__(self, indices):
        """Factorized indexing: integer indices fold the selected factor row into ``weights``; slices / index lists
        subsample the factor.  (reference: factorized_tensors.py:142-176)"""
        factors = list(self.factors)
        if _is_int(indices):
            return self.__class__(contract("r,r->r", self.weights, factors[0][indices, :]), factors[1:])
        if isinstance(indices, slice):
            return self.__class__(self.weights, [factors[0][indices, :], *factors[1:]])
        weights = self.weights
        kept = []
        for pos, index in enumerate(indices):
            if index is Ellipsis:
                raise ValueError(f"Ellipsis is not yet supported, yet got indices={indices} which contains one.")
            row = factors[pos][index, :]
            if _is_int(index):
                weights = contract("r,r->r", weights, row)
            else:
                kept.append(row)
        rest = kept + factors[len(indices):]
        if not rest:
            ones = torch.ones((), dtype=weights.dtype, device=weights.device).expand(weights.shape[0])
            return contract("r,r->", weights, ones)
        return self.__class__(weights, rest)

    def transduct(self, new_dim, mode=0, new_factor=None):
        """Add a new mode of size ``new_dim`` at position ``mode`` (all-ones factor: the tensor is repeated along it)."""
        factors = list(self.factors)
        self.order += 1
        self.shape = self.shape[:mode] + (new_dim,) + self.shape[mode:]
        if new_factor is None:
            new_factor = torch.ones(new_dim, self.rank)
        factors.insert(mode, nn.Parameter(new_factor.to(factors[0].device).contiguous()))
        self.factors = FactorList(factors)
        return self


class TuckerTensor(FactorizedTensor, name="Tucker"):
    """Tucker factorization: T = core x_0 factors[0] x_1 factors[1] ..."""

    def __init__(self, core, factors, shape=None, rank=None):
        super().__init__()
        if shape is not None and rank is not None:
            self.shape, self.rank = shape, rank
        else:
            self.shape, self.rank = tenalg.tucker_shape_rank(core, factors)
        self.order = len(self.shape)
        _register(self, "core", core)
        self.factors = FactorList(factors)

    @classmethod
    def new(cls, shape, rank, fixed_rank_modes=None, device=None, dtype=None, **kwargs):
        rank = tenalg.validate_tucker_rank(shape, rank, fixed_modes=fixed_rank_modes)
        core = nn.Parameter(torch.empty(rank, device=device, dtype=dtype))
        factors = [nn.Parameter(torch.empty((s, r), device=device, dtype=dtype)) for s, r in zip(shape, rank)]
        return cls(core, factors)

    @classmethod
    def from_tensor(cls, tensor, rank="same", fixed_rank_modes=None, **kwargs):
        """HOSVD + HOOI of a dense tensor (reference: factorized_tensors.py:246-254)."""
        rank = tenalg.validate_tucker_rank(tuple(tensor.shape), rank, fixed_modes=fixed_rank_modes)
        with torch.no_grad():
            core, factors = decomposition.tucker(tensor, rank, **kwargs)
        return cls(nn.Parameter(core.contiguous()), [nn.Parameter(f.contiguous()) for f in factors])

    def init_from_tensor(self, tensor, unsqueezed_modes=None, unsqueezed_init="average", **kwargs):
        """Re-initialise core and factors from a dense tensor (reference: factorized_tensors.py:256-301).  ``unsqueezed_modes``:
        rank-1 modes of this factorization that ``tensor`` does not have; the other modes are decomposed and the extra
        factors are filled with ``1 / size`` ('average') or the given constant."""
        if unsqueezed_modes is not None:
            unsqueezed_modes = sorted(unsqueezed_modes)
            for mode in unsqueezed_modes[::-1]:
                if self.rank[mode] != 1:
                    raise ValueError("It is only possible to initialize by averagig over mode for which rank=1."
                                     f"However, got unsqueezed_modes={unsqueezed_modes} but rank[{mode}]={self.rank[mode]} != 1.")
            rank = tuple(r for i, r in enumerate(self.rank) if i not in unsqueezed_modes)
        else:
            rank = self.rank
        with torch.no_grad():
            core, factors = decomposition.tucker(tensor, rank, **kwargs)
            if unsqueezed_modes is not None:
                for mode in unsqueezed_modes:
                    size = self.shape[mode]
                    fill = 1.0 / size if unsqueezed_init == "average" else float(unsqueezed_init)
                    factors.insert(mode, torch.full((size, 1), fill, dtype=core.dtype, device=core.device))
                    core = core.unsqueeze(mode)
        self.core = nn.Parameter(core.contiguous())
        self.factors = FactorList([nn.Parameter(f.contiguous()) for f in factors])
        return self

    @property
    def decomposition(self):
        return self.core, self.factors

    def to_tensor(self):
        return tenalg.tucker_to_tensor(self.core, list(self.factors))

    def normal_(self, mean=0, std=1):
        super().normal_(mean, std)
        r = float(np.prod([math.sqrt(r) for r in self.rank]))
        std_factors = (std / r) ** (1 / (self.order + 1))
        with torch.no_grad():
            self.core.data.normal_(0, std_factors)
            for factor in self.factors:
                factor.data.normal_(0, std_factors)
        return self

    def __getitem__(self, indices):
        """reference: factorized_tensors.py:323-359."""
        factors = list(self.factors)
        if _is_int(indices):
            return self.__class__(tenalg.mode_dot(self.core, factors[0][indices, :], 0), factors[1:])
        if isinstance(indices, slice):
            return self.__class__(self.core, [factors[0][indices, :], *factors[1:]])
        core = self.core
        kept, removed = [], 0
        for pos, index in enumerate(indices):
            if index is Ellipsis:
                raise ValueError(f"Ellipsis is not yet supported, yet got indices={indices}, indices[{pos}]={index}.")
            if _is_int(index):
                core = tenalg.mode_dot(core, factors[pos][index, :], pos - removed)
                removed += 1
            else:
                kept.append(factors[pos][index, :])
        rest = kept + factors[len(indices):]
        return self.__class__(core, rest) if rest else core


class TTTensor(FactorizedTensor, name="TT"):
    """Tensor-Train (MPS) factorization: T[i_1..i_d] = F_1[:, i_1, :] F_2[:, i_2, :] ... F_d[:, i_d, :]."""

    def __init__(self, factors, shape=None, rank=None):
        super().__init__()
        if shape is None or rank is None:
            self.shape, self.rank = tenalg.tt_shape_rank(factors)
        else:
            self.shape, self.rank = shape, rank
        self.order = len(self.shape)
        self.factors = FactorList(factors)

    @classmethod
    def new(cls, shape, rank, device=None, dtype=None, **kwargs):
        rank = tenalg.validate_tt_rank(shape, rank)
        factors = [nn.Parameter(torch.empty((rank[i], s, rank[i + 1]), device=device, dtype=dtype))
                   for i, s in enumerate(shape)]
        return cls(factors)

    @classmethod
    def from_tensor(cls, tensor, rank="same", **kwargs):
        """TT-SVD of a dense tensor (reference: factorized_tensors.py:391-400)."""
        rank = tenalg.validate_tt_rank(tuple(tensor.shape), rank)
        with torch.no_grad():
            factors = decomposition.tensor_train(tensor, rank)
        return cls([nn.Parameter(f.contiguous()) for f in factors])

    def init_from_tensor(self, tensor, **kwargs):
        """Re-initialise the cores from a dense tensor (reference: factorized_tensors.py:402-409)."""
        with torch.no_grad():
            factors = decomposition.tensor_train(tensor, self.rank)
        self.factors = FactorList([nn.Parameter(f.contiguous()) for f in factors])
        self.rank = tuple([f.shape[0] for f in factors] + [1])
        return self

    @property
    def decomposition(self):
        return self.factors

    def to_tensor(self):
        return tenalg.tt_to_tensor(list(self.factors))

    def normal_(self, mean=0, std=1):
        super().normal_(mean, std)
        r = float(np.prod(self.rank))
        std_factors = (std / r) ** (1 / self.order)
        with torch.no_grad():
            for factor in self.factors:
                factor.data.normal_(0, std_factors)
        return self

    def __getitem__(self, indices):
        """Leading-mode indexing (reference: factorized_tensors.py:428-471): an integer index contracts the selected
        slice of the first core into the next core; a slice subsamples the first core."""
        factors = list(self.factors)
        if _is_int(indices):
            indices = (indices,)
        elif isinstance(indices, slice):
            return self.__class__([factors[0][:, indices], *factors[1:]])
        carry = None                                   # (1, r) matrix of the cores contracted so far
        pos = 0
        for pos, index in enumerate(indices):
            if not _is_int(index):
                raise NotImplementedError("TTTensor indexing supports leading integer indices or one leading slice")
            sl = factors[pos][:, index, :]             # (r_k, r_{k+1})
            carry = sl if carry is None else contract("ar,rs->as", carry, sl)
        rest = factors[len(indices):]
        if not rest:
            return carry.reshape(())
        nxt = contract("ar,rds->ads", carry, rest[0])
        return self.__class__([nxt, *rest[1:]])

    def transduct(self, new_dim, mode=0, new_factor=None):
        """Add a mode whose core is the identity for every index (the tensor is repeated along the new mode)."""
        factors = list(self.factors)
        self.order += 1
        new_rank = self.rank[mode]
        self.rank = self.rank[:mode] + (new_rank,) + self.rank[mode:]
        self.shape = self.shape[:mode] + (new_dim,) + self.shape[mode:]
        if new_factor is None:
            new_factor = torch.eye(new_rank).unsqueeze(1).repeat(1, new_dim, 1)
        factors.insert(mode, nn.Parameter(new_factor.to(factors[0].device).contiguous()))
        self.factors = FactorList(factors)
        return self
