"""Complex-valued factorized tensors (reference: tltorch/factorized_tensors/complex_factorized_tensors.py:13-80).

``ComplexHandler`` keeps every parameter / buffer as a REAL tensor with a trailing dimension of 2 (so optimisers, DDP buckets and
checkpoints only ever see real tensors) and hands it out as a complex view on attribute access; factor lists become
``ComplexFactorList``.  The contractions behind ``to_tensor`` / indexing run through ``torch_b200.ops.contract``, which splits
complex operands into real and imaginary planes for the real CUDA kernels.
"""
import torch
from torch import nn

from ..utils.parameter_list import FactorList, ComplexFactorList
from .factorized_tensors import TuckerTensor, CPTensor, TTTensor, DenseTensor


class ComplexHandler:
    def __setattr__(self, key, value):
        if isinstance(value, FactorList):
            value = value if isinstance(value, ComplexFactorList) else ComplexFactorList(list(value))
            super().__setattr__(key, value)
        elif isinstance(value, nn.Parameter):
            self.register_parameter(key, value)
        elif torch.is_tensor(value):
            self.register_buffer(key, value)
        else:
            super().__setattr__(key, value)

    def __getattr__(self, key):
        value = super().__getattr__(key)
        if torch.is_tensor(value) and not value.is_complex() and value.dim() and value.shape[-1] == 2:
            value = torch.view_as_complex(value)
        return value

    def register_parameter(self, key, value):
        if value is not None and value.is_complex():
            value = nn.Parameter(torch.view_as_real(value.detach()), requires_grad=value.requires_grad)
        super().register_parameter(key, value)

    def register_buffer(self, key, value, *args, **kwargs):
        if value is not None and value.is_complex():
            value = torch.view_as_real(value)
        super().register_buffer(key, value, *args, **kwargs)


class ComplexDenseTensor(ComplexHandler, DenseTensor, name="ComplexDense"):
    """Complex dense tensor."""

    @classmethod
    def new(cls, shape, rank=None, device=None, dtype=torch.cfloat, **kwargs):
        return super().new(shape, rank, device=device, dtype=dtype, **kwargs)


class ComplexTuckerTensor(ComplexHandler, TuckerTensor, name="ComplexTucker"):
    """Complex Tucker factorization."""

    @classmethod
    def new(cls, shape, rank="same", fixed_rank_modes=None, device=None, dtype=torch.cfloat, **kwargs):
        return super().new(shape, rank, fixed_rank_modes=fixed_rank_modes, device=device, dtype=dtype, **kwargs)


class ComplexTTTensor(ComplexHandler, TTTensor, name="ComplexTT"):
    """Complex tensor-train factorization."""

    @classmethod
    def new(cls, shape, rank="same", fixed_rank_modes=None, device=None, dtype=torch.cfloat, **kwargs):
        return super().new(shape, rank, device=device, dtype=dtype, **kwargs)


class ComplexCPTensor(ComplexHandler, CPTensor, name="ComplexCP"):
    """Complex CP factorization."""

    @classmethod
    def new(cls, shape, rank="same", fixed_rank_modes=None, device=None, dtype=torch.cfloat, **kwargs):
        return super().new(shape, rank, device=device, dtype=dtype, **kwargs)
