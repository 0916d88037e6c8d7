"""Complex-valued tensorized matrices (reference: tltorch/factorized_tensors/complex_tensorized_matrices.py:10-52)."""
import torch

from .tensorized_matrices import TuckerTensorized, DenseTensorized, CPTensorized, BlockTT
from .complex_factorized_tensors import ComplexHandler


class ComplexDenseTensorized(ComplexHandler, DenseTensorized, name="ComplexDense"):
    _complex_params = ["tensor"]

    @classmethod
    def new(cls, tensorized_shape, rank=None, device=None, dtype=torch.cfloat, **kwargs):
        return super().new(tensorized_shape, rank, device=device, dtype=dtype, **kwargs)


class ComplexTuckerTensorized(ComplexHandler, TuckerTensorized, name="ComplexTucker"):
    _complex_params = ["core", "factors"]

    @classmethod
    def new(cls, tensorized_shape, rank=None, device=None, dtype=torch.cfloat, **kwargs):
        return super().new(tensorized_shape, rank, device=device, dtype=dtype, **kwargs)


class ComplexBlockTT(ComplexHandler, BlockTT, name="ComplexTT"):
    _complex_params = ["factors"]

    @classmethod
    def new(cls, tensorized_shape, rank=None, device=None, dtype=torch.cfloat, **kwargs):
        return super().new(tensorized_shape, rank, device=device, dtype=dtype, **kwargs)


class ComplexCPTensorized(ComplexHandler, CPTensorized, name="ComplexCP"):
    _complex_params = ["weights", "factors"]

    @classmethod
    def new(cls, tensorized_shape, rank=None, device=None, dtype=torch.cfloat, **kwargs):
        return super().new(tensorized_shape, rank, device=device, dtype=dtype, **kwargs)
