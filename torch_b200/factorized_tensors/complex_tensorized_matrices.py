"""Complex-valued tensorized matrices: the four real tensorized factorizations with ``ComplexHandler`` mixed in, registered under
the names ``ComplexDense / ComplexTucker / ComplexTT / ComplexCP`` of the tensorized registry (what the reference declares class by
class in tltorch/factorized_tensors/complex_tensorized_matrices.py:10-52)."""
import torch

from . import tensorized_matrices as _real
from .complex_factorized_tensors import ComplexHandler


def _complex_variant(base, registry_name, complex_params, doc):
    """Subclass ``base`` with complex storage handling; ``new`` defaults to ``torch.cfloat`` parameters."""

    def new(cls, tensorized_shape, rank=None, device=None, dtype=torch.cfloat, **kwargs):
        return super(variant, cls).new(tensorized_shape, rank, device=device, dtype=dtype, **kwargs)

    variant = type(base)("Complex" + base.__name__, (ComplexHandler, base),
                         {"__doc__": doc, "__module__": __name__, "_complex_params": list(complex_params), "new": classmethod(new)},
                         name=registry_name)
    return variant


ComplexDenseTensorized = _complex_variant(_real.DenseTensorized, "ComplexDense", ["tensor"], "Complex dense tensorized matrix.")
ComplexTuckerTensorized = _complex_variant(_real.TuckerTensorized, "ComplexTucker", ["core", "factors"],
                                           "Complex Tucker factorization of a tensorized matrix.")
ComplexBlockTT = _complex_variant(_real.BlockTT, "ComplexTT", ["factors"], "Complex block tensor-train (TT-matrix).")
ComplexCPTensorized = _complex_variant(_real.CPTensorized, "ComplexCP", ["weights", "factors"],
                                       "Complex CP factorization of a tensorized matrix.")
