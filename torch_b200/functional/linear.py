"""Dispatch of the dense-x . factorized-W linear (reference: tltorch/functional/linear.py:15-50)."""
import numpy as np
import torch

from ..ops import contract
from .factorized_linear import linear_blocktt, linear_cp, linear_tucker


def _dense_linear(x, weight2d, bias):
    """x (..., I) . W(O, I)^T + b in one contraction launch (bias fused in the epilogue)."""
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1])
    y = contract("bi,oi->bo", x2, weight2d, bias, "o") if bias is not None else contract("bi,oi->bo", x2, weight2d)
    return y.reshape(*lead, weight2d.shape[0])


def factorized_linear(x, weight, bias=None, in_features=None, implementation="factorized"):
    """Linear layer with a dense input x and factorized weight."""
    from ..factorized_tensors import TensorizedTensor
    from ..factorized_tensors.tensorized_matrices import CPTensorized, TuckerTensorized, BlockTT
    assert implementation in {"factorized", "reconstructed"}, \
        f"Expect implementation from [factorized, reconstructed], but got {implementation}"
    if in_features is None:
        in_features = int(np.prod(x.shape[-1]))
    if not torch.is_tensor(weight):
        if isinstance(weight, TensorizedTensor):
            if implementation == "factorized":
                x_shape = tuple(x.shape[:-1]) + tuple(weight.tensorized_shape[1])
                out_shape = tuple(x.shape[:-1]) + (-1,)
                fn = None
                if isinstance(weight, CPTensorized):
                    fn = linear_cp
                elif isinstance(weight, TuckerTensorized):
                    fn = linear_tucker
                elif isinstance(weight, BlockTT):
                    fn = linear_blocktt
                if fn is not None:
                    # the reference adds the bias as a separate elementwise pass (linear.py:32-43); here it rides in
                    # the epilogue of the last contraction launch
                    return fn(x.reshape(x_shape), weight, bias=bias).reshape(out_shape)
            weight = weight.to_matrix()
        else:
            weight = weight.to_tensor()
    return _dense_linear(x, weight.reshape(-1, in_features), bias)


