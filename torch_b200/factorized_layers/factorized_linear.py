"""FactorizedLinear: nn.Module wrapper owning a tensorized, factorized weight (reference:
tltorch/factorized_layers/factorized_linear.py:15-280, forward dispatch :96-123)."""
import math

import numpy as np
from ..utils.nvtx import annotate
import torch
from torch import nn
from torch.utils import checkpoint

from ..functional import factorized_linear
from ..factorized_tensors import TensorizedTensor
from ..utils import get_tensorized_shape


class FactorizedLinear(nn.Module):
    """Fully-connected layer whose (out, in) weight is tensorized to ``out_tensorized_features + in_tensorized_features``
    and stored in factorized form (``factorization`` in {'cp', 'tucker', 'blocktt', 'dense'}).

    implementation='factorized' contracts the input with the factors directly; 'reconstructed' rebuilds the dense
    weight and runs a regular linear.  ``n_layers > 1`` parametrizes several layers with one jointly factorized tensor.
    """

    def __init__(self, in_tensorized_features, out_tensorized_features, bias=True, factorization="cp", rank="same",
                 implementation="factorized", n_layers=1, checkpointing=False, device=None, dtype=None):
        super().__init__()
        if factorization == "TTM" and n_layers != 1:
            raise ValueError(f"TTM factorization only support single factorized layers but got n_layers={n_layers}.")
        self.in_features = int(np.prod(in_tensorized_features))
        self.out_features = int(np.prod(out_tensorized_features))
        self.in_tensorized_features = in_tensorized_features
        self.out_tensorized_features = out_tensorized_features
        self.tensorized_shape = out_tensorized_features + in_tensorized_features
        self.weight_shape = (self.out_features, self.in_features)
        self.input_rank = rank
        self.implementation = implementation
        self.checkpointing = checkpointing
        self.n_layers = n_layers

        if bias:
            if n_layers == 1:
                self.bias = nn.Parameter(torch.empty(self.out_features, device=device, dtype=dtype))
                self.has_bias = True
            else:
                self.bias = nn.Parameter(torch.empty((n_layers, self.out_features), device=device, dtype=dtype))
                self.has_bias = np.zeros(n_layers)
        else:
            self.register_parameter("bias", None)

        if n_layers > 1:
            tensor_shape = (n_layers, out_tensorized_features, in_tensorized_features)
        else:
            tensor_shape = (out_tensorized_features, in_tensorized_features)
        if isinstance(factorization, TensorizedTensor):
            self.weight = factorization.to(device).to(dtype)
        else:
            self.weight = TensorizedTensor.new(tensor_shape, rank=rank, factorization=factorization, device=device,
                                               dtype=dtype)
            self.reset_parameters()
        self.rank = self.weight.rank

    def reset_parameters(self):
        with torch.no_grad():
            self.weight.normal_(0, math.sqrt(5) / math.sqrt(self.in_features))
            if self.bias is not None:
                bound = 1 / math.sqrt(self.in_features)
                torch.nn.init.uniform_(self.bias, -bound, bound)

    @annotate(lambda m: f"tltorch.FactorizedLinear[{type(m.weight).__name__}] {m.in_features}->{m.out_features}")
    def forward(self, x, indices=0):
        if self.n_layers == 1:
            if indices != 0:
                raise ValueError(f"Only one convolution was parametrized (n_layers=1) but tried to access {indices}.")
            weight, bias = self.weight(), self.bias
        elif isinstance(self.n_layers, int):
            if not isinstance(indices, int):
                raise ValueError(f"Expected indices to be in int but got indices={indices}"
                                 f", but this conv was created with n_layers={self.n_layers}.")
            weight = self.weight(indices)
            bias = self.bias[indices] if self.bias is not None else None
        elif len(indices) != len(self.n_layers):
            raise ValueError(f"Got indices={indices}, but this conv was created with n_layers={self.n_layers}.")
        else:
            weight = self.weight(indices)
            bias = self.bias[indices] if self.bias is not None else None

        def _inner_forward(inp):   # weight() stays outside so hooks do not fire twice under recomputation
            return factorized_linear(inp, weight, bias=bias, in_features=self.in_features,
                                     implementation=self.implementation)

        if self.checkpointing and x.requires_grad:
            return checkpoint.checkpoint(_inner_forward, x, use_reentrant=False)
        return _inner_forward(x)

    def get_linear(self, indices):
        if self.n_layers == 1:
            raise ValueError("A single linear is parametrized, directly use the main class.")
        return SubFactorizedLinear(self, indices)

    def __getitem__(self, indices):
        return self.get_linear(indices)

    @classmethod
    def from_linear(cls, linear, rank="same", auto_tensorize=True, n_tensorized_modes=3, in_tensorized_features=None,
                    out_tensorized_features=None, bias=True, factorization="CP", implementation="reconstructed",
                    checkpointing=False, decomposition_kwargs=dict(), verbose=False):
        """Create a FactorizedLinear whose weight approximates an existing nn.Linear (needs a tensor decomposition:
        init-time code outside the hot path -- raises NotImplementedError unless factorization='dense')."""
        out_features, in_features = linear.weight.shape
        if auto_tensorize:
            if out_tensorized_features is not None and in_tensorized_features is not None:
                raise ValueError("Either use auto_reshape or specify out_tensorized_features and in_tensorized_features.")
            in_tensorized_features, out_tensorized_features = get_tensorized_shape(
                in_features=in_features, out_features=out_features, order=n_tensorized_modes, min_dim=2, verbose=verbose)
        else:
            assert out_features == np.prod(out_tensorized_features)
            assert in_features == np.prod(in_tensorized_features)
        instance = cls(in_tensorized_features, out_tensorized_features, bias=bias, factorization=factorization,
                       rank=rank, implementation=implementation, n_layers=1, checkpointing=checkpointing,
                       device=linear.weight.device, dtype=linear.weight.dtype)
        instance.weight.init_from_matrix(linear.weight.data, **decomposition_kwargs)
        if bias and linear.bias is not None:
            instance.bias.data = linear.bias.data
        return instance

    @classmethod
    def from_linear_list(cls, linear_list, in_tensorized_features, out_tensorized_features, rank, bias=True,
                         factorization="CP", implementation="reconstructed", checkpointing=False,
                         decomposition_kwargs=dict(init="random")):
        if factorization == "TTM" and len(linear_list) > 1:
            raise ValueError(f"TTM factorization only support single factorized layers but got {len(linear_list)} layers.")
        for linear in linear_list:
            out_features, in_features = linear.weight.shape
            assert out_features == np.prod(out_tensorized_features)
            assert in_features == np.prod(in_tensorized_features)
        instance = cls(in_tensorized_features, out_tensorized_features, bias=bias, factorization=factorization,
                       rank=rank, implementation=implementation, n_layers=len(linear_list), checkpointing=checkpointing,
                       device=linear.weight.device, dtype=linear.weight.dtype)
        instance.weight.init_from_matrix(torch.stack([layer.weight.data for layer in linear_list]), **decomposition_kwargs)
        if bias:
            for i, layer in enumerate(linear_list):
                if layer.bias is not None:
                    instance.bias.data[i] = layer.bias.data
                    instance.has_bias[i] = 1
        return instance

    def __repr__(self):
        msg = (f"{self.__class__.__name__}(in_features={self.in_features}, out_features={self.out_features},"
               f" weight of size ({self.out_features}, {self.in_features}) tensorized to "
               f"({self.out_tensorized_features}, {self.in_tensorized_features}),"
               f"factorization={self.weight._name}, rank={self.rank}, implementation={self.implementation}")
        if self.bias is None:
            msg += ", bias=False"
        if self.n_layers == 1:
            return msg + ", with a single layer parametrized, "
        return msg + f" with {self.n_layers} layers jointly parametrized."


class SubFactorizedLinear(nn.Module):
    """One of the linears of a jointly factorized FactorizedLinear (shares the parent's parameters)."""

    def __init__(self, main_linear, indices):
        super().__init__()
        self.main_linear = main_linear
        self.indices = indices

    def forward(self, x):
        return self.main_linear(x, self.indices)

    def extra_repr(self):
        msg = f"in_features={self.main_linear.in_features}, out_features={self.main_linear.out_features}"
        if self.main_linear.has_bias[self.indices]:
            msg += ", bias=True"
        return msg
