"""Tensor Regression Layer (reference: tltorch/factorized_layers/tensor_regression_layers.py:16-132)."""
from ..utils.nvtx import annotate
import torch
import torch.nn as nn

from ..functional.tensor_regression import trl
from ..factorized_tensors import FactorizedTensor


class TRL(nn.Module):
    """y[b, o...] = <x[b, i...], W[i..., o...]> (+ bias) with W in factorized form.

    Reference: Tensor Regression Networks, Kossaifi et al., JMLR 2020."""

    def __init__(self, input_shape, output_shape, bias=False, verbose=0, factorization="cp", rank="same", n_layers=1,
                 device=None, dtype=None, **kwargs):
        super().__init__(**kwargs)
        self.verbose = verbose
        self.input_shape = (input_shape,) if isinstance(input_shape, int) else tuple(input_shape)
        self.output_shape = (output_shape,) if isinstance(output_shape, int) else tuple(output_shape)
        self.n_input = len(self.input_shape)
        self.n_output = len(self.output_shape)
        self.weight_shape = self.input_shape + self.output_shape
        self.order = len(self.weight_shape)
        if bias:
            self.bias = nn.Parameter(torch.empty(self.output_shape, device=device, dtype=dtype))
        else:
            self.bias = None
        if n_layers == 1:
            factorization_shape = self.weight_shape
        elif isinstance(n_layers, int):
            factorization_shape = (n_layers,) + self.weight_shape
        else:
            factorization_shape = tuple(n_layers) + self.weight_shape
        if isinstance(factorization, FactorizedTensor):
            self.weight = factorization.to(device).to(dtype)
        else:
            self.weight = FactorizedTensor.new(factorization_shape, rank=rank, factorization=factorization,
                                               device=device, dtype=dtype)
            self.init_from_random()
        self.factorization = self.weight.name

    @annotate(lambda m: "tltorch.TRL")
    def forward(self, x):
        # NB: like the reference (:75-77) the weight MODULE is passed, not weight(): hooks registered on a TRL weight
        # (tensor dropout / lasso) therefore do not fire here.  Call functional.tensor_regression.trl
        # with ``weight()`` to apply them.
        return trl(x, self.weight, bias=self.bias)

    def init_from_random(self, decompose_full_weight=False):
        with torch.no_grad():
            if decompose_full_weight:
                self.weight.init_from_tensor(torch.normal(0.0, 0.02, size=self.weight_shape))
            else:
                self.weight.normal_()
            if self.bias is not None:
                self.bias.uniform_(-1, 1)

    def init_from_linear(self, linear, unsqueezed_modes=None, **kwargs):
        """Initialise from a fully connected layer: its weight, viewed as the regression tensor, is decomposed into this layer's
        factorization (reference: tensor_regression_layers.py:96-133)."""
        if unsqueezed_modes is not None:
            if self.factorization != "Tucker":
                raise ValueError(f'unsqueezed_modes is only supported for factorization="tucker" but factorization '
                                 f"is {self.factorization}.")
            unsqueezed_modes = sorted(unsqueezed_modes)
            weight_shape = list(self.weight_shape)
            for mode in unsqueezed_modes[::-1]:
                if mode == 0:
                    raise ValueError("Cannot learn pooling for mode-0 (channels).")
                if mode > self.n_input:
                    raise ValueError(f"Can only learn pooling for the input tensor. The input has only {self.n_input} "
                                     f"modes, yet got a unsqueezed_mode for mode {mode}.")
                weight_shape.pop(mode)
            kwargs["unsqueezed_modes"] = unsqueezed_modes
        else:
            weight_shape = self.weight_shape
        with torch.no_grad():
            weight = torch.t(linear.weight).contiguous().view(weight_shape)
            self.weight.init_from_tensor(weight, **kwargs)
            if self.bias is not None:
                self.bias.data = linear.bias.data
