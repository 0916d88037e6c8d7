"""Tensor Contraction Layer (reference: tltorch/factorized_layers/tensor_contraction_layers.py:19-94)."""
import math

from ..utils.nvtx import annotate
import torch
import torch.nn as nn
from torch.nn import init

from .. import tenalg
from ..ops import contract


class TCL(nn.Module):
    """y = x x_1 M_0 x_2 M_1 ... : every non-batch mode of x is projected by its own (rank_k, input_k) matrix.

    Reference: Kossaifi et al., "Tensor Contraction Layers for Parsimonious Deep Nets", CVPR-W 2017."""

    def __init__(self, input_shape, rank, verbose=0, bias=False, device=None, dtype=None, **kwargs):
        super().__init__(**kwargs)
        self.verbose = verbose
        self.input_shape = (input_shape,) if isinstance(input_shape, int) else tuple(input_shape)
        self.order = len(self.input_shape)
        self.rank = (rank,) * self.order if isinstance(rank, int) else tuple(rank)
        self.output_shape = self.rank            # the reference reads self.output_shape without defining it (App. C.5)
        self.contraction_modes = list(range(1, self.order + 1))
        for i, (s, r) in enumerate(zip(self.input_shape, self.rank)):
            self.register_parameter(f"factor_{i}", nn.Parameter(torch.empty((r, s), device=device, dtype=dtype)))
        if bias:
            self.bias = nn.Parameter(torch.empty(self.output_shape, device=device, dtype=dtype), requires_grad=True)
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    @property
    def factors(self):
        return [getattr(self, f"factor_{i}") for i in range(self.order)]

    @annotate(lambda m: "tltorch.TCL")
    def forward(self, x):
        # one fused launch when a sample fits shared memory (bias added in the same kernel), else a per-mode chain
        return tenalg.multi_mode_dot(x, self.factors, modes=self.contraction_modes, addend=self.bias)

    def reset_parameters(self):
        for i in range(self.order):
            init.kaiming_uniform_(getattr(self, f"factor_{i}"), a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(self.input_shape[0])
            init.uniform_(self.bias, -bound, bound)
