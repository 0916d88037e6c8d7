"""FactorizedConv: N-D convolution with a factorized kernel (reference:
tltorch/factorized_layers/factorized_convolution.py:109-457, forward dispatch :189-208)."""
import warnings

import numpy as np
from ..utils.nvtx import annotate
import torch
import torch.nn as nn

from ..functional.convolution import _get_factorized_conv
from ..factorized_tensors import FactorizedTensor


def _ensure_list(order, value):
    if isinstance(value, int):
        return [value] * order
    assert len(value) == order
    return value


def _ensure_array(layers_shape, order, value, one_per_order=True):
    """Per-(sub-)layer array of stride / padding / dilation (shape layers_shape + (order,)) or flags (layers_shape)."""
    target = tuple(layers_shape) + ((order,) if one_per_order else ())
    if isinstance(value, np.ndarray):
        assert value.shape == target
        return value
    out = np.ones(target, dtype=np.int32)
    if isinstance(value, (int, bool, np.integer)):
        return out * int(value)
    assert len(value) == order
    out[..., :] = value
    return out


def kernel_shape_to_factorization_shape(factorization, kernel_shape):
    """TT kernels are factorized as (C_in, k_1..k_d, C_out): the output channel moves last (:71-82)."""
    if factorization.lower() == "tt":
        kernel_shape = list(kernel_shape)
        return tuple(kernel_shape[1:] + kernel_shape[:1])
    return kernel_shape


def factorization_shape_to_kernel_shape(factorization, factorization_shape):
    if factorization.lower() == "tt":
        shape = list(factorization_shape)
        return tuple(shape[-1:] + shape[:-1])
    return factorization_shape


def kernel_to_tensor(factorization, kernel):
    return torch.movedim(kernel, 0, -1) if factorization.lower() == "tt" else kernel


def tensor_to_kernel(factorization, tensor):
    return torch.movedim(tensor, -1, 0) if factorization.lower() == "tt" else tensor


class FactorizedConv(nn.Module):
    """Factorized convolution of arbitrary order (1-D, 2-D, 3-D)."""
    _version = 1

    def __init__(self, in_channels, out_channels, kernel_size, order=None, stride=1, padding=0, dilation=1, bias=False,
                 has_bias=False, n_layers=1, factorization="cp", rank="same", implementation="factorized",
                 fixed_rank_modes=None, device=None, dtype=None):
        super().__init__()
        if isinstance(kernel_size, int):
            if order is None:
                raise ValueError("If int given for kernel_size, order (dimension of the convolution) should also be provided.")
            if not isinstance(order, int) or order <= 0:
                raise ValueError(f"order should be the (positive integer) order of the convolution"
                                 f"but got order={order} of type {type(order)}.")
            kernel_size = (kernel_size,) * order
        else:
            kernel_size = tuple(kernel_size)
            order = len(kernel_size)
        self.order = order
        self.kernel_size = kernel_size
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.implementation = implementation
        self.input_rank = rank
        self.n_layers = n_layers
        self.factorization = factorization

        if isinstance(n_layers, int):
            layers_shape = () if n_layers == 1 else (n_layers,)
        else:
            layers_shape = tuple(n_layers)
        self.layers_shape = layers_shape
        self.padding = _ensure_array(layers_shape, order, padding)
        self.stride = _ensure_array(layers_shape, order, stride)
        self.dilation = _ensure_array(layers_shape, order, dilation)
        self.has_bias = _ensure_array(layers_shape, order, has_bias, one_per_order=False)

        if bias:
            self.bias = nn.Parameter(torch.empty(*layers_shape, out_channels, device=device, dtype=dtype))
        else:
            self.register_parameter("bias", None)

        if isinstance(factorization, FactorizedTensor):
            self.weight = factorization.to(device).to(dtype)
            kernel_shape = factorization_shape_to_kernel_shape(factorization._name, factorization.shape)
        else:
            kernel_shape = kernel_shape_to_factorization_shape(factorization, (out_channels, in_channels) + kernel_size)
            factorization_shape = layers_shape + tuple(kernel_shape)
            if fixed_rank_modes is not None:
                if factorization.lower() != "tucker":
                    warnings.warn(f"Got fixed_rank_modes={fixed_rank_modes} which is only used for factorization=tucker "
                                  f"but got factorization={factorization}.")
                elif fixed_rank_modes == "spatial":
                    fixed_rank_modes = list(range(2 + len(layers_shape), 2 + len(layers_shape) + order))
            self.weight = FactorizedTensor.new(factorization_shape, rank=rank, factorization=factorization,
                                               fixed_rank_modes=fixed_rank_modes, device=device, dtype=dtype)
        self.rank = self.weight.rank
        self.shape = self.weight.shape
        self.kernel_shape = kernel_shape
        self.forward_fun = _get_factorized_conv(self.weight, self.implementation)

    @annotate(lambda m: f"tltorch.FactorizedConv[{type(m.weight).__name__}] {m.in_channels}->{m.out_channels}")
    def forward(self, x, indices=0):
        if self.n_layers == 1:
            if indices != 0:
                raise ValueError(f"Only one convolution was parametrized (n_layers=1) but tried to access {indices}.")
            return self.forward_fun(x, self.weight(), bias=self.bias, stride=self.stride.tolist(),
                                    padding=self.padding.tolist(), dilation=self.dilation.tolist())
        if isinstance(self.n_layers, int):
            if not isinstance(indices, int):
                raise ValueError(f"Expected indices to be in int but got indices={indices}"
                                 f", but this conv was created with n_layers={self.n_layers}.")
        elif len(indices) != len(self.n_layers):
            raise ValueError(f"Got indices={indices}, but this conv was created with n_layers={self.n_layers}.")
        bias = self.bias[indices] if (self.bias is not None and self.has_bias[indices]) else None
        return self.forward_fun(x, self.weight(indices), bias=bias, stride=self.stride[indices].tolist(),
                                padding=self.padding[indices].tolist(), dilation=self.dilation[indices].tolist())

    def reset_parameters(self, std=0.02):
        if self.bias is not None:
            self.bias.data.zero_()
        self.weight = self.weight.normal_(0, std)

    def set(self, indices, stride=1, padding=0, dilation=1, bias=None):
        """Set stride / padding / dilation (and bias) of the sub-convolution ``self[indices]``."""
        self.padding[indices] = _ensure_list(self.order, padding)
        self.stride[indices] = _ensure_list(self.order, stride)
        self.dilation[indices] = _ensure_list(self.order, dilation)
        if bias is not None:
            self.bias.data[indices] = bias.data
            self.has_bias[indices] = True

    def get_conv(self, indices):
        if self.n_layers == 1:
            raise ValueError("A single convolution is parametrized, directly use the main class.")
        return SubFactorizedConv(self, indices)

    def __getitem__(self, indices):
        return self.get_conv(indices)

    @classmethod
    def from_factorization(cls, factorization, implementation="factorized", stride=1, padding=0, dilation=1, bias=None,
                           n_layers=1):
        kernel_shape = factorization_shape_to_kernel_shape(factorization._name, factorization.shape)
        if n_layers == 1:
            out_channels, in_channels, *kernel_size = kernel_shape
        elif isinstance(n_layers, int):
            layer_size, out_channels, in_channels, *kernel_size = kernel_shape
            assert layer_size == n_layers
        else:
            out_channels, in_channels, *kernel_size = kernel_shape[len(n_layers):]
        instance = cls(in_channels, out_channels, kernel_size, order=len(kernel_size), implementation=implementation,
                       padding=padding, stride=stride, bias=(bias is not None), n_layers=n_layers, dilation=dilation,
                       factorization=factorization, rank=factorization.rank)
        instance.weight = factorization
        if bias is not None:
            instance.bias.data = bias
        return instance

    @classmethod
    def from_conv(cls, conv_layer, rank="same", implementation="reconstructed", factorization="CP",
                  decompose_weights=True, decomposition_kwargs=dict(), fixed_rank_modes=None, **kwargs):
        """Factorized version of an existing nn.ConvNd.  ``decompose_weights=True`` needs a tensor decomposition
        (init-time, outside the hot path -> NotImplementedError); ``False`` gives a randomly initialised layer."""
        out_channels, in_channels, *kernel_size = conv_layer.weight.shape
        instance = cls(in_channels, out_channels, kernel_size, factorization=factorization,
                       implementation=implementation, rank=rank, dilation=conv_layer.dilation,
                       padding=conv_layer.padding, stride=conv_layer.stride[0], bias=conv_layer.bias is not None,
                       fixed_rank_modes=fixed_rank_modes, device=conv_layer.weight.device,
                       dtype=conv_layer.weight.dtype, **kwargs)
        if decompose_weights:
            if conv_layer.bias is not None:
                instance.bias.data = conv_layer.bias.data
            with torch.no_grad():
                instance.weight.init_from_tensor(kernel_to_tensor(factorization, conv_layer.weight.data),
                                                 **decomposition_kwargs)
        else:
            instance.reset_parameters()
        return instance

    @classmethod
    def from_conv_list(cls, conv_list, rank="same", implementation="reconstructed", factorization="cp",
                       decompose_weights=True, decomposition_kwargs=dict(), **kwargs):
        """Jointly factorized version of a list of nn.ConvNd layers of identical kernel shape (reference: :318-344)."""
        conv_layer = conv_list[0]
        out_channels, in_channels, *kernel_size = conv_layer.weight.shape
        instance = cls(in_channels, out_channels, kernel_size, implementation=implementation, rank=rank, factorization=factorization,
                       padding=conv_layer.padding, stride=conv_layer.stride[0], bias=True, dilation=conv_layer.dilation,
                       n_layers=len(conv_list), fixed_rank_modes=None, device=conv_layer.weight.device,
                       dtype=conv_layer.weight.dtype, **kwargs)
        if decompose_weights:
            with torch.no_grad():
                weight_tensor = torch.stack([kernel_to_tensor(factorization, layer.weight.data) for layer in conv_list])
                instance.weight.init_from_tensor(weight_tensor, **decomposition_kwargs)
        else:
            instance.reset_parameters()
        for i, layer in enumerate(conv_list):
            instance.set(i, stride=layer.stride, padding=layer.padding, dilation=layer.dilation, bias=layer.bias)
        return instance

    def transduct(self, kernel_size, mode=0, padding=0, stride=1, dilation=1, fine_tune_transduction_only=True):
        """Add a new spatial dimension to the convolution (e.g. 2-D -> 3-D) by transducting the factorization."""
        if fine_tune_transduction_only:
            for param in self.parameters():
                param.requires_grad = False
        mode += len(self.layers_shape)
        self.order += 1
        self.padding = np.insert(self.padding, mode, padding, axis=-1)
        self.stride = np.insert(self.stride, mode, stride, axis=-1)
        self.dilation = np.insert(self.dilation, mode, dilation, axis=-1)
        self.kernel_size = self.kernel_size[:mode] + (kernel_size,) + self.kernel_size[mode:]
        self.kernel_shape = self.kernel_shape[:mode + 2] + (kernel_size,) + self.kernel_shape[mode + 2:]
        # like the reference (:380-391): dispatch on the weight's TYPE (``self.factorization`` may be a tensor object when the
        # layer was built from one); a CP weight gets the centre-delta factor, so the transducted conv starts as the
        # frame-wise application of the old one (identity along the new dimension) instead of a sum over the window
        from ..factorized_tensors import CPTensor, TTTensor
        if isinstance(self.weight, CPTensor):
            ref = self.weight.factors[0]
            new_factor = torch.zeros(kernel_size, self.weight.rank, dtype=ref.dtype, device=ref.device)
            new_factor[kernel_size // 2, :] = 1
            transduction_mode = mode + 2
        elif isinstance(self.weight, TTTensor):
            new_factor, transduction_mode = None, mode + 1
        else:
            new_factor, transduction_mode = None, mode + 2
        self.weight = self.weight.transduct(kernel_size, transduction_mode, new_factor)
        self.shape = self.weight.shape
        self.forward_fun = _get_factorized_conv(self.weight, self.implementation)
        return self

    def extra_repr(self):
        s = (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, rank={self.rank}, "
             f"order={self.order}, factorization={self.weight.__class__.__name__}, implementation={self.implementation}")
        if self.n_layers != 1:
            s += f", n_layers={self.n_layers}"
        else:
            s += f", padding={self.padding.tolist()}, stride={self.stride.tolist()}, dilation={self.dilation.tolist()}"
        if self.bias is None:
            s += ", bias=False"
        return s


class SubFactorizedConv(nn.Module):
    """One of the convolutions of a jointly factorized FactorizedConv (shares the parent's parameters)."""

    def __init__(self, main_conv, indices):
        super().__init__()
        self.main_conv = main_conv
        self.indices = indices

    def forward(self, x):
        return self.main_conv.forward(x, self.indices)

    def extra_repr(self):
        return self.main_conv.extra_repr()
