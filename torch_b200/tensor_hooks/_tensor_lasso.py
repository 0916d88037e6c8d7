"""Generalised tensor lasso: an l1 penalty on per-rank gates of a factorized tensor, applied as a forward hook on the weight
module (reference: tltorch/tensor_hooks/_tensor_lasso.py:15-345).

    regulariser = tensor_lasso("tucker", penalty=0.01)
    regulariser.apply(layer.weight)        # adds ``lasso_weights`` parameters (ones) and the hook
    y = layer(x); loss = criterion(y) + regulariser.loss; loss.backward(); regulariser.reset()

CP: the factorization already carries ``weights``; the hook only clamps / thresholds the gates and accumulates their l1 norm.
Tucker / TT: the gates scale the columns of every factor (the trailing rank of every TT core but the last) before the layer
contracts them; the scaling is one launch of the contraction kernel per factor (``contract('ar,r->ar')``), differentiable in
both the factor and the gate.  The clamp / threshold of the few gate values and the |.| sum are plain tensor bookkeeping on
rank-sized vectors.
"""
import warnings

import torch
from torch import nn
from torch.nn import functional as F

from ..factorized_tensors import TuckerTensor, TTTensor, CPTensor
from ..ops import contract
from ..utils import ParameterList

_L = "abcdefghijklmnopqrstuvwxyz"


def _scale_last(t, gate):
    """t[..., r] * gate[r] through the contraction kernel (autograd reaches both)."""
    letters = _L[:t.dim()]
    return contract(f"{letters},{letters[-1]}->{letters}", t, gate.reshape(-1).to(t.dtype))


class TensorLasso:
    """Base class + registry: ``TensorLasso.from_factorization(tensor)`` picks the hook for the tensor's class."""
    _factorizations = dict()

    def __init_subclass__(cls, factorization, **kwargs):
        super().__init_subclass__(**kwargs)
        cls._factorizations[factorization.__name__] = cls

    def __init__(self, penalty=0.01, clamp_weights=True, threshold=1e-6, normalize_loss=True):
        self.penalty = penalty
        self.clamp_weights = clamp_weights
        self.threshold = threshold
        self.normalize_loss = normalize_loss
        self.reset()

    def reset(self):
        """Forget the accumulated loss; call once per iteration after the optimiser step."""
        self._loss = 0
        self.n_element = 0

    @property
    def loss(self):
        if self.n_element == 0:
            warnings.warn("The L1Regularization was not applied to any weights.")
            return 0
        if self.normalize_loss:
            return self.penalty * self._loss / self.n_element
        return self.penalty * self._loss

    def __call__(self, module, input, output):
        raise NotImplementedError

    def apply_lasso(self, tensor, lasso_weights):
        raise NotImplementedError

    @classmethod
    def from_factorization(cls, factorization, penalty=0.01, clamp_weights=True, threshold=1e-6, normalize_loss=True):
        for klass in type(factorization).__mro__:            # tensorized subclasses use their base factorization's hook
            if klass.__name__ in cls._factorizations:
                return cls.from_factorization_name(klass.__name__, penalty=penalty, clamp_weights=clamp_weights,
                                                   threshold=threshold, normalize_loss=normalize_loss)
        raise KeyError(factorization.__class__.__name__)

    @classmethod
    def from_factorization_name(cls, factorization_name, penalty=0.01, clamp_weights=True, threshold=1e-6, normalize_loss=True):
        return cls._factorizations[factorization_name](penalty=penalty, clamp_weights=clamp_weights, threshold=threshold,
                                                       normalize_loss=normalize_loss)

    # -- shared by the Tucker / TT hooks -----------------------------------------------------------------------------
    def _condition(self, gates):
        """Clamp to [-1, 1] and zero everything at or below the threshold, in place, outside autograd (:222-231, :296-305)."""
        with torch.no_grad():
            for g in gates:
                if self.clamp_weights:
                    g.data.clamp_(-1, 1)
                if self.threshold:
                    F.threshold(g.data, threshold=self.threshold, value=0, inplace=True)

    def _accumulate(self, gates):
        for g in gates:
            self.n_element += g.numel()
            self._loss = self._loss + torch.sum(torch.abs(g))

    def remove(self, module):
        delattr(module, "lasso_weights")

    def set_weights(self, module, value):
        with torch.no_grad():
            for g in module.lasso_weights:
                g.data.fill_(value)


class CPLasso(TensorLasso, factorization=CPTensor):
    """The CP weights' own gate vector: nothing is rescaled, the hook accumulates ``penalty * |gates|_1`` (:155-171)."""

    def __call__(self, module, input, cp_tensor):
        gates = module.lasso_weights
        self._condition([gates])
        self.n_element += gates.numel()
        self._loss = self._loss + self.penalty * torch.norm(gates, 1)
        return cp_tensor

    def apply(self, module):
        ref = module.factors[0]
        module.lasso_weights = nn.Parameter(torch.ones(module.rank, dtype=ref.dtype, device=ref.device))
        module.register_forward_hook(self)
        return module

    def set_weights(self, module, value):
        with torch.no_grad():
            module.lasso_weights.data.fill_(value)


class TuckerLasso(TensorLasso, factorization=TuckerTensor):
    """One gate vector per mode scaling that mode's factor columns (:220-246)."""

    def __call__(self, module, input, tucker_tensor):
        gates = list(module.lasso_weights)
        self._condition(gates)
        self._accumulate(gates)
        return self.apply_lasso(tucker_tensor, gates)

    def apply_lasso(self, tucker_tensor, lasso_weights):
        factors = [_scale_last(f, w) for f, w in zip(tucker_tensor.factors, lasso_weights)]
        return TuckerTensor(tucker_tensor.core, factors)

    def apply(self, module):
        core = module.core
        module.lasso_weights = ParameterList([nn.Parameter(torch.ones(r, dtype=core.dtype, device=core.device)) for r in module.rank])
        module.register_forward_hook(self)
        return module


class TTLasso(TensorLasso, factorization=TTTensor):
    """One gate per internal TT rank scaling the trailing rank of the core to its left (:294-320)."""

    def __call__(self, module, input, tt_tensor):
        gates = list(module.lasso_weights)
        self._condition(gates)
        self._accumulate(gates)
        return self.apply_lasso(tt_tensor, gates)

    def apply_lasso(self, tt_tensor, lasso_weights):
        cores = list(tt_tensor.factors)
        scaled = [_scale_last(f, w) for f, w in zip(cores, lasso_weights)] + [cores[-1]]
        return TTTensor(scaled)

    def apply(self, module):
        ref = module.factors[0]
        module.lasso_weights = ParameterList([nn.Parameter(torch.ones(1, 1, r, dtype=ref.dtype, device=ref.device))
                                              for r in module.rank[1:-1]])
        module.register_forward_hook(self)
        return module


def tensor_lasso(factorization="CP", penalty=0.01, clamp_weights=True, threshold=1e-6, normalize_loss=True):
    """Create a lasso regulariser for 'cp' / 'tucker' / 'tt' factorized tensors (:323-389); ``.apply(tensor)`` attaches it."""
    mapping = dict(cp="CPTensor", tucker="TuckerTensor", tt="TTTensor")
    return TensorLasso.from_factorization_name(mapping[factorization.lower()], penalty=penalty, clamp_weights=clamp_weights,
                                               threshold=threshold, normalize_loss=normalize_loss)


def remove_tensor_lasso(factorized_tensor):
    """Detach the lasso hook and its gate parameters from a factorized tensor (:391-413)."""
    for key, hook in factorized_tensor._forward_hooks.items():
        if isinstance(hook, TensorLasso):
            hook.remove(factorized_tensor)
            del factorized_tensor._forward_hooks[key]
            return factorized_tensor
    raise ValueError(f"TensorLasso not found in factorized tensor {factorized_tensor}")
