"""Tensor dropout: Bernoulli sub-sampling of the rank of a factorized tensor, applied as a forward hook on the weight
module (reference: tltorch/tensor_hooks/_tensor_dropout.py:12-196).

Mask GENERATION replays the reference's torch RNG calls in the same order (one ``torch.bernoulli`` per rank mode, a
``torch.randint`` fallback when nothing survives), so a seeded run drops the same components.  Mask APPLICATION -- the
gathers of factor columns / core sub-blocks and the 1/(1-p) rescaling -- is one ``torch.index_select`` chain on tiny
tensors followed by a contraction-kernel scale, and produces a fresh (smaller) factorized tensor that the contraction
kernels then consume unchanged.
"""
import torch

from ..factorized_tensors import TuckerTensor, CPTensor, TTTensor
from ..ops import contract


def _sample(rank, proba, min_dim, min_values, device):
    """Keep-index vector for one rank mode (reference sampling loop: :64-73, :93-98, :116-127)."""
    idx = torch.arange(rank, device=device, dtype=torch.int64)
    if rank > min_dim:
        keep = torch.bernoulli(torch.ones(rank, device=device) * (1 - proba),
                               out=torch.empty(rank, device=device, dtype=torch.bool))
        idx = idx[keep]
        if len(idx) == 0:
            idx = torch.randint(0, rank, size=(min_values,), device=device, dtype=torch.int64)
    return idx


def _scale(t, factor):
    """t * factor through the contraction kernel (keeps autograd on the custom-op path)."""
    if factor == 1:
        return t
    letters = "abcdefghijklmnopqrstuvwxyz"[:t.dim()]
    one = torch.ones((), dtype=t.dtype, device=t.device)
    return contract(f"{letters},->{letters}", t, one, alpha=float(factor))


class TensorDropout:
    """Forward hook applying tensor dropout to a FactorizedTensor.

    proba : probability of dropping a rank component; min_dim : modes with rank <= min_dim are left untouched;
    min_values : components kept when the Bernoulli draw is empty; drop_test : also drop in eval mode."""
    _factorizations = dict()

    def __init_subclass__(cls, factorization, **kwargs):
        cls._factorizations[factorization.__name__] = cls

    def __init__(self, proba, min_dim=1, min_values=1, drop_test=False):
        assert 0 <= proba < 1, f"Got prob={proba} but tensor dropout is defined for 0 <= proba < 1."
        self.proba = proba
        self.min_dim = min_dim
        self.min_values = min_values
        self.drop_test = drop_test

    def __call__(self, module, input, factorized_tensor):
        return self._apply_tensor_dropout(factorized_tensor, training=module.training)

    def _inactive(self, training):
        return (not self.proba) or ((not training) and (not self.drop_test))

    def _apply_tensor_dropout(self, factorized_tensor, training=True):
        raise NotImplementedError()

    @classmethod
    def apply(cls, module, proba, min_dim=3, min_values=1, drop_test=False):
        # keyed by class NAME like the reference (:26-28, :45): the *Tensorized / BlockTT classes have no entry, so
        # tensor dropout on a FactorizedLinear weight raises KeyError there too
        cls = cls._factorizations[module.__class__.__name__]
        for hook in module._forward_hooks.values():
            if isinstance(hook, cls):
                raise RuntimeError("Cannot register two weight_norm hooks on the same parameter")
        return module.register_forward_hook(cls(proba, min_dim=min_dim, min_values=min_values, drop_test=drop_test))


class TuckerDropout(TensorDropout, factorization=TuckerTensor):
    def _apply_tensor_dropout(self, tucker_tensor, training=True):
        if self._inactive(training):
            return tucker_tensor
        core, factors = tucker_tensor.core, list(tucker_tensor.factors)
        kept = [_sample(r, self.proba, self.min_dim, self.min_values, core.device) for r in tucker_tensor.rank]
        sub = core
        for mode, idx in enumerate(kept):
            sub = sub.index_select(mode, idx)
        if training:   # the reference rescales by 1/(1-p)^core.ndim even for modes it did not drop (App. C.4)
            sub = _scale(sub, 1 / ((1 - self.proba) ** core.dim()))
        return TuckerTensor(sub, [f.index_select(1, idx) for f, idx in zip(factors, kept)])


class CPDropout(TensorDropout, factorization=CPTensor):
    def _apply_tensor_dropout(self, cp_tensor, training=True):
        if self._inactive(training):
            return cp_tensor
        rank = cp_tensor.rank
        if rank <= self.min_dim:
            # the reference leaves `weights` / `factors` unbound here and crashes with UnboundLocalError (App. C.3);
            # returning the tensor untouched is the evident intent
            return cp_tensor
        idx = _sample(rank, self.proba, self.min_dim, self.min_values, cp_tensor.factors[0].device)
        weights = cp_tensor.weights.index_select(0, idx)
        if training:
            weights = _scale(weights, 1 / (1 - self.proba))
        return CPTensor(weights, [f.index_select(1, idx) for f in cp_tensor.factors])


class TTDropout(TensorDropout, factorization=TTTensor):
    def _apply_tensor_dropout(self, tt_tensor, training=True):
        if self._inactive(training):
            return tt_tensor
        device = tt_tensor.factors[0].device
        kept = [_sample(r, self.proba, self.min_dim, self.min_values, device) for r in tt_tensor.rank[1:]]
        scaling = 1 / (1 - self.proba) if training else 1
        out = []
        last = tt_tensor.order - 1
        for i, f in enumerate(tt_tensor.factors):
            if i > 0:
                f = f.index_select(0, kept[i - 1])
            if i < last:
                f = _scale(f.index_select(2, kept[i]), scaling)
            out.append(f)
        return TTTensor(out)


def tensor_dropout(factorized_tensor, p=0, min_dim=3, min_values=1, drop_test=False):
    """Attach tensor dropout (probability ``p``) to a factorized tensor module and return the module.

    >>> tensor = FactorizedTensor.new((3, 4, 2), rank=0.5, factorization='CP').normal_()
    >>> tensor = tensor_dropout(tensor, p=0.5)
    >>> remove_tensor_dropout(tensor)
    """
    TensorDropout.apply(factorized_tensor, p, min_dim=min_dim, min_values=min_values, drop_test=drop_test)
    return factorized_tensor


def remove_tensor_dropout(factorized_tensor):
    for key, hook in factorized_tensor._forward_hooks.items():
        if isinstance(hook, TensorDropout):
            del factorized_tensor._forward_hooks[key]
            return factorized_tensor
    raise ValueError(f"TensorLasso not found in factorized tensor {factorized_tensor}")
