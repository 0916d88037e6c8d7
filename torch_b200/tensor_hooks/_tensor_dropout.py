"""Tensor dropout: Bernoulli sub-sampling of the rank of a factorized tensor, applied as a forward hook on the weight
module (reference: tltorch/tensor_hooks/_tensor_dropout.py:12-196).

Mask GENERATION replays the reference's torch RNG calls in the same order (one ``torch.bernoulli`` per rank mode, a
``torch.randint`` fallback when nothing survives), so a seeded run drops the same components.

Mask APPLICATION has two forms that give the same tensor:
  * fixed shape (default whenever an empty draw is impossible in practice, p ** rank < 1e-30): the 0/1 keep vectors and
    the 1/(1-p) factor are MULTIPLIED into the core / weights / cores by the contraction kernel (dropped components
    contribute exact zeros), the factors are used as they are.  No index vector ever reaches the host, shapes never
    change, so the whole training step -- mask draw included -- is CUDA-graph capturable;
  * gather (small ranks, where the reference's ``len(idx) == 0`` fallback can actually fire and has to be checked on the
    host): ``index_select`` of factor columns / core sub-blocks like the reference.
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


def _empty_draw_impossible(rank, proba):
    return proba == 0 or rank * (-__import__("math").log10(proba)) > 30


def _keep_mask(rank, proba, dtype, device):
    """0/1 keep vector drawn with the SAME RNG call as ``_sample`` (the reference's torch.bernoulli, :64-73)."""
    keep = torch.bernoulli(torch.ones(rank, device=device) * (1 - proba),
                           out=torch.empty(rank, device=device, dtype=torch.bool))
    return keep.to(dtype)


def _mask_mode(t, mode, mask, alpha=1.0):
    """t * mask along ``mode`` (* alpha) through the contraction kernel."""
    letters = "abcdefghijklmnopqrstuvwxyz"[:t.dim()]
    return contract(f"{letters},{letters[mode]}->{letters}", t, mask, alpha=float(alpha))


def _apply_masks(t, masks, alpha):
    """t * alpha * prod_k masks[k][i_k] in ONE launch (tlt_scale_modes); wider / exotic tensors go mode by mode."""
    from ..rows_ops import scale_modes
    if t.dim() <= 6 and t.dtype in (torch.float32, torch.bfloat16):
        return scale_modes(t, masks, alpha)
    out, first = t, True
    for mode, m in enumerate(masks):
        if m is not None:
            out = _mask_mode(out, mode, m, alpha if first else 1.0)
            first = False
    return _scale(out, alpha) if first else out


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
        dropped = [r > self.min_dim for r in tucker_tensor.rank]
        if all(_empty_draw_impossible(r, self.proba) for r, d in zip(tucker_tensor.rank, dropped) if d):
            # fixed-shape form: core * (m_0 x m_1 x ...) / (1-p)^ndim, factors untouched
            scale = 1 / ((1 - self.proba) ** core.dim()) if training else 1.0
            masks = [_keep_mask(r, self.proba, core.dtype, core.device) if d else None for r, d in zip(tucker_tensor.rank, dropped)]
            return TuckerTensor(_apply_masks(core, masks, scale), factors)
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
        if _empty_draw_impossible(rank, self.proba):
            w = cp_tensor.weights
            mask = _keep_mask(rank, self.proba, w.dtype, w.device)
            return CPTensor(_apply_masks(w, [mask], 1 / (1 - self.proba) if training else 1.0), list(cp_tensor.factors))
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
        inner = list(tt_tensor.rank[1:-1])
        if all(r > self.min_dim and _empty_draw_impossible(r, self.proba) for r in inner) and tt_tensor.rank[-1] <= self.min_dim:
            # fixed-shape form: every core but the last is masked along its right rank (which zeroes the matching left
            # slice of the next core's contribution) and rescaled by 1/(1-p)
            scaling = 1 / (1 - self.proba) if training else 1.0
            out = []
            for i, f in enumerate(tt_tensor.factors):
                if i < tt_tensor.order - 1:
                    f = _apply_masks(f, [None, None, _keep_mask(f.shape[2], self.proba, f.dtype, device)], scaling)
                out.append(f)
            return TTTensor(out)
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
