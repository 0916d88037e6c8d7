"""Tensor regression layer contractions: B200 counterparts of tltorch/functional/tensor_regression.py."""
import math

from .. import tenalg
from .. import trl_ops


def trl(x, weight, bias=None, **kwargs):
    """Tensor Regression Layer: generalised inner product of x with the (factorized) regression weights.
    reference: functional/tensor_regression.py:11-34."""
    from ..factorized_tensors import TuckerTensor
    if isinstance(weight, TuckerTensor):
        return tucker_trl(x, weight, bias=bias, **kwargs)
    return tenalg.inner(x, weight.to_tensor(), n_modes=x.dim() - 1, bias=bias)


def _tucker_trl_cost(x_shape, weight):
    """MACs of (reconstruct + inner) vs (project input + expand core + inner)."""
    n_in = len(x_shape) - 1
    batch = x_shape[0]
    shape, rank = list(weight.shape), list(weight.rank)
    full = math.prod(shape)
    rec = 0
    cur = list(rank)
    for k in range(len(shape)):                   # core x_k F_k, mode by mode
        cur[k] = shape[k]
        rec += math.prod(cur) * rank[k]
    cost_rec = rec + batch * full
    proj = 0
    cur = list(x_shape[1:])
    for k in range(n_in):
        proj += batch * math.prod(cur) * rank[k]
        cur[k] = rank[k]
    exp = 0
    cur_w = list(rank)
    for k in range(n_in, len(shape)):
        cur_w[k] = shape[k]
        exp += math.prod(cur_w) * rank[k]
    cost_proj = proj + exp + batch * math.prod(rank[:n_in]) * math.prod(shape[n_in:])
    return cost_rec, cost_proj


def tucker_trl(x, weight, project_input=False, bias=None):
    """reference: functional/tensor_regression.py:37-49.

    ``project_input=True`` contracts x with the input factors first.  With the default (False) the reference rebuilds
    the full regression tensor every step; the result is identical either way, so when the projection order is
    cheaper (it is >10x cheaper at SURVEY config 4) it is used regardless."""
    n_input = x.dim() - 1
    if not project_input:
        cost_rec, cost_proj = _tucker_trl_cost(tuple(x.shape), weight)
        project_input = cost_proj < cost_rec
    factors = list(weight.factors)
    if project_input and trl_ops.eligible(x, weight, bias):
        # projection (one pass over the activation + the channel GEMM) -> ONE launch for core contraction, output expansion and bias
        return trl_ops.tucker_trl_fused(x, weight, bias)
    if project_input:
        x = tenalg.multi_mode_dot(x, factors[:n_input], modes=range(1, n_input + 1), transpose=True)
        regression_weights = tenalg.multi_mode_dot(weight.core, factors[n_input:], modes=range(n_input, weight.order))
    else:
        regression_weights = weight.to_tensor()
    return tenalg.inner(x, regression_weights, n_modes=x.dim() - 1, bias=bias)
