"""Std-preserving random initialisation helpers (reference: tltorch/factorized_tensors/init.py)."""
import math

import numpy as np
import torch


def tensor_init(tensor, std=0.02):
    """Initialise a factorized tensor so that its reconstruction has (approximately) standard deviation ``std``."""
    from .factorized_tensors import CPTensor, TuckerTensor, TTTensor
    from .tensorized_matrices import BlockTT
    if isinstance(tensor, CPTensor):
        return cp_init(tensor, std=std)
    if isinstance(tensor, TuckerTensor):
        return tucker_init(tensor, std=std)
    if isinstance(tensor, TTTensor):
        return tt_init(tensor, std=std)
    if isinstance(tensor, BlockTT):
        return block_tt_init(tensor, std=std)
    if torch.is_tensor(tensor):
        return tensor.normal_(0, std)
    raise ValueError(f"Got tensor of class {tensor.__class__.__name__} but expected a factorized tensor")


def cp_init(cp_tensor, std=0.02):
    rank = cp_tensor.rank
    std_factors = (std / math.sqrt(rank)) ** (1 / cp_tensor.order)
    with torch.no_grad():
        cp_tensor.weights.fill_(1)
        for factor in cp_tensor.factors:
            factor.normal_(0, std_factors)
    return cp_tensor


def tucker_init(tucker_tensor, std=0.02):
    r = float(np.prod([math.sqrt(r) for r in tucker_tensor.rank]))
    std_factors = (std / r) ** (1 / (tucker_tensor.order + 1))
    with torch.no_grad():
        tucker_tensor.core.normal_(0, std_factors)
        for factor in tucker_tensor.factors:
            factor.normal_(0, std_factors)
    return tucker_tensor


def tt_init(tt_tensor, std=0.02):
    r = float(np.prod(tt_tensor.rank))
    std_factors = (std / r) ** (1 / tt_tensor.order)
    with torch.no_grad():
        for factor in tt_tensor.factors:
            factor.normal_(0, std_factors)
    return tt_tensor


def block_tt_init(block_tt, std=0.02):
    return tt_init(block_tt, std=std)
