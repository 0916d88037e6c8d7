"""Factorized embedding lookup: rows of a tensorized, factorized (N, D) table selected by an index vector WITHOUT forming
the table (reference: FactorizedEmbedding.forward, tltorch/factorized_layers/factorized_embedding.py:96-117, which indexes
the weight in factorized form -- tensorized_matrices.py:136-193 (CP), :230-311 (Tucker), :386-489 (BlockTT)).

A row index i of the tensorized row shape (n_1..n_m) unravels to (i_1..i_m); the gathered row of every row-mode factor is
then contracted with the column-mode factors by the library's own contraction kernels (``ops.contract``), batched over the
looked-up positions:

  CP       rows[p, :] = (lambda * prod_k U_k[i_k(p), :]) . KhatriRao(V_l)^T
  Tucker   rows[p, :] = core x_k U_k[i_k(p), :] x_l V_l
  BlockTT  rows[p, :] = G_1[:, i_1(p), :, :] G_2[:, i_2(p), :, :] ...      (TT-Rec style chain)

The gather of factor rows / core slices itself is a torch ``index_select`` (its backward is the matching scatter-add).
"""
import math

import numpy as np
import torch

from .ops import contract
from . import tenalg

_L = "abcdefghijklmopqrstuvwxyz"          # 'n' is reserved for the looked-up position


def unravel(flat_index, shape):
    """torch version of np.unravel_index (row-major) that stays on the index tensor's device."""
    out = []
    rem = flat_index
    for s in reversed(shape):
        out.append(rem % s)
        rem = torch.div(rem, s, rounding_mode="floor")
    return list(reversed(out))


def cp_row_factor(weights, row_factors, flat_index, row_shape):
    """(P, R) matrix  prod_k U_k[i_k(p), r]  (lambda is NOT folded in)."""
    idx = unravel(flat_index, row_shape)
    out = None
    for f, i in zip(row_factors, idx):
        g = f.index_select(0, i)
        out = g if out is None else contract("nr,nr->nr", out, g)
    return out


def cp_rows(weights, row_factors, col_factors, flat_index, row_shape):
    lead = cp_row_factor(weights, row_factors, flat_index, row_shape)
    full = tenalg.cp_to_tensor(weights, [lead] + list(col_factors))
    return full.reshape(full.shape[0], -1)


def tucker_rows(core, row_factors, col_factors, flat_index, row_shape):
    idx = unravel(flat_index, row_shape)
    res = core
    letters = _L[:core.dim()]
    for pos, (f, i) in enumerate(zip(row_factors, idx)):
        g = f.index_select(0, i)                                       # (P, a_pos)
        rest = letters[pos + 1:]
        if pos == 0:
            res = contract(f"{letters[0]}{rest},n{letters[0]}->n{rest}", res, g)
        else:
            res = contract(f"n{letters[pos]}{rest},n{letters[pos]}->n{rest}", res, g)
    for pos, f in enumerate(col_factors):                              # expand the column modes: res x_{1 + pos} V_pos
        res = tenalg.mode_dot(res, f, 1 + pos)
    return res.reshape(res.shape[0], -1)


def blocktt_rows(cores, flat_index, row_shape):
    """cores[k] (r_k, n_k, d_k, r_{k+1}) -> (P, prod d)."""
    idx = unravel(flat_index, row_shape)
    res = None
    for core, i in zip(cores, idx):
        g = core.index_select(1, i)                                    # (r_k, P, d_k, r_{k+1})
        if res is None:
            res = g[0]                                                 # r_0 = 1 -> (P, d_0, r_1)
        else:
            nxt = contract("nxr,rnys->nxys", res, g)
            res = nxt.reshape(nxt.shape[0], nxt.shape[1] * nxt.shape[2], nxt.shape[3])
    return res[:, :, 0]


def _as_index(index, device):
    if torch.is_tensor(index):
        return index.reshape(-1).to(device=device, dtype=torch.long)
    return torch.as_tensor(np.asarray(index).reshape(-1), dtype=torch.long, device=device)


def parse_row_gather(indices, tensorized_shape):
    """Recognise ``weight[(int, ...)*, rows, :]``: integer indices on leading plain modes, an index vector on the tensorized ROW mode
    and a full slice on the tensorized COLUMN mode.  Returns (lead ints, rows) or None."""
    if not isinstance(indices, tuple):
        return None
    n_lead = 0
    for s in tensorized_shape:
        if isinstance(s, (int, np.integer)):
            n_lead += 1
        else:
            break
    if len(tensorized_shape) != n_lead + 2 or len(indices) != n_lead + 2:
        return None
    lead, rows, cols = indices[:n_lead], indices[n_lead], indices[n_lead + 1]
    if not all(isinstance(i, (int, np.integer)) for i in lead):
        return None
    if not (isinstance(cols, slice) and cols == slice(None)):
        return None
    if isinstance(rows, (slice, int, np.integer)) or rows is Ellipsis:
        return None
    return tuple(int(i) for i in lead), rows
