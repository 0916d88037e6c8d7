"""Import shim named ``tensorly`` -- TEST INFRASTRUCTURE ONLY.

Adapter over ``oracle.tensorly_semantics`` exposing just the names the reference package imports, so
that ``tests/golden/make_golden.py`` can import and execute the REAL reference (``/root/reference/tltorch``)
in the build container, where tensorly itself is not installable.  Never put this directory on
``sys.path`` outside that script / the oracle-vs-reference test.
"""
import sys as _sys
import os as _os

import numpy as _np
import torch as _torch

_here = _os.path.dirname(_os.path.abspath(__file__))
_repo = _os.path.abspath(_os.path.join(_here, "..", "..", ".."))
if _repo not in _sys.path:
    _sys.path.insert(0, _repo)

from oracle import tensorly_semantics as _sem  # noqa: E402

__version__ = "shim-of-oracle.tensorly_semantics"

float32 = _torch.float32
float64 = _torch.float64
int64 = _torch.int64


def set_backend(name, local_threadsafe=False):
    if name != "pytorch":
        raise ValueError("the shim only restates the pytorch backend")


def get_backend():
    return "pytorch"


def einsum(eq, *operands):
    return _torch.einsum(eq, *operands)


shape = _sem.shape
ndim = _sem.ndim
transpose = _sem.transpose
moveaxis = _sem.moveaxis


def reshape(t, shp):
    return _torch.reshape(t, tuple(shp))


def copy(t):
    return t.clone()


def sum(t, axis=None, keepdims=False):
    if axis is None:
        return _torch.sum(t)
    return _torch.sum(t, dim=axis, keepdim=keepdims)


def arange(*args, **kwargs):
    return _torch.arange(*args, **kwargs)


def context(t):
    return {"dtype": t.dtype, "device": t.device, "requires_grad": t.requires_grad}


def tensor(data, dtype=None, device=None, requires_grad=False, **kw):
    if isinstance(data, _np.ndarray):
        return _torch.tensor(data.copy(), dtype=dtype, device=device, requires_grad=requires_grad)
    return _torch.tensor(data, dtype=dtype, device=device, requires_grad=requires_grad)


def to_numpy(t):
    if _torch.is_tensor(t):
        return t.detach().cpu().numpy()
    return _np.asarray(t)


def check_random_state(seed):
    if seed is None:
        return _np.random.mtrand._rand
    if isinstance(seed, int):
        return _np.random.RandomState(seed)
    if isinstance(seed, _np.random.RandomState):
        return seed
    raise ValueError("Seed should be None, int or np.random.RandomState")


def randn(shp, seed=None, **ctx):
    rng = check_random_state(seed)
    return tensor(rng.randn(*shp), **ctx)


def norm(t, order=2, axis=None):
    if axis is None:
        return _torch.linalg.vector_norm(t.reshape(-1), ord=order)
    return _torch.linalg.vector_norm(t, ord=order, dim=axis)


def abs(t):
    return _torch.abs(t)


def max(t):
    return _torch.max(t)


cp_to_tensor = _sem.cp_to_tensor
tucker_to_tensor = _sem.tucker_to_tensor
tt_to_tensor = _sem.tt_to_tensor

from . import tenalg, cp_tensor, tucker_tensor, tt_tensor, tt_matrix, decomposition, testing, random  # noqa: E402,F401
