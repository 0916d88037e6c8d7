"""Base classes and registries of the factorized parameter containers.

API-compatible with tltorch/factorized_tensors/core.py (constructor arguments, ``new`` / ``from_tensor`` /
``from_matrix`` class methods, ``forward(indices)`` so that forward hooks fire, ``to_tensor`` / ``to_matrix``,
``normal_``), re-authored.  Containers are ``nn.Module``s that own their factors as ``nn.Parameter``s under the same
names as the reference (``weights``, ``core``, ``factors.factor_{i}``), so reference state_dicts load unchanged.
"""
import numpy as np
import torch
from torch import nn


def _ensure_tuple(value):
    if isinstance(value, int):
        return () if value == 1 else (value,)
    value = tuple(value)
    return () if value == (1,) else value


def _canonical(name):
    return "dense" if name is None else name.lower()


def flatten_tensorized_shape(tensorized_shape):
    out = ()
    for e in tensorized_shape:
        out += (int(e),) if isinstance(e, (int, np.integer)) else tuple(int(s) for s in e)
    return out


class _FactorizationMeta(type(nn.Module)):
    """``FactorizedTensor(..., factorization='cp')`` instantiates the registered subclass; calling a concrete class
    directly ignores a stray ``factorization`` keyword."""

    def __call__(cls, *args, **kwargs):
        factorization = kwargs.pop("factorization", None)
        if cls.__dict__.get("_is_registry_root", False):
            cls = cls._lookup(factorization)
        return super(_FactorizationMeta, cls).__call__(*args, **kwargs)


class FactorizedTensor(nn.Module, metaclass=_FactorizationMeta):
    """Tensor in factorized form.  Every factorization exposes ``shape``, ``rank`` and ``order``."""
    _is_registry_root = True
    _factorizations = {}

    def __init_subclass__(cls, name=None, **kwargs):
        super().__init_subclass__(**kwargs)
        if name:
            cls._registry()[_canonical(name)] = cls
            cls._name = name

    @classmethod
    def _registry(cls):
        return TensorizedTensor._factorizations if _is_tensorized(cls) else FactorizedTensor._factorizations

    @classmethod
    def _lookup(cls, factorization):
        try:
            return cls._registry()[_canonical(factorization)]
        except KeyError:
            raise ValueError(f"Got factorization={factorization} but expected"
                             f"one of {cls._registry().keys()}")

    # ---- creation ----------------------------------------------------------------------------------------------
    @classmethod
    def new(cls, shape, rank="same", factorization="Tucker", **kwargs):
        """Main way to create a factorized tensor: ``FactorizedTensor.new((3, 4, 2), rank=0.5, factorization='tucker')``."""
        return cls._lookup(factorization).new(shape, rank, **kwargs)

    @classmethod
    def from_tensor(cls, tensor, rank, factorization="CP", **kwargs):
        return cls._lookup(factorization).from_tensor(tensor, rank, **kwargs)

    # ---- use inside a network ----------------------------------------------------------------------------------
    def forward(self, indices=None, **kwargs):
        """``tensor()`` returns the factorization itself (after forward hooks, e.g. tensor dropout);
        ``tensor(indices)`` returns ``tensor[indices]``."""
        return self if indices is None else self[indices]

    def __getitem__(self, indices):
        raise NotImplementedError

    @property
    def decomposition(self):
        raise NotImplementedError

    def to_tensor(self):
        raise NotImplementedError

    def normal_(self, mean=0, std=1):
        if mean != 0:
            raise ValueError(f"Currently only mean=0 is supported, but got mean={mean}")
        return self

    # ---- tensor-like introspection -----------------------------------------------------------------------------
    def dim(self):
        return len(self.shape)

    @property
    def ndim(self):
        return len(self.shape)

    def numel(self):
        return int(np.prod(self.shape))

    def size(self, index=None):
        return self.shape if index is None else self.shape[index]

    @property
    def name(self):
        return self._name

    @property
    def tensor_shape(self):
        return self.shape

    def __repr__(self):
        return f"{self.__class__.__name__}(shape={self.shape}, rank={self.rank})"

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        """Passing a factorized tensor to a torch function reconstructs it first (reference: core.py:376-383)."""
        kwargs = kwargs or {}
        args = [a.to_tensor() if hasattr(a, "to_tensor") else a for a in args]
        return func(*args, **kwargs)


class TensorizedTensor(FactorizedTensor):
    """Matrix (or batch of matrices) whose rows / columns are tensorized.  ``order`` and ``tensorized_shape`` describe the
    underlying tensor; ``shape``, ``dim`` and ``ndim`` describe the matrix."""
    _is_registry_root = True
    _factorizations = {}

    @classmethod
    def new(cls, tensorized_shape, rank, factorization="CP", **kwargs):
        return cls._lookup(factorization).new(tensorized_shape, rank, **kwargs)

    @classmethod
    def from_tensor(cls, tensor, shape, rank, factorization="CP", **kwargs):
        return cls._lookup(factorization).from_tensor(tensor, shape, rank, **kwargs)

    @classmethod
    def from_matrix(cls, matrix, tensorized_row_shape, tensorized_column_shape, rank, factorization="CP", **kwargs):
        batch_dims = _ensure_tuple(tuple(matrix.shape[:-2])) if matrix.ndim > 2 else ()
        tensor = matrix.reshape((*batch_dims, *tensorized_row_shape, *tensorized_column_shape))
        return cls.from_tensor(tensor, batch_dims + (tensorized_row_shape, tensorized_column_shape), rank,
                               factorization=factorization, **kwargs)

    def to_matrix(self):
        """Dense (batch of) matrix: ``to_tensor().reshape(shape)`` (reference: core.py:536-543)."""
        return self.to_tensor().reshape(tuple(int(s) for s in self.shape))

    @property
    def tensor_shape(self):
        return flatten_tensorized_shape(self.tensorized_shape)

    def init_from_matrix(self, matrix, **kwargs):
        return self.init_from_tensor(matrix.reshape(self.tensor_shape), **kwargs)

    def __repr__(self):
        return (f"{self.__class__.__name__}(shape={self.shape}, tensorized_shape={self.tensorized_shape}, "
                f"rank={self.rank})")

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        args = [a.to_matrix() if hasattr(a, "to_matrix") else a for a in args]
        return func(*args, **kwargs)


def _is_tensorized(cls):
    try:
        return issubclass(cls, TensorizedTensor)
    except NameError:          # while TensorizedTensor itself is being defined
        return False
