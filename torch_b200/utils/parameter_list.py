"""Parameter containers.  state_dict keys must stay ``factor_{i}`` / ``param_{i}`` so that checkpoints written by the
reference load unchanged (reference: tltorch/utils/parameter_list.py:5-82, :110-173)."""
import torch
from torch import nn


class _KeyedList(nn.Module):
    _prefix = "item"

    def __init__(self, parameters=None):
        super().__init__()
        self.keys = []
        self.counter = 0
        if parameters is not None:
            self.extend(parameters)

    def _new_key(self):
        key = f"{self._prefix}_{self.counter}"
        self.counter += 1
        return key

    def _store(self, key, element):
        raise NotImplementedError

    def append(self, element):
        key = self._new_key()
        self._store(key, element)
        self.keys.append(key)

    def insert(self, index, element):
        key = self._new_key()
        self._store(key, element)
        self.keys.insert(index, key)

    def extend(self, parameters):
        for p in parameters:
            self.append(p)
        return self

    def pop(self, index=-1):
        item = self[index]
        del self[index]
        return item

    def __getitem__(self, index):
        picked = self.keys[index]
        if isinstance(picked, list):
            return self.__class__([getattr(self, k) for k in picked])
        return getattr(self, picked)

    def __setitem__(self, index, value):
        self._store(self.keys[index], value)

    def __delitem__(self, index):
        delattr(self, self.keys[index])
        del self.keys[index]

    def __len__(self):
        return len(self.keys)

    def __iter__(self):
        return (getattr(self, k) for k in self.keys)

    def __iadd__(self, parameters):
        return self.extend(parameters)

    def __add__(self, parameters):
        out = self.__class__(list(self))
        out.extend(parameters)
        return out

    def __radd__(self, parameters):
        out = self.__class__(parameters)
        out.extend(list(self))
        return out

    def extra_repr(self):
        lines = []
        for k, p in self._parameters.items():
            where = f" (GPU {p.get_device()})" if p.is_cuda else ""
            lines.append(f"  ({k}): Parameter containing: [{torch.typename(p)} of size "
                         f"{'x'.join(str(s) for s in p.size())}{where}]")
        return "\n".join(lines)


class FactorList(_KeyedList):
    """List of factors: nn.Parameters are registered as parameters, plain tensors as buffers, lists recursively."""
    _prefix = "factor"

    def _store(self, key, element):
        if key in self._buffers:
            del self._buffers[key]
        if key in self._parameters:
            del self._parameters[key]
        if torch.is_tensor(element):
            if isinstance(element, nn.Parameter):
                self.register_parameter(key, element)
            else:
                self.register_buffer(key, element)
        else:
            setattr(self, key, self.__class__(element))


class ParameterList(_KeyedList):
    _prefix = "param"

    def _store(self, key, element):
        self.register_parameter(key, element)


class ComplexFactorList(FactorList):
    """FactorList of complex factors stored as real (..., 2) parameters -- optimisers and DDP see real tensors -- and handed
    out as complex views (reference: tltorch/utils/parameter_list.py:85-107)."""

    def __getitem__(self, index):
        if isinstance(index, int):
            value = getattr(self, self.keys[index])
            return torch.view_as_complex(value) if torch.is_tensor(value) else value
        return self.__class__([torch.view_as_complex(getattr(self, k)) for k in self.keys[index]])

    def __iter__(self):
        return (self[i] for i in range(len(self.keys)))

    def _store(self, key, element):
        if torch.is_tensor(element) and element.is_complex():
            real = torch.view_as_real(element.detach() if isinstance(element, nn.Parameter) else element)
            element = nn.Parameter(real, requires_grad=element.requires_grad) if isinstance(element, nn.Parameter) else real
        super()._store(key, element)
