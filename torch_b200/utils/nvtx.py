"""Optional NVTX ranges around the layers' forward passes: ``TLT_NVTX=1`` names every factorized layer call on the
timeline of Nsight Systems / ``ncu --nvtx`` (`tltorch.FactorizedConv[cp] 64->64`), off by default (a range push / pop per call)."""
import functools
import os

import torch

ENABLED = os.environ.get("TLT_NVTX", "0") == "1"


def annotate(label_of):
    """Decorator for a module's ``forward``: wraps the call in an NVTX range named by ``label_of(self)`` when enabled."""
    def wrap(fn):
        if not ENABLED:
            return fn

        @functools.wraps(fn)
        def inner(self, *args, **kwargs):
            if not torch.cuda.is_available():
                return fn(self, *args, **kwargs)
            torch.cuda.nvtx.range_push(label_of(self))
            try:
                return fn(self, *args, **kwargs)
            finally:
                torch.cuda.nvtx.range_pop()
        return inner
    return wrap
