"""``import tltorch`` drop-in: resolves to the B200-native implementation (torch_b200) of the TensorLy-Torch factorized-layer
hot path, so code written against the reference's package name runs unchanged:

    import tltorch
    layer = tltorch.FactorizedLinear((8, 8, 12), (8, 12, 32), factorization='blocktt', rank=16, device='cuda')
    from tltorch.functional import factorized_linear
    from tltorch.factorized_tensors import CPTensor

Every submodule of torch_b200 is registered under the matching ``tltorch.*`` name (same module objects, no copies).
"""
import importlib
import pkgutil
import sys

import torch_b200 as _impl

_self = sys.modules[__name__]
for _k, _v in vars(_impl).items():
    if not _k.startswith("__") or _k in ("__version__", "__doc__"):
        setattr(_self, _k, _v)
__path__ = list(_impl.__path__)          # ``import tltorch.x.y`` finds torch_b200's sources ...


def _alias(prefix_src, prefix_dst):
    for name, mod in list(sys.modules.items()):
        if name == prefix_src or name.startswith(prefix_src + "."):
            sys.modules.setdefault(prefix_dst + name[len(prefix_src):], mod)


for _m in pkgutil.walk_packages(_impl.__path__, "torch_b200."):
    if _m.name.startswith("torch_b200.csrc"):
        continue                         # the build script is not part of the API
    try:
        importlib.import_module(_m.name)
    except Exception:                    # optional pieces must not break ``import tltorch``
        pass
_alias("torch_b200", "tltorch")          # ... and already-imported ones are the SAME module objects
