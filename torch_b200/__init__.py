"""torch_b200: B200-native (sm_100a) implementation of the TensorLy-Torch factorized-layer hot path.

Same public API as ``tltorch`` for that path: FactorizedLinear, FactorizedConv, TRL, TCL, the CP / Tucker / TT /
BlockTT tensor classes, ``functional`` and the tensor-dropout hook.  All contractions run in hand-written CUDA kernels
(torch_b200/csrc) reached through a C ABI (include/tltorch_cuda.h); there is no CPU fallback.
"""
__version__ = "0.5.0+b200.r1"

from . import utils  # noqa: F401
from . import ops  # noqa: F401
from . import tenalg  # noqa: F401
from . import factorized_tensors  # noqa: F401
from .factorized_tensors import init  # noqa: F401
from . import functional  # noqa: F401
from . import factorized_layers  # noqa: F401
from . import tensor_hooks  # noqa: F401

from .factorized_layers import FactorizedLinear, FactorizedConv, TRL, TCL, FactorizedEmbedding  # noqa: F401
from .factorized_tensors import (FactorizedTensor, DenseTensor, CPTensor, TTTensor, TuckerTensor, tensor_init,  # noqa: F401
                                 TensorizedTensor, CPTensorized, BlockTT, TTMatrix, DenseTensorized, TuckerTensorized)
from .tensor_hooks import tensor_dropout, remove_tensor_dropout, tensor_lasso, remove_tensor_lasso  # noqa: F401
from .factorized_tensors import (ComplexCPTensor, ComplexTuckerTensor, ComplexTTTensor, ComplexDenseTensor,  # noqa: F401
                                 ComplexCPTensorized, ComplexBlockTT, ComplexDenseTensorized, ComplexTuckerTensorized)
from . import decomposition  # noqa: F401
from .utils import FactorList, ParameterList, ComplexFactorList, get_tensorized_shape  # noqa: F401
