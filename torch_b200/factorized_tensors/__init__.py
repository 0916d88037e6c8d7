from .core import FactorizedTensor, TensorizedTensor  # noqa: F401
from .factorized_tensors import CPTensor, TuckerTensor, TTTensor, DenseTensor  # noqa: F401
from .tensorized_matrices import (CPTensorized, TuckerTensorized, BlockTT, DenseTensorized, TTMatrix,  # noqa: F401
                                  validate_block_tt_rank)
from .init import tensor_init, cp_init, tucker_init, tt_init, block_tt_init  # noqa: F401
