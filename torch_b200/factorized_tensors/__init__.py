from .core import FactorizedTensor, TensorizedTensor  # noqa: F401
from .factorized_tensors import CPTensor, TuckerTensor, TTTensor, DenseTensor  # noqa: F401
from .tensorized_matrices import (CPTensorized, TuckerTensorized, BlockTT, DenseTensorized, TTMatrix,  # noqa: F401
                                  validate_block_tt_rank)
from .complex_factorized_tensors import ComplexCPTensor, ComplexTuckerTensor, ComplexTTTensor, ComplexDenseTensor  # noqa: F401
from .complex_tensorized_matrices import (ComplexCPTensorized, ComplexBlockTT, ComplexDenseTensorized,  # noqa: F401
                                          ComplexTuckerTensorized)
from .init import tensor_init, cp_init, tucker_init, tt_init, block_tt_init  # noqa: F401
