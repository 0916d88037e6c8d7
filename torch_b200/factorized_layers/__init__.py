from .factorized_convolution import FactorizedConv  # noqa: F401
from .tensor_regression_layers import TRL  # noqa: F401
from .tensor_contraction_layers import TCL  # noqa: F401
from .factorized_linear import FactorizedLinear  # noqa: F401
from .factorized_embedding import FactorizedEmbedding  # noqa: F401
