from .convolution import convolve, tucker_conv  # noqa: F401
from .linear import factorized_linear  # noqa: F401
