from .parameter_list import FactorList, ParameterList, ComplexFactorList  # noqa: F401
from .tensorize_shape import get_tensorized_shape  # noqa: F401
