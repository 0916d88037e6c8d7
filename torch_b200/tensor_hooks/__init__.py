from ._tensor_lasso import tensor_lasso, remove_tensor_lasso, TensorLasso  # noqa: F401
from ._tensor_dropout import tensor_dropout, remove_tensor_dropout, TensorDropout  # noqa: F401
