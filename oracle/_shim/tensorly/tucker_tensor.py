from oracle.tensorly_semantics import validate_tucker_rank, tucker_to_tensor  # noqa: F401
from oracle.tensorly_semantics import validate_tucker_tensor as _validate_tucker_tensor  # noqa: F401
