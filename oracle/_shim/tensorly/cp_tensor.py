from oracle.tensorly_semantics import validate_cp_rank, cp_to_tensor  # noqa: F401
from oracle.tensorly_semantics import validate_cp_tensor as _validate_cp_tensor  # noqa: F401
