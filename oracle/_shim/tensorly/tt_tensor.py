from oracle.tensorly_semantics import validate_tt_rank, tt_to_tensor  # noqa: F401
from oracle.tensorly_semantics import validate_tt_tensor as _validate_tt_tensor  # noqa: F401
