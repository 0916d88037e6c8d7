"""tensorly.tenalg shim (see ../__init__.py)."""
from oracle.tensorly_semantics import mode_dot, multi_mode_dot, inner, tensordot, khatri_rao  # noqa: F401
from . import tenalg_utils  # noqa: F401
