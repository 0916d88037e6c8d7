"""tensorly.tenalg.tenalg_utils shim (see ../__init__.py)."""
from oracle.tensorly_semantics import validate_contraction_modes as _validate_contraction_modes  # noqa: F401
