from oracle.tensorly_semantics import validate_tt_matrix_rank  # noqa: F401
