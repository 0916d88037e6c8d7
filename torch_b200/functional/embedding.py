"""Factorized embedding lookup (functional form); implementation in torch_b200/embedding_ops.py."""
from ..embedding_ops import cp_rows, tucker_rows, blocktt_rows, parse_row_gather, unravel  # noqa: F401
