"""tensorly.random is only used by reference tests, not by the hot path."""
