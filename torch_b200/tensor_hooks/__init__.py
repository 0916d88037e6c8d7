from ._tensor_dropout import tensor_dropout, remove_tensor_dropout, TensorDropout  # noqa: F401
