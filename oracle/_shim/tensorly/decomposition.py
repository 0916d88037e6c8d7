"""Init-time decompositions are OUT OF SCOPE of the hot path."""


def _unavailable(*a, **k):
    raise NotImplementedError("tensorly decompositions are not restated: init-time only, outside the hot path")


parafac = tucker = tensor_train = tensor_train_matrix = _unavailable
