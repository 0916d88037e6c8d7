import numpy as np
import torch


def _np(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def assert_array_almost_equal(a, b, decimal=6, err_msg="", **kw):
    np.testing.assert_array_almost_equal(_np(a), _np(b), decimal=decimal, err_msg=err_msg)


def assert_array_equal(a, b, **kw):
    np.testing.assert_array_equal(_np(a), _np(b))


def assert_(cond, msg=""):
    assert cond, msg


assert_equal = np.testing.assert_equal
assert_raises = np.testing.assert_raises
