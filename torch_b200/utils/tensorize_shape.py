"""Automatic choice of tensorized shapes (reference: tltorch/utils/tensorize_shape.py:7-98)."""
import math
from bisect import insort_left


def factorize(value, min_value=2, remaining=-1):
    """Smallest-divisor factorization of ``value``; stops early after ``remaining`` factors (the tail stays unsplit)."""
    out = []
    while True:
        if value <= min_value or remaining == 0:
            out.append(value)
            break
        divisor = next((d for d in range(min_value, math.isqrt(value) + 1) if value % d == 0), None)
        if divisor is None or divisor == value:
            out.append(value)
            break
        out.append(divisor)
        value //= divisor
        remaining -= 1
    return tuple(out)


def merge_ints(values, size):
    """Multiply the two smallest entries together until at most ``size`` entries remain."""
    if len(values) <= 1:
        return values
    values = sorted(values)
    while len(values) > size:
        a, b = values[0], values[1]
        values = values[2:]
        insort_left(values, a * b)
    return tuple(values)


def get_tensorized_shape(in_features, out_features, order=None, min_dim=2, verbose=True):
    """Factorize in_features / out_features into the same number (``order``) of integers >= min_dim."""
    in_ten = factorize(in_features, min_value=min_dim)
    out_ten = factorize(out_features, min_value=min_dim, remaining=len(in_ten))
    target = min(len(in_ten), len(out_ten)) if order is None else min(order, len(in_ten), len(out_ten))
    if len(in_ten) > target:
        in_ten = merge_ints(in_ten, size=target)
    if len(out_ten) > target:
        out_ten = merge_ints(out_ten, size=target)
    if verbose:
        print(f"Tensorizing (in, out)=({in_features, out_features}) -> ({in_ten, out_ten})")
    return in_ten, out_ten
