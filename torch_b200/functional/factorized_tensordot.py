"""Dense (x) factorized contractions: B200 counterparts of tltorch/functional/factorized_tensordot.py.

The reference turns each call into ONE n-ary ``tl.einsum`` whose left-to-right evaluation starts with outer products
(SURVEY.md section 0.3).  Here the same contraction is evaluated as a short chain of ``ops.contract`` launches in a
FLOP-sane order (contract the dense tensor with the factors of the contracted modes first, then the core / weights,
then expand onto the free modes).  The result is the same tensor, index for index.
"""
from ..ops import contract
from .. import tenalg

_L = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def _validate_contraction_modes(shape1, shape2, modes):
    if isinstance(modes, int):
        modes = [modes]
    modes = list(modes)
    if len(modes) == 2 and not isinstance(modes[0], int) and not isinstance(modes[1], int):
        m1, m2 = list(modes[0]), list(modes[1])
    else:
        m1, m2 = list(modes), list(modes)
    if len(m1) != len(m2):
        raise ValueError("Can only contract two tensors along the same number of modes")
    m1 = [m + len(shape1) if m < 0 else m for m in m1]
    m2 = [m + len(shape2) if m < 0 else m for m in m2]
    for a, b in zip(m1, m2):
        if shape1[a] != shape2[b]:
            raise ValueError(f"Contraction dimensions must have the same dimensions: {shape1[a]} != {shape2[b]}")
    return m1, m2


def tensor_dot_tucker(tensor, tucker, modes, batched_modes=(), addend=None):
    """Contract ``tensor`` with a Tucker tensor along ``modes`` (reference: factorized_tensordot.py:10-66).

    Output modes: the kept modes of ``tensor`` followed by the un-contracted modes of ``tucker``.  ``addend`` (shaped
    like the free modes, e.g. a bias) is fused into the epilogue of the last launch."""
    if batched_modes not in ((), [], None):
        raise NotImplementedError("batched_modes is never exercised by the reference layers (SURVEY App. C.9)")
    modes_tensor, modes_tucker = _validate_contraction_modes(tuple(tensor.shape), tuple(tucker.tensor_shape), modes)
    core, factors = tucker.core, list(tucker.factors)
    # 1. project the contracted modes of the dense tensor onto the rank space, in place (mode stays where it is)
    res = tenalg.multi_mode_dot(tensor, [factors[mk] for mk in modes_tucker], modes_tensor, transpose=True)
    # 2. contract with the core: kept tensor modes, then free core modes
    nt = tensor.dim()
    t_letters = list(_L[:nt])
    core_letters = list(_L[nt:nt + core.dim()])
    for mt, mk in zip(modes_tensor, modes_tucker):
        core_letters[mk] = t_letters[mt]
    kept = [l for i, l in enumerate(t_letters) if i not in modes_tensor]
    free = [l for i, l in enumerate(core_letters) if i not in modes_tucker]
    res = contract(f"{''.join(t_letters)},{''.join(core_letters)}->{''.join(kept + free)}", res, core)
    # 3. expand the free modes onto their factors
    free_modes = [i for i in range(core.dim()) if i not in modes_tucker]
    if free_modes:
        res = tenalg.multi_mode_dot(res, [factors[mk] for mk in free_modes], [len(kept) + pos for pos in range(len(free_modes))],
                                    addend=addend)
    if addend is not None and not free_modes:
        raise ValueError("addend needs at least one free mode")
    return res


def tensor_dot_cp(tensor, cp, modes, addend=None):
    """Contract ``tensor`` with a CP tensor along ``modes`` (reference: factorized_tensordot.py:70-103).

    z[kept, r] = <tensor, KhatriRao(contracted factors)>;  out = (z * lambda) . KhatriRao(free factors)^T."""
    try:
        cp_shape = cp.tensor_shape
    except AttributeError:
        cp_shape = cp.shape
    modes_tensor, modes_cp = _validate_contraction_modes(tuple(tensor.shape), tuple(cp_shape), modes)
    factors = list(cp.factors)
    nt = tensor.dim()
    t_letters = _L[:nt]
    kept = "".join(l for i, l in enumerate(t_letters) if i not in modes_tensor)
    contracted = "".join(t_letters[m] for m in modes_tensor)
    kr_in = tenalg.khatri_rao_tensor([factors[m] for m in modes_cp])                  # (contracted dims..., R)
    z = contract(f"{t_letters},{contracted}z->{kept}z", tensor, kr_in)                # (kept..., R)
    free_modes = [i for i in range(len(factors)) if i not in modes_cp]
    if not free_modes:
        return contract(f"{kept}z,z->{kept}", z, cp.weights)
    kr_out = tenalg.khatri_rao_tensor([factors[m] for m in free_modes], cp.weights)   # (free dims..., R), lambda folded in
    free = _L[nt:nt + len(free_modes)]
    if addend is not None:
        return contract(f"{kept}z,{free}z->{kept}{free}", z, kr_out, addend, free)
    return contract(f"{kept}z,{free}z->{kept}{free}", z, kr_out)
