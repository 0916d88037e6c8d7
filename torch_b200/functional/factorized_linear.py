"""x . W^T with W kept in factorized form: B200 counterparts of tltorch/functional/factorized_linear.py."""
import math

from ..ops import contract
from .. import tenalg
from .. import blocktt_ops
from .. import tucker_ops
from .. import cplinear_ops
from .. import cplinear_tc_ops
from .factorized_tensordot import tensor_dot_tucker, tensor_dot_cp

_L = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"


def linear_tucker(tensor, tucker_matrix, transpose=True, channels_first=True, bias=None):
    """reference: functional/factorized_linear.py:7-21 (``channels_first`` is unused there as well)."""
    axis = 1 if transpose else 0
    n = len(tucker_matrix.tensorized_shape[axis])
    if transpose:
        in_shape, out_shape = list(tucker_matrix.tensorized_shape[1]), list(tucker_matrix.tensorized_shape[0])
        x2 = tensor.reshape(-1, math.prod(in_shape))
        factors = list(tucker_matrix.factors)
        if tucker_ops.tucker3_eligible(x2, tucker_matrix.core, factors, bias, in_shape, out_shape):
            # project (1 launch, on chip) -> core GEMM on tcgen05 -> expand + bias (1 launch); backward: one launch per
            # mode-product chain gives the data gradient and all three factor gradients (tucker_ops.py)
            return tucker_ops.tucker3_linear(x2, tucker_matrix.core, factors, bias, in_shape, out_shape)
    tensor = tensor.reshape(-1, *tucker_matrix.tensorized_shape[axis])
    modes_tensor = list(range(tensor.dim() - n, tensor.dim()))
    modes_tucker = list(range(n, tucker_matrix.order)) if transpose else list(range(n))
    addend = None if bias is None else bias.reshape(tucker_matrix.tensorized_shape[1 - axis])
    return tensor_dot_tucker(tensor, tucker_matrix, (modes_tensor, modes_tucker), addend=addend)


def linear_cp(tensor, cp_matrix, transpose=True, bias=None):
    """reference: functional/factorized_linear.py:23-36."""
    if transpose:
        n_out = len(cp_matrix.tensorized_shape[0])
        in_shape = cp_matrix.tensorized_shape[1]
        modes_cp = list(range(n_out, cp_matrix.order))
    else:
        n_in = len(cp_matrix.tensorized_shape[0])
        in_shape = cp_matrix.tensorized_shape[0]
        modes_cp = list(range(n_in))
    factors = list(cp_matrix.factors)
    f_in = [factors[m] for m in modes_cp]
    f_out = [f for m, f in enumerate(factors) if m not in modes_cp]
    x2 = tensor.reshape(-1, math.prod(in_shape))
    if f_out and cplinear_tc_ops.eligible(x2, cp_matrix.weights, f_out, f_in, bias):
        # every product of the layer as one tcgen05 GEMM on 3-way bf16-split operands (fp32-class accuracy, csrc/split3.cu)
        y = cplinear_tc_ops.cp_linear_tc(x2, cp_matrix.weights, f_out, f_in, bias)
        return y.reshape(-1, *[f.shape[0] for f in f_out])
    if f_out and cplinear_ops.eligible(x2, cp_matrix.weights, f_out, f_in, bias):
        # the whole chain in 2 launches forward / 4 backward, Khatri-Rao tiles rebuilt in shared memory (csrc/cplinear.cu)
        y = cplinear_ops.cp_linear_fused(x2, cp_matrix.weights, f_out, f_in, bias)
        return y.reshape(-1, *[f.shape[0] for f in f_out])
    tensor = tensor.reshape(-1, *in_shape)
    addend = None if bias is None else bias.reshape(cp_matrix.tensorized_shape[0 if transpose else 1])
    return tensor_dot_cp(tensor, cp_matrix, (list(range(1, tensor.dim())), modes_cp), addend=addend)


def blocktt_sweep_macs(in_shape, out_shape, rank):
    """MACs per sample of the left-to-right TT sweep (the order the reference's einsum takes)."""
    n = len(in_shape)
    total = 0
    for k in range(n):
        rows = math.prod(out_shape[:k]) * math.prod(in_shape[k + 1:])
        total += rows * (rank[k] * in_shape[k]) * (out_shape[k] * rank[k + 1])
    return total


def blocktt_plan(batch, in_shape, out_shape, rank):
    """'dense' (rebuild W from the cores once, then one GEMM) or 'sweep' (chain of mode products), by MAC count."""
    dense = math.prod(in_shape) * math.prod(out_shape)
    rebuild = 0
    rows_o, rows_i = 1, 1
    for k in range(len(in_shape) - 1):
        rows_o *= out_shape[k]
        rows_i *= in_shape[k]
        rebuild += rows_o * rows_i * rank[k + 1] * out_shape[k + 1] * in_shape[k + 1] * rank[k + 2]
    dense_per_sample = dense + rebuild / max(batch, 1)
    return "dense" if dense_per_sample <= blocktt_sweep_macs(in_shape, out_shape, rank) else "sweep"


def linear_blocktt(tensor, tt_matrix, transpose=True, plan=None, bias=None):
    """reference: functional/factorized_linear.py:39-60.  Returns (batch, prod(out_shape)).

    The reference's einsum sweeps the cores left to right.  At BERT-FFN shapes that sweep costs
    1.3-3x the FLOPs of the dense product, so a cost model picks between the sweep and "rebuild W on chip-resident
    cores (a few MFLOP) + one tensor-core GEMM"; both produce the same result."""
    axis = 1 if transpose else 0
    in_shape = tuple(tt_matrix.tensorized_shape[axis])
    out_shape = tuple(tt_matrix.tensorized_shape[1 - axis])
    cores = list(tt_matrix.factors)
    n = len(in_shape)
    x = tensor.reshape(-1, *in_shape)
    batch = x.shape[0]
    rank = [c.shape[0] for c in cores] + [cores[-1].shape[-1]]
    if plan is None:
        plan = blocktt_plan(batch, in_shape, out_shape, rank)
    if plan == "dense":
        x2 = x.reshape(batch, -1)
        if blocktt_ops.linear_eligible(x2, cores, bias, transpose):
            # rebuild W (1 launch) + tcgen05 GEMM with the bias in its epilogue; fused core-gradient kernels in backward
            return blocktt_ops.blocktt3_linear(x2, cores[0], cores[1], cores[2], bias, transpose)
        w = tenalg.blocktt_to_tensor(cores, tt_matrix.tensorized_shape).reshape(tt_matrix.shape)     # (rows, cols)
        x2 = x.reshape(batch, -1)
        spec = "bi,oi->bo" if transpose else "bi,io->bo"
        return contract(spec, x2, w, bias, "o") if bias is not None else contract(spec, x2, w)
    # sweep: state = (b, remaining in-modes, produced out-modes, rank)
    ins = list(_L[1:1 + n])
    outs = list(_L[1 + n:1 + 2 * n])
    ranks = list(_L[1 + 2 * n:2 + 3 * n])
    state = "a" + "".join(ins)
    res = x
    for k in range(n):
        core = cores[k] if k else cores[0]
        c_spec = ranks[k] + (outs[k] + ins[k] if transpose else ins[k] + outs[k]) + ranks[k + 1]
        if k == 0:
            res = res.unsqueeze(-1)          # boundary rank r0 = 1
            state = state + ranks[0]
        new_state = "a" + "".join(ins[k + 1:]) + "".join(outs[:k + 1]) + ranks[k + 1]
        if k == n - 1 and bias is not None:
            res = contract(f"{state},{c_spec}->{new_state}", res, core, bias.reshape(*out_shape, 1), new_state[1:])
        else:
            res = contract(f"{state},{c_spec}->{new_state}", res, core)
        state = new_state
    return res.reshape(batch, -1)
