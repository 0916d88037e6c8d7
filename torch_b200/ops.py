"""PyTorch custom ops (with autograd) over the C ABI of libtltorch_cuda.so.

Three primitive families carry the whole factorized-layer hot path:

  tltorch::contract   pairwise strided tensor contraction (einsum-like spec) -> tlt_contract (SIMT FFMA, any strides)
                      or, for plain bf16 GEMM shapes, tlt_gemm_bf16_tc (tcgen05 tensor cores, TMA, TMEM)
  tltorch::im2col / tltorch::col2im     unfold for dense (non-grouped) N-D convolutions -> tlt_im2col / tlt_col2im
  tltorch::dwconv (+ dgrad / wgrad)     depthwise N-D convolution                        -> tlt_dwconv_*

The backward of every op is expressed with the same ops, so gradients flow to activations, factors, cores, weights
and biases exactly as autograd does for the reference's einsum / conv calls.  CPU tensors raise
NotImplementedError: there is no fallback path.
"""
import ctypes as C
import functools
import math
from typing import List, Optional, Sequence

import torch

from . import _cuda

__all__ = ["contract", "im2col", "col2im", "dwconv", "launch_count"]

launch_count = _cuda.launch_count

# minimum M*N*K for which a plain bf16 GEMM is routed to the tcgen05 kernel instead of the SIMT contraction kernel
TC_MIN_WORK = 1 << 21
import os as _os
_FORCE_SIMT = _os.environ.get("TLT_FORCE_SIMT", "0") == "1"
_LOG_SIMT = _os.environ.get("TLT_LOG_SIMT", "0") == "1"     # debugging aid: report large contractions that stay on the SIMT kernel
import sys as _sys   # debugging aid: route every contraction to the SIMT kernel


PROFILE = None          # set to a list by bench.py to collect (kernel, start_event, end_event, flops, bytes) per launch


# =====================================================================================================================
# contraction planning (host logic; pure python, exercised on CPU by tests/test_host_logic.py)
# =====================================================================================================================
class ContractPlan:
    __slots__ = ("out_shape", "classes", "desc_fields", "tc", "spec", "swap")

    def __init__(self):
        self.tc = None
        self.swap = False


def _parse(spec):
    try:
        lhs, out = spec.split("->")
        a, b = lhs.split(",")
    except ValueError:
        raise ValueError(f"contract spec must look like 'ab,bc->ac', got {spec!r}")
    for s in (a, b, out):
        if len(set(s)) != len(s):
            raise ValueError(f"repeated index inside one operand is not supported: {spec!r}")
    return a, b, out


def _merge(letters, size, strides_by_op):
    """Drop size-1 letters, then fuse neighbours (outer p, inner q) when stride[p] == stride[q]*size[q] in every operand.
    Returns (sizes, [strides per operand])."""
    letters = [l for l in letters if size[l] != 1]
    sizes = [size[l] for l in letters]
    strides = [[st[l] for l in letters] for st in strides_by_op]
    i = 0
    while i + 1 < len(sizes):
        if all(st[i] == st[i + 1] * sizes[i + 1] for st in strides):
            sizes[i] *= sizes[i + 1]
            del sizes[i + 1]
            for st in strides:
                st[i] = st[i + 1]
                del st[i + 1]
        else:
            i += 1
    return sizes, strides


@functools.lru_cache(maxsize=8192)
def plan_contract(spec, a_shape, a_stride, b_shape, b_stride, d_letters, d_stride):
    """Classify indices (batch / m / n / k), order and merge sub-dims, and produce the tlt_contract_desc fields."""
    la, lb, lo = _parse(spec)
    if len(la) != len(a_shape) or len(lb) != len(b_shape):
        raise ValueError(f"spec {spec!r} does not match operand ranks {len(a_shape)}, {len(b_shape)}")
    size = {}
    for letters, shape in ((la, a_shape), (lb, b_shape)):
        for l, s in zip(letters, shape):
            if size.setdefault(l, s) != s:
                raise ValueError(f"size mismatch for index {l!r} in {spec!r}: {size[l]} vs {s}")
    for l in lo:
        if l not in size:
            raise ValueError(f"output index {l!r} of {spec!r} appears in no operand")
    # The kernels put the column class (indices of the SECOND operand) innermost in their store pattern, and the tensor-
    # core GEMM needs a row-major C.  If the innermost non-trivial output index belongs to the first operand only,
    # swap the operand roles (C^T = B^T A^T): same contraction, same output tensor.
    swap = False
    inner = next((l for l in reversed(lo) if size[l] != 1), None)
    if inner is not None and inner in la and inner not in lb:
        la, lb = lb, la
        a_shape, b_shape = b_shape, a_shape
        a_stride, b_stride = b_stride, a_stride
        swap = True
    sa = dict(zip(la, a_stride))
    sb = dict(zip(lb, b_stride))
    out_shape = tuple(size[l] for l in lo)
    so, acc = {}, 1
    for l in reversed(lo):
        so[l] = acc
        acc *= size[l]
    sd = {l: 0 for l in lo}
    if d_letters is not None:
        for l, st in zip(d_letters, d_stride):
            if l not in so:
                raise ValueError(f"addend index {l!r} is not an output index of {spec!r}")
            sd[l] = st
    batch = [l for l in lo if l in sa and l in sb]
    mm = [l for l in lo if l in sa and l not in sb]
    nn = [l for l in lo if l in sb and l not in sa]
    kk = [l for l in la if l in sb and l not in so]
    for l in la:
        if l not in sb and l not in so:
            raise ValueError(f"index {l!r} of the first operand is neither contracted nor kept in {spec!r}")
    for l in lb:
        if l not in sa and l not in so:
            raise ValueError(f"index {l!r} of the second operand is neither contracted nor kept in {spec!r}")
    # output-ordered (outer -> inner) for b/m/n; A-stride-ordered for k
    kk.sort(key=lambda l: -abs(sa[l]))
    sz_b, (a_b, b_b, c_b, d_b) = _merge(batch, size, [sa, sb, so, sd])
    sz_m, (a_m, c_m, d_m) = _merge(mm, size, [sa, so, sd])
    sz_n, (b_n, c_n, d_n) = _merge(nn, size, [sb, so, sd])
    sz_k, (a_k, b_k) = _merge(kk, size, [sa, sb])
    for name, s in (("batch", sz_b), ("row", sz_m), ("column", sz_n), ("reduction", sz_k)):
        if len(s) > _cuda.MAX_SUBDIMS:
            raise ValueError(f"{spec!r}: {len(s)} non-mergeable {name} sub-dimensions exceed the kernel limit "
                             f"of {_cuda.MAX_SUBDIMS}; make an operand contiguous first")
    plan = ContractPlan()
    plan.swap = swap
    plan.spec = spec
    plan.out_shape = out_shape
    plan.classes = dict(b=sz_b, m=sz_m, n=sz_n, k=sz_k)
    plan.desc_fields = dict(nb=len(sz_b), nm=len(sz_m), nn=len(sz_n), nk=len(sz_k),
                            size_b=sz_b, size_m=sz_m, size_n=sz_n, size_k=sz_k,
                            a_b=a_b, a_m=a_m, a_k=a_k, b_b=b_b, b_k=b_k, b_n=b_n,
                            c_b=c_b, c_m=c_m, c_n=c_n, d_b=d_b, d_m=d_m, d_n=d_n)
    # ---- plain-GEMM view for the tcgen05 kernel -------------------------------------------------------------------
    if len(sz_b) == 0 and len(sz_m) == 1 and len(sz_n) == 1 and len(sz_k) == 1:
        M, N, K = sz_m[0], sz_n[0], sz_k[0]
        am, ak, bk, bn, cm, cn = a_m[0], a_k[0], b_k[0], b_n[0], c_m[0], c_n[0]
        a_major = 0 if ak == 1 else (1 if am == 1 else None)
        b_major = 0 if bk == 1 else (1 if bn == 1 else None)
        if a_major is not None and b_major is not None and cn == 1:
            bias_ok = d_letters is None or (d_m[0] == 0 and d_n[0] == 1)
            plan.tc = dict(M=M, N=N, K=K, lda=am if a_major == 0 else ak, ldb=bn if b_major == 0 else bk, ldc=cm,
                           a_major=a_major, b_major=b_major, bias_ok=bias_ok)
    return plan


def _fill_desc(plan, dt_a, dt_b, dt_c, dt_d, alpha):
    d = _cuda.ContractDesc()
    d.dtype_a, d.dtype_b, d.dtype_c, d.dtype_d = dt_a, dt_b, dt_c, dt_d
    f = plan.desc_fields
    d.nb, d.nm, d.nn, d.nk = f["nb"], f["nm"], f["nn"], f["nk"]
    for name in ("size_b", "size_m", "size_n", "size_k", "a_b", "a_m", "a_k", "b_b", "b_k", "b_n",
                 "c_b", "c_m", "c_n", "d_b", "d_m", "d_n"):
        arr = getattr(d, name)
        vals = f[name]
        for i in range(_cuda.MAX_SUBDIMS):
            arr[i] = vals[i] if i < len(vals) else (1 if name.startswith("size") else 0)
    d.alpha = alpha
    d.accumulate = 0
    d.split_k = 0
    return d


# =====================================================================================================================
# executors (the ONLY functions that touch the shared library; tests on CPU swap them for descriptor emulators)
# =====================================================================================================================
def _execute_contract(plan, a, b, d, out, alpha):
    _cuda.require_cuda(a, b, d, out)
    if plan.swap:
        a, b = b, a
    lib = _cuda.lib()
    dt_d = _cuda.dtype_code(d.dtype) if d is not None else 0
    tc = plan.tc
    if (tc is not None and not _FORCE_SIMT and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
            and (d is None or tc["bias_ok"]) and tc["M"] * tc["N"] * tc["K"] >= TC_MIN_WORK):
        g = _cuda.GemmDesc()
        g.M, g.N, g.K = tc["M"], tc["N"], tc["K"]
        g.lda, g.ldb, g.ldc = tc["lda"], tc["ldb"], tc["ldc"]
        g.a_major, g.b_major = tc["a_major"], tc["b_major"]
        g.dtype_c = _cuda.dtype_code(out.dtype)
        g.accumulate = 0
        g.dtype_bias = dt_d
        g.alpha = alpha
        if (lib.tlt_gemm_bf16_tc_supported(C.byref(g)) and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0
                and out.data_ptr() % 16 == 0):
            if PROFILE is not None:      # bench.py: CUDA-event pair around this launch, on the launching stream
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            _cuda.check(lib.tlt_gemm_bf16_tc(C.byref(g), _cuda.ptr(a), _cuda.ptr(b), _cuda.ptr(d), _cuda.ptr(out),
                                             _cuda.stream_ptr(out)), f"tlt_gemm_bf16_tc[{plan.spec}]")
            if PROFILE is not None:
                e1.record()
                PROFILE.append((f"gemm_bf16_tc M{tc['M']} N{tc['N']} K{tc['K']} a{tc['a_major']}b{tc['b_major']}", e0, e1,
                                2.0 * tc["M"] * tc["N"] * tc["K"],
                                2.0 * (tc["M"] * tc["K"] + tc["N"] * tc["K"]) + out.element_size() * tc["M"] * tc["N"]))
            return
    desc = _fill_desc(plan, _cuda.dtype_code(a.dtype), _cuda.dtype_code(b.dtype), _cuda.dtype_code(out.dtype), dt_d, alpha)
    if (not _FORCE_SIMT and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and out.dtype == torch.bfloat16
            and _packed_tensor_core_contract(plan, desc, a, b, d, out, alpha, lib)):
        return
    if _LOG_SIMT:
        f = plan.desc_fields
        work = math.prod(f["size_b"] or [1]) * math.prod(f["size_m"] or [1]) * math.prod(f["size_n"] or [1]) * math.prod(f["size_k"] or [1])
        if work >= (1 << 22):
            print(f"[tlt simt] {plan.spec} {a.dtype} b{f['size_b']} m{f['size_m']} n{f['size_n']} k{f['size_k']} macs={work:.3g}",
                  file=_sys.stderr)
    ws_bytes = lib.tlt_contract_workspace_size(C.byref(desc))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=out.device) if ws_bytes else None
    _cuda.check(lib.tlt_contract(C.byref(desc), _cuda.ptr(a), _cuda.ptr(b), _cuda.ptr(d), _cuda.ptr(out),
                                 _cuda.ptr(ws), ws_bytes, _cuda.stream_ptr(out)), f"tlt_contract[{plan.spec}]")


# large bf16 contractions whose operands are strided / misaligned for TMA: pack -> tcgen05 GEMM -> unpack
PACK_MIN_WORK = 1 << 26


def _direct_matrix(sizes_r, strides_r, sizes_k, strides_k, t):
    """(major, ld) if the operand already is a TMA-friendly matrix (single row dim, single reduction dim), else None."""
    if len(sizes_r) != 1 or len(sizes_k) != 1 or t.data_ptr() % 16:
        return None
    sr, sk = strides_r[0], strides_k[0]
    if sk == 1 and sr % 8 == 0 and sr >= sizes_k[0]:
        return 0, sr
    if sr == 1 and sk % 8 == 0 and sk >= sizes_r[0]:
        return 1, sk
    return None


def _packed_tensor_core_contract(plan, desc, a, b, d, out, alpha, lib):
    f = plan.desc_fields
    if f["nb"] != 0 or not f["size_m"] or not f["size_n"] or not f["size_k"]:
        return False
    M, N, K = math.prod(f["size_m"]), math.prod(f["size_n"]), math.prod(f["size_k"])
    # skinny outputs with a very long reduction (factor / weight gradients): still worth the tensor cores when the K
    # blocks can be split over all SM pairs (fp32 partials met by TMA reduce-add)
    long_k = K >= 16384 and M * N <= 512 * 512
    if M * N * K < (PACK_MIN_WORK >> 2 if long_k else PACK_MIN_WORK) or K < 8 or (max(M, N) < 128 and not long_k):
        return False
    st = _cuda.stream_ptr(out)
    kp = (K + 7) // 8 * 8
    g = _cuda.GemmDesc()
    g.M, g.N, g.K = M, N, K
    keep = []
    da = _direct_matrix(f["size_m"], f["a_m"], f["size_k"], f["a_k"], a)
    if da is None:
        ap = torch.empty((M, kp), dtype=torch.bfloat16, device=out.device)
        _cuda.check(lib.tlt_contract_pack(C.byref(desc), 0, _cuda.ptr(a), _cuda.ptr(ap), kp, st), "tlt_contract_pack[A]")
        keep.append(ap)
        a_ptr, g.a_major, g.lda = _cuda.ptr(ap), 0, kp
    else:
        a_ptr, (g.a_major, g.lda) = _cuda.ptr(a), da
    db = _direct_matrix(f["size_n"], f["b_n"], f["size_k"], f["b_k"], b)
    if db is None:
        bp = torch.empty((N, kp), dtype=torch.bfloat16, device=out.device)
        _cuda.check(lib.tlt_contract_pack(C.byref(desc), 1, _cuda.ptr(b), _cuda.ptr(bp), kp, st), "tlt_contract_pack[B]")
        keep.append(bp)
        b_ptr, g.b_major, g.ldb = _cuda.ptr(bp), 0, kp
    else:
        b_ptr, (g.b_major, g.ldb) = _cuda.ptr(b), db
    direct_c = (d is None and not long_k and len(f["size_m"]) == 1 and len(f["size_n"]) == 1 and f["c_n"][0] == 1
                and f["c_m"][0] % 8 == 0 and out.data_ptr() % 16 == 0)
    if direct_c:
        cp, g.ldc = out, f["c_m"][0]
    else:
        np_ = (N + 7) // 8 * 8
        cp = torch.empty((M, np_), dtype=torch.float32 if long_k else out.dtype, device=out.device)
        g.ldc = np_
    g.dtype_c = _cuda.dtype_code(cp.dtype)
    g.accumulate, g.dtype_bias, g.alpha, g.split_k = 0, 0, alpha, 0
    if not lib.tlt_gemm_bf16_tc_supported(C.byref(g)):
        return False
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _cuda.check(lib.tlt_gemm_bf16_tc(C.byref(g), a_ptr, b_ptr, None, _cuda.ptr(cp), st), f"tlt_gemm_bf16_tc[packed {plan.spec}]")
    if PROFILE is not None:
        e1.record()
        PROFILE.append((f"gemm_bf16_tc(packed) M{M} N{N} K{K}", e0, e1, 2.0 * M * N * K, 2.0 * (M * K + N * K + M * N)))
    if not direct_c:
        _cuda.check(lib.tlt_contract_unpack(C.byref(desc), _cuda.ptr(cp), _cuda.dtype_code(cp.dtype), g.ldc, _cuda.ptr(d),
                                            _cuda.ptr(out), st),
                    "tlt_contract_unpack")
    return True


def _conv_desc(dtype, batch, channels, in_size, out_size, kernel, stride, padding, dilation):
    d = _cuda.ConvDesc()
    d.dtype = _cuda.dtype_code(dtype) if dtype in (torch.float32, torch.bfloat16) else -1   # -1: rejected by the library
    n = len(kernel)
    d.ndim = n
    d.batch, d.channels = batch, channels
    for i in range(3):
        d.in_size[i] = in_size[i] if i < n else 1
        d.out_size[i] = out_size[i] if i < n else 1
        d.kernel[i] = kernel[i] if i < n else 1
        d.stride[i] = stride[i] if i < n else 1
        d.padding[i] = padding[i] if i < n else 0
        d.dilation[i] = dilation[i] if i < n else 1
    return d


def _execute_conv(which, desc, *tensors):
    _cuda.require_cuda(*tensors)
    fn = getattr(_cuda.lib(), which)
    args = [_cuda.ptr(t) for t in tensors]
    _cuda.check(fn(C.byref(desc), *args, _cuda.stream_ptr(tensors[0])), which)


# =====================================================================================================================
# tltorch::contract
# =====================================================================================================================
@torch.library.custom_op("tltorch::contract", mutates_args=())
def _contract_op(a: torch.Tensor, b: torch.Tensor, d: Optional[torch.Tensor], spec: str, d_letters: str,
                 alpha: float) -> torch.Tensor:
    plan = plan_contract(spec, tuple(a.shape), tuple(a.stride()), tuple(b.shape), tuple(b.stride()),
                         d_letters if d is not None else None, tuple(d.stride()) if d is not None else None)
    out = torch.empty(plan.out_shape, dtype=a.dtype, device=a.device)
    if out.numel():
        _execute_contract(plan, a, b, d, out, alpha)
    return out


@_contract_op.register_fake
def _(a, b, d, spec, d_letters, alpha):
    la, lb, lo = _parse(spec)
    size = dict(zip(la, a.shape))
    size.update(zip(lb, b.shape))
    return a.new_empty([size[l] for l in lo])


def _contract_setup(ctx, inputs, output):
    a, b, d, spec, d_letters, alpha = inputs
    ctx.save_for_backward(a, b)
    ctx.spec, ctx.d_letters, ctx.alpha = spec, d_letters, alpha
    ctx.has_d = d is not None
    ctx.d_shape = tuple(d.shape) if d is not None else None


def _contract_backward(ctx, gy):
    a, b = ctx.saved_tensors
    la, lb, lo = _parse(ctx.spec)
    ga = gb = gd = None
    if ctx.needs_input_grad[0]:
        ga = contract(f"{lo},{lb}->{la}", gy, b, alpha=ctx.alpha)
    if ctx.needs_input_grad[1]:
        gb = contract(f"{la},{lo}->{lb}", a, gy, alpha=ctx.alpha)
    if ctx.has_d and ctx.needs_input_grad[2]:
        gd = reduce_to(gy, lo, ctx.d_letters)
    return ga, gb, gd, None, None, None


_contract_op.register_autograd(_contract_backward, setup_context=_contract_setup)


def contract(spec: str, a: torch.Tensor, b: torch.Tensor, addend: Optional[torch.Tensor] = None,
             addend_letters: str = "", alpha: float = 1.0) -> torch.Tensor:
    """out[spec] = alpha * sum_k a * b (+ addend broadcast along ``addend_letters``).

    ``spec`` is a two-operand einsum string.  Indices shared by both operands and the output are batch indices; no
    operand is ever permuted or copied -- strides go straight into the kernel descriptor."""
    if a.is_complex() or b.is_complex() or (addend is not None and addend.is_complex()):
        return _complex_contract(spec, a, b, addend, addend_letters, float(alpha))
    if a.dtype != b.dtype:
        raise TypeError(f"contract: operand dtypes differ ({a.dtype} vs {b.dtype})")
    if addend is not None and len(addend_letters) != addend.dim():
        raise ValueError("addend_letters must name every dim of the addend")
    return _contract_op(a, b, addend, spec, addend_letters, float(alpha))


def _complex_contract(spec, a, b, addend, addend_letters, alpha):
    """Complex operands (the Complex* factorized tensors, reference: complex_factorized_tensors.py): the real kernels run on the
    real / imaginary planes -- re = ar.br - ai.bi, im = ar.bi + ai.br, the second product of each pair accumulating onto the
    first through the kernel's addend -- and the planes are joined again.  Up to 4 launches, still no CPU path."""
    out_letters = spec.split("->")[1]

    def planes(t):
        if t is None:
            return None, None
        if t.is_complex():
            return t.real.contiguous(), t.imag.contiguous()
        return t, None

    (ar, ai), (br, bi), (dr, di) = planes(a), planes(b), planes(addend)
    real_dtype = ar.dtype

    def to_real(t):
        return None if t is None else t.to(real_dtype)

    br, bi, dr, di = to_real(br), to_real(bi), to_real(dr), to_real(di)

    def acc(x, y, sign, base, base_letters):
        if x is None or y is None:
            return base, base_letters
        return (_contract_op(x, y, base, spec, base_letters if base is not None else "", sign * alpha), out_letters)

    re, re_l = acc(ar, br, 1.0, dr, addend_letters)
    re, re_l = acc(ai, bi, -1.0, re, re_l)
    im, im_l = acc(ar, bi, 1.0, di, addend_letters)
    im, im_l = acc(ai, br, 1.0, im, im_l)
    if im is None:
        im = torch.zeros_like(re)
    elif im_l != out_letters:           # only the addend had an imaginary part: broadcast it to the output's shape
        im = im.reshape([im.shape[im_l.index(l)] if l in im_l else 1 for l in out_letters]).expand_as(re).contiguous()
    return torch.complex(re, im)


def reduce_to(t: torch.Tensor, letters: str, keep: str) -> torch.Tensor:
    """Sum ``t`` (dims named by ``letters``) over every index not in ``keep`` -- a contraction with a stride-0 ones tensor."""
    drop = "".join(l for l in letters if l not in keep)
    if not drop:
        perm = [letters.index(l) for l in keep]
        return t.permute(perm)
    if len(drop) > 1 and letters[-1] in drop and t.is_contiguous() and t.shape[-1] >= 8:
        # e.g. the bias gradient of an NCHW conv ("bos" -> "o"): first the contiguous innermost index (row sums, coalesced),
        # then the rest (a vector-matrix reduction) -- instead of one composite-k reduction on the generic kernel
        inner = reduce_to(t, letters, letters[:-1])
        return reduce_to(inner, letters[:-1], keep)
    sizes = [t.shape[letters.index(l)] for l in drop]
    ones = torch.ones((), dtype=t.dtype, device=t.device).expand(sizes)
    return contract(f"{letters},{drop}->{keep}", t, ones)


# =====================================================================================================================
# unfold ops for dense N-D convolutions
# =====================================================================================================================
def conv_out_size(in_size, kernel, stride, padding, dilation):
    return [(i + 2 * p - d * (k - 1) - 1) // s + 1 for i, k, s, p, d in zip(in_size, kernel, stride, padding, dilation)]


@torch.library.custom_op("tltorch::im2col", mutates_args=())
def _im2col_op(x: torch.Tensor, kernel: List[int], stride: List[int], padding: List[int],
               dilation: List[int]) -> torch.Tensor:
    x = x.contiguous()
    B, Cc = x.shape[0], x.shape[1]
    in_size = list(x.shape[2:])
    out_size = conv_out_size(in_size, kernel, stride, padding, dilation)
    cols = torch.empty((B * math.prod(out_size), Cc * math.prod(kernel)), dtype=x.dtype, device=x.device)
    if cols.numel():
        desc = _conv_desc(x.dtype, B, Cc, in_size, out_size, kernel, stride, padding, dilation)
        _execute_conv("tlt_im2col", desc, x, cols)
    return cols


@_im2col_op.register_fake
def _(x, kernel, stride, padding, dilation):
    out_size = conv_out_size(list(x.shape[2:]), kernel, stride, padding, dilation)
    return x.new_empty((x.shape[0] * math.prod(out_size), x.shape[1] * math.prod(kernel)))


@torch.library.custom_op("tltorch::col2im", mutates_args=())
def _col2im_op(cols: torch.Tensor, x_shape: List[int], kernel: List[int], stride: List[int], padding: List[int],
               dilation: List[int]) -> torch.Tensor:
    cols = cols.contiguous()
    B, Cc = x_shape[0], x_shape[1]
    in_size = list(x_shape[2:])
    out_size = conv_out_size(in_size, kernel, stride, padding, dilation)
    dx = torch.empty(x_shape, dtype=cols.dtype, device=cols.device)
    if dx.numel():
        desc = _conv_desc(cols.dtype, B, Cc, in_size, out_size, kernel, stride, padding, dilation)
        _execute_conv("tlt_col2im", desc, cols, dx)
    return dx


@_col2im_op.register_fake
def _(cols, x_shape, kernel, stride, padding, dilation):
    return cols.new_empty(x_shape)


def _im2col_setup(ctx, inputs, output):
    x, ctx.kernel, ctx.stride, ctx.padding, ctx.dilation = inputs
    ctx.x_shape = list(x.shape)


def _im2col_backward(ctx, g):
    return _col2im_op(g, ctx.x_shape, ctx.kernel, ctx.stride, ctx.padding, ctx.dilation), None, None, None, None


def _col2im_setup(ctx, inputs, output):
    _, _, ctx.kernel, ctx.stride, ctx.padding, ctx.dilation = inputs


def _col2im_backward(ctx, g):
    return _im2col_op(g, ctx.kernel, ctx.stride, ctx.padding, ctx.dilation), None, None, None, None, None


_im2col_op.register_autograd(_im2col_backward, setup_context=_im2col_setup)
_col2im_op.register_autograd(_col2im_backward, setup_context=_col2im_setup)


def im2col(x, kernel, stride, padding, dilation):
    return _im2col_op(x, list(kernel), list(stride), list(padding), list(dilation))


def col2im(cols, x_shape, kernel, stride, padding, dilation):
    return _col2im_op(cols, list(x_shape), list(kernel), list(stride), list(padding), list(dilation))


# =====================================================================================================================
# depthwise N-D convolution
# =====================================================================================================================
@torch.library.custom_op("tltorch::dwconv", mutates_args=())
def _dwconv_op(x: torch.Tensor, w: torch.Tensor, stride: List[int], padding: List[int],
               dilation: List[int]) -> torch.Tensor:
    x, w = x.contiguous(), w.contiguous()
    B, Cc = x.shape[0], x.shape[1]
    kernel = list(w.shape[1:])
    in_size = list(x.shape[2:])
    out_size = conv_out_size(in_size, kernel, stride, padding, dilation)
    y = torch.empty((B, Cc, *out_size), dtype=x.dtype, device=x.device)
    if y.numel():
        desc = _conv_desc(x.dtype, B, Cc, in_size, out_size, kernel, stride, padding, dilation)
        _execute_conv("tlt_dwconv_fwd", desc, x, w, y)
    return y


@_dwconv_op.register_fake
def _(x, w, stride, padding, dilation):
    out_size = conv_out_size(list(x.shape[2:]), list(w.shape[1:]), stride, padding, dilation)
    return x.new_empty((x.shape[0], x.shape[1], *out_size))


@torch.library.custom_op("tltorch::dwconv_dgrad", mutates_args=())
def _dwconv_dgrad_op(gy: torch.Tensor, w: torch.Tensor, x_shape: List[int], stride: List[int], padding: List[int],
                     dilation: List[int]) -> torch.Tensor:
    gy, w = gy.contiguous(), w.contiguous()
    kernel = list(w.shape[1:])
    dx = torch.empty(x_shape, dtype=gy.dtype, device=gy.device)
    if dx.numel():
        desc = _conv_desc(gy.dtype, x_shape[0], x_shape[1], list(x_shape[2:]), list(gy.shape[2:]), kernel, stride,
                          padding, dilation)
        _execute_conv("tlt_dwconv_dgrad", desc, gy, w, dx)
    return dx


@_dwconv_dgrad_op.register_fake
def _(gy, w, x_shape, stride, padding, dilation):
    return gy.new_empty(x_shape)


@torch.library.custom_op("tltorch::dwconv_wgrad", mutates_args=())
def _dwconv_wgrad_op(x: torch.Tensor, gy: torch.Tensor, kernel: List[int], stride: List[int], padding: List[int],
                     dilation: List[int]) -> torch.Tensor:
    x, gy = x.contiguous(), gy.contiguous()
    # fp32 accumulation buffer (the kernel adds per-block partial sums with fp32 atomics); fp64 only ever reaches this
    # point under the CPU descriptor emulator of the test-suite
    acc_dtype = torch.float64 if x.dtype == torch.float64 else torch.float32
    dw = torch.zeros((x.shape[1], *kernel), dtype=acc_dtype, device=x.device)
    if x.numel() and gy.numel():
        desc = _conv_desc(x.dtype, x.shape[0], x.shape[1], list(x.shape[2:]), list(gy.shape[2:]), kernel, stride,
                          padding, dilation)
        _execute_conv("tlt_dwconv_wgrad", desc, x, gy, dw)
    return dw if x.dtype == acc_dtype else dw.to(x.dtype)


@_dwconv_wgrad_op.register_fake
def _(x, gy, kernel, stride, padding, dilation):
    return x.new_empty((x.shape[1], *kernel))


def _dwconv_setup(ctx, inputs, output):
    x, w, ctx.stride, ctx.padding, ctx.dilation = inputs
    ctx.save_for_backward(x, w)


def _dwconv_backward(ctx, gy):
    x, w = ctx.saved_tensors
    dx = dw = None
    if ctx.needs_input_grad[0]:
        dx = _dwconv_dgrad_op(gy, w, list(x.shape), ctx.stride, ctx.padding, ctx.dilation)
    if ctx.needs_input_grad[1]:
        dw = _dwconv_wgrad_op(x, gy, list(w.shape[1:]), ctx.stride, ctx.padding, ctx.dilation)
    return dx, dw, None, None, None


_dwconv_op.register_autograd(_dwconv_backward, setup_context=_dwconv_setup)


def dwconv(x, w, stride, padding, dilation):
    """Depthwise conv: x (B, C, *S), w (C, *kernel) -> (B, C, *out)."""
    return _dwconv_op(x, w, list(stride), list(padding), list(dilation))
