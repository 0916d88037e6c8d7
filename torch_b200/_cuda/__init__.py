"""ctypes binding of libtltorch_cuda.so (the C ABI declared in include/tltorch_cuda.h).

There is NO CPU fallback: every launcher raises if the tensors are not CUDA tensors, and loading fails loudly if the
shared library is missing on a box that has a GPU.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtltorch_cuda.so")
MAX_SUBDIMS = 4
F32, BF16 = 0, 1

_lib = None
PROFILE = None     # bench.py / tools: set to a list to collect (entry point, start event, end event) for every launcher call


class TltError(RuntimeError):
    pass


class ContractDesc(C.Structure):
    _A = C.c_int64 * MAX_SUBDIMS
    _fields_ = [("dtype_a", C.c_int32), ("dtype_b", C.c_int32), ("dtype_c", C.c_int32), ("dtype_d", C.c_int32),
                ("nb", C.c_int32), ("nm", C.c_int32), ("nn", C.c_int32), ("nk", C.c_int32),
                ("size_b", _A), ("size_m", _A), ("size_n", _A), ("size_k", _A),
                ("a_b", _A), ("a_m", _A), ("a_k", _A),
                ("b_b", _A), ("b_k", _A), ("b_n", _A),
                ("c_b", _A), ("c_m", _A), ("c_n", _A),
                ("d_b", _A), ("d_m", _A), ("d_n", _A),
                ("alpha", C.c_float), ("accumulate", C.c_int32), ("split_k", C.c_int32)]


class GemmDesc(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
                ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64),
                ("a_major", C.c_int32), ("b_major", C.c_int32), ("dtype_c", C.c_int32), ("accumulate", C.c_int32),
                ("dtype_bias", C.c_int32), ("alpha", C.c_float), ("split_k", C.c_int32)]


class ConvDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("ndim", C.c_int32), ("batch", C.c_int64), ("channels", C.c_int64),
                ("in_size", C.c_int64 * 3), ("out_size", C.c_int64 * 3),
                ("kernel", C.c_int32 * 3), ("stride", C.c_int32 * 3), ("padding", C.c_int32 * 3),
                ("dilation", C.c_int32 * 3)]


class BlockTT3Desc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("rows", C.c_int32 * 3), ("cols", C.c_int32 * 3),
                ("rank1", C.c_int32), ("rank2", C.c_int32)]


class PointwiseDesc(C.Structure):
    _fields_ = [("batch", C.c_int64), ("in_channels", C.c_int64), ("out_channels", C.c_int64), ("spatial", C.c_int64),
                ("w_ld", C.c_int64), ("dtype_bias", C.c_int32)]


class MmdDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("batch", C.c_int64), ("in_dims", C.c_int32 * 3), ("out_dims", C.c_int32 * 3),
                ("transpose", C.c_int32 * 3)]


class Mmd16Desc(C.Structure):
    _fields_ = [("batch", C.c_int64), ("in_dims", C.c_int32 * 3), ("out_dims", C.c_int32 * 3), ("transpose", C.c_int32 * 3),
                ("ld_in", C.c_int64), ("ld_out", C.c_int64)]


class DwSepDesc(C.Structure):
    _fields_ = [("batch", C.c_int64), ("height", C.c_int64), ("width", C.c_int64), ("ld", C.c_int64), ("stride", C.c_int32)]


class CpConvRowsDesc(C.Structure):
    _fields_ = [("batch", C.c_int64), ("in_channels", C.c_int64), ("out_channels", C.c_int64), ("rank", C.c_int64),
                ("height", C.c_int64), ("width", C.c_int64), ("stride", C.c_int32)]


class CpLinearDesc(C.Structure):
    _fields_ = [("batch", C.c_int64), ("rank", C.c_int64), ("n_in", C.c_int32), ("n_out", C.c_int32),
                ("in_dims", C.c_int64 * 4), ("out_dims", C.c_int64 * 4)]


class BlockTT3LinearDesc(C.Structure):
    _fields_ = [("tt", BlockTT3Desc), ("batch", C.c_int64), ("transpose", C.c_int32), ("dtype_bias", C.c_int32)]


class TuckerTrlDesc(C.Structure):
    _fields_ = [("batch", C.c_int64), ("channels", C.c_int64), ("n_spatial", C.c_int32), ("spatial", C.c_int64 * 3),
                ("spatial_rank", C.c_int64 * 3), ("rc", C.c_int64), ("ro", C.c_int64), ("out", C.c_int64)]


class TrlTailDesc(C.Structure):
    _fields_ = [("batch", C.c_int64), ("j", C.c_int64), ("rc", C.c_int64), ("ldz", C.c_int64), ("ro", C.c_int64), ("out", C.c_int64)]


# every symbol include/tltorch_cuda.h declares: (name, restype, argtypes)
_vp = C.c_void_p
SYMBOLS = {
    "tlt_version": (C.c_int, []),
    "tlt_last_error_string": (C.c_char_p, []),
    "tlt_launch_count": (C.c_uint64, []),
    "tlt_contract_workspace_size": (C.c_size_t, [C.POINTER(ContractDesc)]),
    "tlt_contract": (C.c_int, [C.POINTER(ContractDesc), _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "tlt_contract_pack": (C.c_int, [C.POINTER(ContractDesc), C.c_int32, _vp, _vp, C.c_int64, _vp]),
    "tlt_contract_unpack": (C.c_int, [C.POINTER(ContractDesc), _vp, C.c_int32, C.c_int64, _vp, _vp, _vp]),
    "tlt_gemm_bf16_tc_supported": (C.c_int, [C.POINTER(GemmDesc)]),
    "tlt_gemm_bf16_tc": (C.c_int, [C.POINTER(GemmDesc), _vp, _vp, _vp, _vp, _vp]),
    "tlt_blocktt3_reconstruct": (C.c_int, [C.POINTER(BlockTT3Desc), _vp, _vp, _vp, _vp, _vp]),
    "tlt_blocktt3_core_grads_workspace_size": (C.c_size_t, [C.POINTER(BlockTT3Desc)]),
    "tlt_blocktt3_core_grads": (C.c_int, [C.POINTER(BlockTT3Desc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "tlt_multi_mode_dot": (C.c_int, [C.POINTER(MmdDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tlt_mmd16_fwd": (C.c_int, [C.POINTER(Mmd16Desc), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tlt_mmd16_bwd": (C.c_int, [C.POINTER(Mmd16Desc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tlt_pointwise_tc_supported": (C.c_int, [C.POINTER(PointwiseDesc)]),
    "tlt_pointwise_tc": (C.c_int, [C.POINTER(PointwiseDesc), _vp, _vp, _vp, _vp, _vp]),
    "tlt_pointwise_wgrad_tc": (C.c_int, [C.POINTER(PointwiseDesc), _vp, _vp, _vp, _vp]),
    "tlt_pack_kmajor": (C.c_int, [_vp, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, C.c_int64, _vp]),
    "tlt_im2col": (C.c_int, [C.POINTER(ConvDesc), _vp, _vp, _vp]),
    "tlt_col2im": (C.c_int, [C.POINTER(ConvDesc), _vp, _vp, _vp]),
    "tlt_dwconv_fwd": (C.c_int, [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp]),
    "tlt_dwconv_dgrad": (C.c_int, [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp]),
    "tlt_dwconv_wgrad": (C.c_int, [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp]),
    "tlt_nchw_to_rows": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int64, _vp, C.c_int64, _vp]),
    "tlt_rows_to_nchw": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _vp, _vp]),
    "tlt_im2col_rows": (C.c_int, [C.POINTER(ConvDesc), _vp, _vp, _vp]),
    "tlt_col2im_rows": (C.c_int, [C.POINTER(ConvDesc), _vp, _vp, _vp]),
    "tlt_dwconv_rows_fwd": (C.c_int, [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp]),
    "tlt_dwconv_rows_dgrad": (C.c_int, [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp]),
    "tlt_dwconv_rows_wgrad": (C.c_int, [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp]),
    "tlt_skinny_fwd": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp]),
    "tlt_skinny_bwd": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tlt_scale_modes": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_int64), _vp, C.POINTER(C.c_void_p), C.c_float, _vp, _vp]),
    "tlt_khatri_rao_fwd": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_void_p), _vp, _vp, _vp]),
    "tlt_khatri_rao_bwd": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_void_p), _vp, _vp,
                                     C.POINTER(C.c_void_p), _vp, _vp]),
    "tlt_dwsep_rows_taps": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int32, _vp, _vp]),
    "tlt_dwsep_rows_fwd": (C.c_int, [C.POINTER(DwSepDesc), _vp, _vp, _vp, _vp]),
    "tlt_dwsep_rows_dgrad": (C.c_int, [C.POINTER(DwSepDesc), _vp, _vp, _vp, _vp]),
    "tlt_dwsep_rows_wgrad": (C.c_int, [C.POINTER(DwSepDesc), _vp, _vp, _vp, _vp]),
    "tlt_dwsep_rows_wgrad_parts": (C.c_int, []),
    "tlt_dwsep_rows_finalize": (C.c_int, [_vp, C.c_int32, _vp, _vp, _vp, C.c_int64, C.c_int64, _vp, _vp, _vp, _vp]),
    "tlt_cpchain_prep": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tlt_cpchain_post": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, _vp, _vp, _vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tlt_cpconv_rows_supported": (C.c_int, [C.POINTER(CpConvRowsDesc)]),
    "tlt_cpconv_rows_workspace_size": (C.c_size_t, [C.POINTER(CpConvRowsDesc)]),
    "tlt_cpconv_rows_grad_workspace_size": (C.c_size_t, [C.POINTER(CpConvRowsDesc)]),
    "tlt_cpconv_rows_pack": (C.c_int, [C.POINTER(CpConvRowsDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tlt_cpconv_rows_fwd": (C.c_int, [C.POINTER(CpConvRowsDesc), _vp, _vp, _vp, _vp, _vp]),
    "tlt_cpconv_rows_bwd": (C.c_int, [C.POINTER(CpConvRowsDesc), _vp, _vp, _vp, _vp, _vp, _vp]),
    "tlt_cpconv_rows_finalize": (C.c_int, [C.POINTER(CpConvRowsDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tlt_cplinear_workspace_size": (C.c_size_t, [C.POINTER(CpLinearDesc)]),
    "tlt_cplinear_fwd": (C.c_int, [C.POINTER(CpLinearDesc), _vp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _vp, _vp, _vp, _vp, _vp]),
    "tlt_cplinear_bwd": (C.c_int, [C.POINTER(CpLinearDesc), _vp, _vp, _vp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _vp, _vp, _vp,
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _vp, _vp, _vp]),
    "tlt_trl_tail_fwd": (C.c_int, [C.POINTER(TrlTailDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tlt_trl_tail_bwd": (C.c_int, [C.POINTER(TrlTailDesc), C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tlt_kron2_fwd": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp]),
    "tlt_kron2_bwd": (C.c_int, [_vp, _vp, _vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp]),
    "tlt_khatri_rao_bwd_ld": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_void_p), _vp, _vp, C.c_int64,
                                        C.POINTER(C.c_void_p), _vp, _vp]),
    "tlt_khatri_rao_bwd_all": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_void_p), _vp, _vp, C.c_int64,
                                         C.POINTER(C.c_void_p), _vp, _vp]),
    "tlt_khatri_rao_split3": (C.c_int, [C.c_int32, C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_void_p), _vp, C.c_int32, _vp, C.c_int32,
                                        _vp, _vp]),
    "tlt_split3_bf16": (C.c_int, [_vp, C.c_int64, C.c_int64, C.c_int64, C.c_int32, _vp, C.c_int32, _vp, _vp]),
    "tlt_blocktt3_linear_workspace_size": (C.c_size_t, [C.POINTER(BlockTT3LinearDesc)]),
    "tlt_blocktt3_linear_grad_workspace_size": (C.c_size_t, [C.POINTER(BlockTT3LinearDesc)]),
    "tlt_blocktt3_linear_fwd": (C.c_int, [C.POINTER(BlockTT3LinearDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tlt_blocktt3_linear_bwd": (C.c_int, [C.POINTER(BlockTT3LinearDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tlt_tucker_trl_workspace_size": (C.c_size_t, [C.POINTER(TuckerTrlDesc)]),
    "tlt_tucker_trl_grad_workspace_size": (C.c_size_t, [C.POINTER(TuckerTrlDesc)]),
    "tlt_tucker_trl_fwd": (C.c_int, [C.POINTER(TuckerTrlDesc), _vp, _vp, C.POINTER(C.c_void_p), _vp, _vp, _vp, _vp]),
    "tlt_tucker_trl_bwd": (C.c_int, [C.POINTER(TuckerTrlDesc), _vp, _vp, _vp, C.POINTER(C.c_void_p), _vp, _vp, _vp, _vp,
                                     C.POINTER(C.c_void_p), _vp, _vp, _vp]),
    "tlt_spatial_project_supported": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, C.c_int32]),
    "tlt_spatial_project_fwd": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp]),
    "tlt_spatial_project_bwd": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tlt_cast": (C.c_int, [_vp, C.c_int32, _vp, C.c_int32, C.c_int64, _vp]),
    "tlt_l2_flush": (C.c_int, [_vp, C.c_size_t, _vp]),
}


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TltError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU or PyTorch fallback for the factorized-layer hot path)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.tlt_version() != 1:
            raise TltError(f"ABI version mismatch: library reports {handle.tlt_version()}")
        _lib = _Library(handle)
    return _lib


class _Library:
    """Attribute access to the C entry points.  Launchers (status-returning entry points that take a stream) are wrapped so
    that, while ``PROFILE`` is a list, each call is bracketed by a CUDA-event pair recorded on torch's current stream -- the
    stream every launcher is given (``stream_ptr``) -- which is how bench.py times individual kernels inside a step."""

    def __init__(self, handle):
        self._handle = handle

    def __getattr__(self, name):
        fn = getattr(self._handle, name)
        res, args = SYMBOLS.get(name, (None, []))
        is_launcher = res is C.c_int and name not in ("tlt_version", "tlt_dwsep_rows_wgrad_parts") and not name.endswith("_supported")
        if is_launcher:
            def call(*a, _fn=fn, _name=name):
                if PROFILE is None:
                    return _fn(*a)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = _fn(*a)
                e1.record()
                PROFILE.append((_name, e0, e1))
                return rc
            fn = call
        setattr(self, name, fn)
        return fn


def check(rc, what):
    if rc != 0:
        msg = lib().tlt_last_error_string()
        raise TltError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")


def dtype_code(dt):
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise TypeError(f"the tltorch CUDA path computes in float32 or bfloat16, got {dt}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NotImplementedError(
                "tltorch-b200 runs the factorized-layer hot path on CUDA (sm_100a) only; got a CPU tensor. "
                "There is deliberately no CPU fallback (the CPU oracle under oracle/ is test infrastructure).")


def stream_ptr(t):
    """torch's current stream on ``t``'s device.  Kernels launch on the CURRENT device, so a tensor living on another GPU
    would be launched on the wrong one: refuse instead (wrap the call in ``torch.cuda.device(t.device)``)."""
    if t.device.index is not None and t.device.index != torch.cuda.current_device():
        raise TltError(f"tensor on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}: "
                       "launches go to the current device -- use `with torch.cuda.device(tensor.device):`")
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def launch_count():
    return int(lib().tlt_launch_count())
