"""The C-ABI library loads and exports every symbol include/tltorch_cuda.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "tltorch_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tlt_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from torch_b200 import _cuda
    assert os.path.exists(_cuda.LIB_PATH)
    handle = ctypes.CDLL(_cuda.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(handle, name), f"{name} is declared in include/tltorch_cuda.h but not exported"
    assert set(declared) == set(_cuda.SYMBOLS), "python binding table and header disagree"
    assert _cuda.lib().tlt_version() == 1


def test_descriptor_structs_match_header_layout(tmp_path):
    """sizeof / offsetof as gcc sees include/tltorch_cuda.h must equal the ctypes mirror."""
    import subprocess
    from torch_b200 import _cuda
    probes = [("tlt_contract_desc", _cuda.ContractDesc, ["dtype_a", "nb", "size_b", "a_b", "b_n", "c_b", "d_n", "alpha", "split_k"]),
              ("tlt_gemm_desc", _cuda.GemmDesc, ["M", "ldc", "a_major", "dtype_c", "dtype_bias", "alpha"]),
              ("tlt_conv_desc", _cuda.ConvDesc, ["dtype", "batch", "in_size", "kernel", "dilation"]),
              ("tlt_cplinear_desc", _cuda.CpLinearDesc, ["batch", "rank", "n_in", "n_out", "in_dims", "out_dims"]),
              ("tlt_cpconv_rows_desc", _cuda.CpConvRowsDesc, ["batch", "rank", "height", "stride"]),
              ("tlt_trl_tail_desc", _cuda.TrlTailDesc, ["batch", "j", "rc", "ldz", "ro", "out"]),
              ("tlt_blocktt3_linear_desc", _cuda.BlockTT3LinearDesc, ["tt", "batch", "transpose", "dtype_bias"]),
              ("tlt_tucker_trl_desc", _cuda.TuckerTrlDesc, ["batch", "channels", "n_spatial", "spatial", "spatial_rank", "rc", "ro", "out"])]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{REPO}/include/tltorch_cuda.h"', "int main(){"]
    for cname, _, fields in probes:
        lines.append(f'printf("%zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'printf("%zu\\n", offsetof({cname}, {f}));')
    lines.append("return 0;}")
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = []
    for _, cls, fields in probes:
        want.append(ctypes.sizeof(cls))
        want += [getattr(cls, f).offset for f in fields]
    assert got == want


def test_invalid_descriptors_are_rejected_with_an_error_string():
    from torch_b200 import _cuda
    lib = _cuda.lib()
    d = _cuda.ContractDesc()
    d.nb = 9
    rc = lib.tlt_contract(ctypes.byref(d), None, None, None, None, None, 0, None)
    assert rc == -1 and b"sub-dimension" in lib.tlt_last_error_string()
    g = _cuda.GemmDesc()
    g.M, g.N, g.K, g.lda, g.ldb, g.ldc = 128, 128, 64, 63, 64, 128      # lda not a multiple of 8
    assert lib.tlt_gemm_bf16_tc_supported(ctypes.byref(g)) == 0
    g.lda = 64
    g.dtype_c = 1
    assert lib.tlt_gemm_bf16_tc_supported(ctypes.byref(g)) == 1
