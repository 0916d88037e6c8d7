"""GPU parity tests (run on the B200 with -m gpu): the CUDA path, called through the C ABI, against
  (a) golden fixtures produced by the real reference (tests/golden/*.pt),
  (b) the CPU oracle on seeded inputs at sizes the oracle finishes in seconds,
  (c) size-independent properties at BASELINE.json's full sizes.
Tolerances are north_star's: relative Frobenius error <= 1e-5 (fp32), <= 2e-2 (bf16)."""
import math

import pytest
import torch

import _golden as G
import _cases

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-5, torch.bfloat16: 2e-2}


def dev():
    return torch.device("cuda:0")


def q(t, dtype):
    """values representable in `dtype`, held in fp64 on the CPU (what the oracle consumes)"""
    return t.to(dtype).double().cpu()


# ------------------------------------------------------------------------------------------------- (a) goldens
def _tag():
    import inspect
    return inspect.stack()[1].function


def _report(tag, errs, tol):
    """Print the measured relative Frobenius errors (pytest -rP / -s shows them; the worst one is in the assertion message)."""
    worst = max(errs.items(), key=lambda kv: kv[1])
    print(f"[parity] {tag}: max rel err {worst[1]:.3e} ({worst[0]}) tol {tol:g}; " + ", ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    assert worst[1] < tol, f"{tag}: {worst[0]} rel err {worst[1]:.3e} >= {tol:g}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name,sub", _cases.all_cases())
def test_golden_cases(name, sub, dtype):
    """Goldens hold bf16-representable inputs (tests/golden/make_golden.py), so fp32 AND bf16 runs consume exactly the numbers the
    reference consumed: north_star's tolerances apply unchanged to y and to every gradient."""
    import torch_b200 as tlt
    g, sub2 = _cases.load_case(name, sub)
    before = tlt.ops.launch_count()
    y, grads, ey, eg = _cases.run_case(name, g, dev(), dtype, sub2)
    assert tlt.ops.launch_count() > before, "no kernel of libtltorch_cuda.so was launched"
    errs = {"y": G.rel_err(y, ey.reshape(y.shape))}
    for n, gr in grads.items():
        errs[n] = G.rel_err(gr, eg[n])
    _report(f"{name}/{sub}/{dtype}", errs, TOL[dtype])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", _cases.module_cases())
def test_golden_module_cases(name, dtype):
    """SURVEY 8(f1) / (f2) on the GPU: n_layers > 1 FactorizedLinear / FactorizedConv (reference: factorized_linear.py:103-113,
    factorized_convolution.py:198-208, tensorized_matrices.py:136-193,386-489) and FactorizedEmbedding (factorized_embedding.py:96-117)
    built from the reference's state_dict, against outputs of the real reference."""
    import torch_b200 as tlt
    g = G.load(name)
    before = tlt.ops.launch_count()
    y, grads, ey, eg = _cases.run_module_case(name, g, dev(), dtype)
    assert tlt.ops.launch_count() > before, "no kernel of libtltorch_cuda.so was launched"
    errs = {"y": G.rel_err(y, ey.reshape(y.shape))}
    assert set(grads) == set(eg)
    for n, gr in grads.items():
        errs[n] = G.rel_err(gr, eg[n])
    _report(f"{name}/{dtype}", errs, TOL[dtype])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_golden_reconstruction(dtype):
    allg = G.load("reconstruction")
    errs = {}
    for key, g in allg.items():
        w, _ = _cases.build_weight(g["kind"], g["params"], dev(), dtype, g.get("tensorized_shape"))
        full = w.to_matrix() if key.startswith("matrix_") else w.to_tensor()
        errs[key] = G.rel_err(full, g["full"])
    _report(f"reconstruction/{dtype}", errs, TOL[dtype])


# ------------------------------------------------------------------------------------------------- (b) oracle, seeded
def _contract_vs_einsum(spec, a_shape, b_shape, dtype, addend=None, a_perm=None, b_perm=None, alpha=1.0):
    from torch_b200.ops import contract
    torch.manual_seed(hash(spec) % 1000)
    a = torch.randn(a_shape, device=dev(), dtype=dtype)
    b = torch.randn(b_shape, device=dev(), dtype=dtype)
    if a_perm:
        a = a.permute(a_perm).contiguous().permute([a_perm.index(i) for i in range(len(a_perm))])
    if b_perm:
        b = b.permute(b_perm).contiguous().permute([b_perm.index(i) for i in range(len(b_perm))])
    want = torch.einsum(spec, a.double().cpu(), b.double().cpu()) * alpha
    d = None
    if addend:
        letters = spec.split("->")[1]
        d = torch.randn([want.shape[letters.index(l)] for l in addend], device=dev(), dtype=dtype)
        view = [want.shape[i] if l in addend else 1 for i, l in enumerate(letters)]
        want = want + d.double().cpu().reshape(view)
    got = contract(spec, a, b, d, addend or "", alpha)
    return G.rel_err(got, want)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_contract_kernel_shapes_and_edges(dtype):
    tol = TOL[dtype] if dtype == torch.float32 else 1e-2
    cases = [
        ("mk,kn->mn", (70, 33), (33, 65), {}),                       # ragged tiles
        ("mk,nk->mn", (128, 64), (96, 64), {}),                      # both K-major
        ("km,kn->mn", (40, 130), (40, 17), {}),                      # both MN-major, small N tile
        ("bmk,bkn->bmn", (3, 20, 7), (3, 7, 9), {}),                 # batched
        ("abcd,dcef->abef", (4, 3, 5, 6), (6, 5, 2, 7), {}),         # composite K (2 sub-dims, different order)
        ("abc,bd->adc", (6, 7, 8), (7, 5), {}),                      # mode product writing a middle mode
        ("ar,br->abr", (5, 16), (7, 16), {}),                        # Khatri-Rao (K empty, batch r)
        ("a,a->", (1000,), (1000,), {}),                             # full reduction to a scalar
        ("mk,kn->mn", (3000, 5), (5, 4), {}),                        # tall-skinny (128x16 tile)
        ("mk,kn->mn", (7, 5), (5, 3000), {}),                        # short-wide (16x128 tile)
        ("mk,kn->mn", (8, 40000), (40000, 8), {}),                   # split-K + workspace
        ("mk,kn->mn", (64, 32), (32, 48), dict(addend="n")),         # bias along n
        ("mk,kn->mn", (64, 32), (32, 48), dict(addend="m", alpha=0.5)),
        ("mk,kn->nm", (33, 20), (20, 17), {}),                       # transposed output
        ("mk,kn->mn", (0, 4), (4, 5), {}),                           # empty M
        ("mk,kn->mn", (4, 0), (0, 5), dict(addend="n")),             # empty K: output is the addend
        ("abc,bd->adc", (6, 7, 8), (7, 5), dict(a_perm=(2, 0, 1))),  # non-contiguous operand
        ("ko,k->o", (5000, 300), (5000,), {}),                       # vector-matrix path, matrix contiguous along j
        ("ok,k->o", (300, 5000), (5000,), dict(addend="o")),         # vector-matrix path, matrix contiguous along k
        ("bko,bk->bo", (3, 2000, 40), (3, 2000), {}),                # batched vector-matrix
        ("k,kon->on", (700,), (700, 9, 11), dict(alpha=2.0)),        # vector first, composite j
    ]
    for spec, sa, sb, kw in cases:
        err = _contract_vs_einsum(spec, sa, sb, dtype, **kw)
        assert err < tol, (spec, sa, sb, kw, err)


@pytest.mark.parametrize("shape", [(256, 256, 128, 0, 0), (512, 512, 64, 0, 0), (4096, 768, 3072, 0, 0), (768, 3072, 4096, 1, 1), (300, 520, 200, 0, 0), (128, 128, 64, 1, 1), (384, 136, 4096, 1, 1),
                                   (1000, 264, 72, 0, 1), (136, 1024, 256, 1, 0), (2048, 3072, 768, 0, 0)])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_tcgen05_gemm_against_fp64(shape, out_dtype):
    """tlt_gemm_bf16_tc directly through the C ABI: all four operand-major combinations, ragged M/N/K tails, bias."""
    import ctypes as C
    from torch_b200 import _cuda
    M, N, K, a_mn, b_mn = shape
    torch.manual_seed(M + N + K)
    a = torch.randn((K, M) if a_mn else (M, K), device=dev(), dtype=torch.bfloat16)
    b = torch.randn((K, N) if b_mn else (N, K), device=dev(), dtype=torch.bfloat16)
    bias = torch.randn(N, device=dev(), dtype=torch.float32)
    c = torch.full((M, N), float("nan"), device=dev(), dtype=out_dtype)
    g = _cuda.GemmDesc()
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldc = a.shape[1], b.shape[1], N
    g.a_major, g.b_major = a_mn, b_mn
    g.dtype_c = _cuda.dtype_code(out_dtype)
    g.dtype_bias = _cuda.F32
    g.alpha = 0.5
    lib = _cuda.lib()
    assert lib.tlt_gemm_bf16_tc_supported(C.byref(g)) == 1
    _cuda.check(lib.tlt_gemm_bf16_tc(C.byref(g), _cuda.ptr(a), _cuda.ptr(b), _cuda.ptr(bias), _cuda.ptr(c),
                                     _cuda.stream_ptr(c)), "tlt_gemm_bf16_tc")
    torch.cuda.synchronize()
    A = (a.t() if a_mn else a).double().cpu()
    Bm = (b.t() if b_mn else b).double().cpu()
    want = 0.5 * (A @ Bm.t()) + bias.double().cpu()
    assert not torch.isnan(c.float()).any()
    assert G.rel_err(c, want) < (1e-5 if out_dtype == torch.float32 else 4e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(768, 3072, 8192, 1, 1, 0), (768, 3072, 8192, 1, 1, 4), (300, 520, 4104, 0, 0, 3),
                                   (512, 264, 2048, 1, 0, 16), (256, 256, 1024, 0, 1, 2)])
@pytest.mark.parametrize("accumulate", [0, 1])
def test_tcgen05_gemm_split_k(shape, accumulate):
    """Split-K work units: fp32 partials meet in C through TMA reduce-add (C zeroed by the call unless accumulate)."""
    import ctypes as C
    from torch_b200 import _cuda
    M, N, K, a_mn, b_mn, split = shape
    torch.manual_seed(M + N + K + split)
    a = torch.randn((K, M) if a_mn else (M, K), device=dev(), dtype=torch.bfloat16)
    b = torch.randn((K, N) if b_mn else (N, K), device=dev(), dtype=torch.bfloat16)
    bias = torch.randn(N, device=dev(), dtype=torch.float32)
    c0 = torch.randn(M, N, device=dev(), dtype=torch.float32)
    c = c0.clone() if accumulate else torch.full((M, N), float("nan"), device=dev(), dtype=torch.float32)
    g = _cuda.GemmDesc()
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldc = a.shape[1], b.shape[1], N
    g.a_major, g.b_major = a_mn, b_mn
    g.dtype_c = _cuda.F32
    g.dtype_bias = _cuda.F32
    g.alpha = 0.25
    g.accumulate = accumulate
    g.split_k = split
    lib = _cuda.lib()
    _cuda.check(lib.tlt_gemm_bf16_tc(C.byref(g), _cuda.ptr(a), _cuda.ptr(b), _cuda.ptr(bias), _cuda.ptr(c),
                                     _cuda.stream_ptr(c)), "tlt_gemm_bf16_tc")
    torch.cuda.synchronize()
    A = (a.t() if a_mn else a).double().cpu()
    Bm = (b.t() if b_mn else b).double().cpu()
    want = 0.25 * (A @ Bm.t()) + bias.double().cpu()
    if accumulate:
        want = want + c0.double().cpu()
    assert not torch.isnan(c).any()
    assert G.rel_err(c, want) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,cols,r1,r2", [((8, 12, 32), (8, 8, 12), 16, 16), ((8, 8, 12), (8, 12, 32), 16, 16),
                                             ((3, 5, 7), (2, 6, 5), 4, 9), ((4, 4, 4), (4, 4, 4), 1, 1),
                                             ((2, 3, 50), (3, 2, 3), 8, 5)])
def test_blocktt3_fused_kernels_vs_fp64(rows, cols, r1, r2, dtype):
    """tlt_blocktt3_reconstruct / tlt_blocktt3_core_grads (one fused launch / two) against the fp64 einsum definition
    of BlockTT.to_tensor (tensorized_matrices.py:354-384) and its autograd, incl. ragged o3 groups."""
    from torch_b200 import blocktt_ops
    torch.manual_seed(sum(rows) + sum(cols) + r1 + r2)
    shapes = [(1, rows[0], cols[0], r1), (r1, rows[1], cols[1], r2), (r2, rows[2], cols[2], 1)]
    cores = [(torch.randn(s, device=dev()) * 0.5).to(dtype).requires_grad_(True) for s in shapes]
    W = blocktt_ops.blocktt3_to_matrix(*cores)
    gW = torch.randn_like(W)
    grads = torch.autograd.grad(W, cores, gW)
    ref = [c.detach().double().cpu().requires_grad_(True) for c in cores]
    Wr = torch.einsum("aoib,bpjc,cqkd->opqijk", *ref).reshape(W.shape)
    gr = torch.autograd.grad(Wr, ref, gW.double().cpu())
    tol = TOL[dtype]
    assert G.rel_err(W, Wr) < tol
    for a, b in zip(grads, gr):
        assert G.rel_err(a, b) < tol


@pytest.mark.parametrize("transpose", [True, False])
def test_blocktt3_fused_linear_vs_oracle(transpose):
    """The fused dense-regime route (reconstruct + tcgen05 GEMMs + split-K dW + fused core grads) vs the oracle's
    restatement of linear_blocktt (functional/factorized_linear.py:39-60), fwd and every gradient, bf16."""
    from oracle import reference_path as rp
    from torch_b200 import blocktt_ops
    torch.manual_seed(11)
    rows, cols = (8, 12, 32), (8, 8, 12)
    shapes = [(1, rows[0], cols[0], 16), (16, rows[1], cols[1], 16), (16, rows[2], cols[2], 1)]
    cores = [(torch.randn(s, device=dev()) * 0.3).bfloat16().requires_grad_(True) for s in shapes]
    fin, fout = (768, 3072) if transpose else (3072, 768)
    x = torch.randn(1000, fin, device=dev()).bfloat16().requires_grad_(True)
    bias = torch.randn(fout, device=dev()).bfloat16().requires_grad_(True)
    assert blocktt_ops.linear_eligible(x, cores, bias, transpose)
    before = tlt_launches()
    y = blocktt_ops.blocktt3_linear(x, *cores, bias, transpose)
    gy = torch.randn_like(y)
    got = torch.autograd.grad(y, [x, bias] + cores, gy)
    assert tlt_launches() - before >= 7
    xo = x.detach().double().cpu().requires_grad_(True)
    bo = bias.detach().double().cpu().requires_grad_(True)
    co = [c.detach().double().cpu().requires_grad_(True) for c in cores]
    Wo = torch.einsum("aoib,bpjc,cqkd->opqijk", *co).reshape(3072, 768)
    yo = (xo @ Wo.t() if transpose else xo @ Wo) + bo
    if transpose:
        assert G.rel_err(yo, rp.blocktt_linear(xo, co, cols) + bo) < 1e-12        # the oracle agrees with the definition
    want = torch.autograd.grad(yo, [xo, bo] + co, gy.double().cpu())
    assert G.rel_err(y, yo) < 2e-2
    for a, b in zip(got, want):
        assert G.rel_err(a, b) < 2e-2


@pytest.mark.parametrize("B,K,N,S", [(3, 64, 69, 3136), (2, 69, 64, 784), (5, 130, 300, 64), (2, 512, 573, 200), (1, 8, 8, 8),
                                     (4, 93, 128, 784),
                                     # transposed-role kernel (N <= 128, S >= 2048): ragged last position tile, N = 128 exactly,
                                     # tiny N / K with a 16-channel tail block, K = 96 (1.5 k blocks)
                                     (2, 48, 128, 2056), (1, 17, 5, 4104), (2, 96, 33, 2048)])
def test_pointwise_tc_kernels_vs_fp64(B, K, N, S):
    """tlt_pointwise_tc (fwd / dgrad) and tlt_pointwise_wgrad_tc straight from NCHW memory, ragged channel counts and
    partial position tiles, against fp64 einsum (the F.conv1d(k=1) stages of functional/convolution.py)."""
    from torch_b200 import pointwise_ops
    torch.manual_seed(B + K + N + S)
    x = torch.randn(B, K, S, device=dev()).bfloat16().requires_grad_(True)
    w_kn = (torch.randn(K, N, device=dev()) / math.sqrt(K)).bfloat16().requires_grad_(True)    # stored (in, out): w_nk is a strided view
    bias = torch.randn(N, device=dev()).bfloat16().requires_grad_(True)
    y = pointwise_ops.pointwise_tc(x, w_kn.t(), bias)
    gy = torch.randn_like(y)
    gx, gw, gb = torch.autograd.grad(y, [x, w_kn, bias], gy)
    xr, wr, br = (t.detach().double().cpu().requires_grad_(True) for t in (x, w_kn, bias))
    yr = torch.einsum("kn,bks->bns", wr, xr) + br[None, :, None]
    want = torch.autograd.grad(yr, [xr, wr, br], gy.double().cpu())
    assert G.rel_err(y, yr) < 1e-2
    for a, b in zip((gx, gw, gb), want):
        assert G.rel_err(a, b) < 1e-2


@pytest.mark.parametrize("spec,sa,sb,bias_letters", [
    ("mk,nk->mn", (1024, 3375), (3375, 3375), None),                 # Tucker core GEMM of config 5: K, N not multiples of 8
    ("bijk,ri->brjk", (256, 64, 16, 16), (60, 64), None),            # mode product along a middle mode
    ("oc,bcs->bos", (388, 512), (64, 512, 49), "o"),                 # 1x1 conv on 7x7 maps (S = 49 breaks the TMA pitch rule) + bias
    ("bsct,qct->bqs", (32, 49, 388, 9), (388, 388, 9), None),        # unfolded Tucker core conv, K = 388 * 9
    ("bk,kn->bn", (4096, 1000), (1000, 264), "n"),
    ("bijk,brjk->ir", (512, 16, 16, 16), (512, 15, 16, 16), None),   # factor gradient of a mode product: 16 x 15 output, K = 131072
])
def test_packed_tensor_core_contract(spec, sa, sb, bias_letters):
    """Strided / TMA-misaligned bf16 contractions take pack -> tcgen05 GEMM -> unpack; result against fp64 einsum."""
    from torch_b200 import ops
    torch.manual_seed(len(spec) + sum(sa))
    a = (torch.randn(sa, device=dev()) * 0.5).bfloat16()
    b = (torch.randn(sb, device=dev()) * 0.5).bfloat16()
    lo = spec.split("->")[1]
    size = dict(zip(spec.split("->")[0].split(",")[0], sa))
    size.update(zip(spec.split("->")[0].split(",")[1], sb))
    bias = torch.randn([size[l] for l in bias_letters], device=dev()).bfloat16() if bias_letters else None
    ops.PROFILE = []
    try:
        y = ops.contract(spec, a, b, bias, bias_letters) if bias is not None else ops.contract(spec, a, b)
        torch.cuda.synchronize()
        took_tc = any("gemm_bf16_tc" in e[0] for e in ops.PROFILE)
    finally:
        ops.PROFILE = None
    assert took_tc, "the contraction did not reach the tensor cores"
    want = torch.einsum(spec, a.double().cpu(), b.double().cpu())
    if bias is not None:
        shape = [size[l] if l in bias_letters else 1 for l in lo]
        want = want + bias.double().cpu().reshape(shape)
    assert tuple(y.shape) == tuple(want.shape)
    assert G.rel_err(y, want) < 1e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,C,H,W,k,stride,pad,dil", [
    (3, 5, 56, 56, (3, 3), (1, 1), (1, 1), (1, 1)),      # fast path, 16-byte staging
    (2, 7, 28, 28, (3, 3), (1, 1), (1, 1), (1, 1)),      # fast path, 8-byte staging (bf16)
    (4, 6, 14, 14, (3, 3), (1, 1), (1, 1), (1, 1)),      # fast path, ragged last strip
    (5, 9, 7, 7, (3, 3), (1, 1), (1, 1), (1, 1)),        # fast path, scalar staging
    (37, 69, 56, 56, (3, 3), (1, 1), (1, 1), (1, 1)),    # TMA-fed path (bf16): many groups per persistent block, tail group
    (9, 141, 28, 28, (3, 3), (1, 1), (1, 1), (1, 1)),    # TMA-fed path (bf16): 10 planes per group
    (3, 5, 8, 12, (3, 3), (1, 1), (1, 1), (1, 1)),       # TMA-fed path (bf16): tiny planes
    (2, 4, 10, 13, (3, 3), (1, 1), (1, 1), (1, 1)),
    (3, 7, 23, 64, (3, 3), (1, 1), (1, 1), (1, 1)),      # tensor-core path (bf16): W = 64 (last tile odd), ragged 16-row slabs
    (5, 3, 17, 32, (3, 3), (1, 1), (1, 1), (1, 1)),      # tensor-core path: W = 32
    (1, 1, 5, 16, (3, 3), (1, 1), (1, 1), (1, 1)),       # tensor-core path: W = 16, a single short plane
    (2, 9, 64, 56, (3, 3), (1, 1), (1, 1), (1, 1)),      # tensor-core path: W = 56 with H = 64 (four full slabs), tail group
    (3, 5, 56, 56, (3, 3), (2, 2), (1, 1), (1, 1)),      # generic staged path: stride 2
    (2, 3, 17, 9, (3, 1), (1, 1), (1, 0), (1, 1)),       # 1-D taps along H
    (2, 3, 12, 11, (2, 3), (1, 2), (0, 2), (2, 1)),      # dilation, asymmetric everything
])
def test_depthwise_conv_kernels_vs_torch(B, C, H, W, k, stride, pad, dil, dtype):
    """tlt_dwconv_{fwd,dgrad,wgrad} (3x3 fast path and generic smem-staged path) against fp64 F.conv2d(groups=C) autograd --
    the depthwise stages of cp_conv / cp_conv_mobilenet (functional/convolution.py:268-271, 321-332)."""
    from torch_b200 import ops
    torch.manual_seed(B * C + H + W)
    x = torch.randn(B, C, H, W, device=dev()).to(dtype).requires_grad_(True)
    w = (torch.randn(C, *k, device=dev()) * 0.5).to(dtype).requires_grad_(True)
    y = ops.dwconv(x, w, list(stride), list(pad), list(dil))
    gy = torch.randn_like(y)
    gx, gw = torch.autograd.grad(y, [x, w], gy)
    xr, wr = x.detach().double().cpu().requires_grad_(True), w.detach().double().cpu().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr, wr.unsqueeze(1), stride=stride, padding=pad, dilation=dil, groups=C)
    gxr, gwr = torch.autograd.grad(yr, [xr, wr], gy.double().cpu())
    tol = TOL[dtype]
    assert tuple(y.shape) == tuple(yr.shape)
    assert G.rel_err(y, yr) < tol
    assert G.rel_err(gx, gxr) < tol
    assert G.rel_err(gw, gwr) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,ranks,modes,transpose,with_bias", [
    ((512, 16, 16, 16), (15, 15, 15), [1, 2, 3], True, False),       # config 5 Tucker projection (per token 16^3 -> 15^3)
    ((512, 15, 15, 15), (16, 16, 16), [1, 2, 3], False, True),       # ... and the expansion, with the bias fused
    ((256, 40, 7, 7), (24, 4, 4), [1, 2, 3], False, True),           # TCL on (C, H, W)
    ((300, 3, 9, 5), (4, 7), [2, 3], True, False),                   # only two trailing modes, an untouched middle dim ahead
    ((128, 33), (70,), [1], False, False),                           # a single mode with r > 64 (several register tiles)
    ((200, 5, 6, 7), (3, 2), [1, 3], False, False),                  # a skipped mode in the middle
])
def test_fused_multi_mode_dot_vs_fp64(shape, ranks, modes, transpose, with_bias, dtype):
    """tlt_multi_mode_dot (one launch, intermediates in shared memory) and its autograd against fp64 tensordot chains --
    tenalg.multi_mode_dot of TCL / tucker_trl / linear_tucker (tensor_contraction_layers.py:75, tensor_regression.py:40)."""
    from torch_b200 import tenalg, ops, modedot_ops
    monkey = modedot_ops.MIN_SAMPLE_ELEMS
    modedot_ops.MIN_SAMPLE_ELEMS = 1
    torch.manual_seed(sum(shape) + len(modes))
    x = torch.randn(shape, device=dev()).to(dtype).requires_grad_(True)
    mats = [(torch.randn((shape[m], r) if transpose else (r, shape[m]), device=dev()) / math.sqrt(shape[m])).to(dtype).requires_grad_(True)
            for r, m in zip(ranks, modes)]
    out_tail = list(shape[min(modes):])
    for r, m in zip(ranks, modes):
        out_tail[m - min(modes)] = r
    bias = torch.randn(out_tail, device=dev()).to(dtype).requires_grad_(True) if with_bias else None
    before = ops.launch_count()
    y = tenalg.multi_mode_dot(x, mats, modes, transpose=transpose, addend=bias)
    assert ops.launch_count() - before == 1, "expected ONE fused launch"
    gy = torch.randn_like(y)
    leaves = [x] + mats + ([bias] if with_bias else [])
    got = torch.autograd.grad(y, leaves, gy)
    modedot_ops.MIN_SAMPLE_ELEMS = monkey
    ref_leaves = [t.detach().double().cpu().requires_grad_(True) for t in leaves]
    ref = ref_leaves[0]
    for m, mode in zip(ref_leaves[1:1 + len(mats)], modes):
        mm = m.t() if transpose else m
        ref = torch.movedim(torch.tensordot(ref, mm, dims=([mode], [1])), -1, mode)
    if with_bias:
        ref = ref + ref_leaves[-1]
    want = torch.autograd.grad(ref, ref_leaves, gy.double().cpu())
    tol = TOL[dtype]
    assert G.rel_err(y, ref) < tol
    for a, b in zip(got, want):
        assert G.rel_err(a, b) < tol


def tlt_launches():
    from torch_b200 import ops
    return ops.launch_count()


def test_tcgen05_and_simt_agree_through_contract():
    from torch_b200 import ops
    torch.manual_seed(3)
    x = torch.randn(512, 768, device=dev(), dtype=torch.bfloat16)
    w = torch.randn(3072, 768, device=dev(), dtype=torch.bfloat16)
    before = ops.launch_count()
    y_tc = ops.contract("bi,oi->bo", x, w)
    ops._FORCE_SIMT = True
    try:
        y_simt = ops.contract("bi,oi->bo", x, w)
    finally:
        ops._FORCE_SIMT = False
    assert ops.launch_count() - before == 2
    assert G.rel_err(y_tc, y_simt.double()) < 1e-2
    want = x.double().cpu() @ w.double().cpu().t()
    assert G.rel_err(y_tc, want) < 4e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("fact", ["cp", "tucker", "blocktt"])
def test_factorized_linear_layer_vs_oracle(fact, dtype):
    """FactorizedLinear module fwd+bwd vs the oracle (fp64, same rounded parameter values)."""
    import torch_b200 as tlt
    from oracle import reference_path as rp
    torch.manual_seed(11)
    in_ts, out_ts = (4, 8, 4), (8, 4, 8)
    layer = tlt.FactorizedLinear(in_ts, out_ts, bias=True, factorization=fact, rank=0.3 if fact != "blocktt" else 8,
                                 device=dev()).to(dtype)
    if fact == "cp":
        layer.weight.weights.data.uniform_(0.5, 1.5)
    x = torch.randn(96, 128, device=dev(), dtype=dtype, requires_grad=True)
    y = layer(x)
    gy = torch.randn_like(y)
    params = dict(layer.named_parameters())
    grads = torch.autograd.grad(y, [x] + list(params.values()), gy)
    # oracle
    leaves = {n: p.detach().double().cpu().requires_grad_(True) for n, p in params.items()}
    xo = x.detach().double().cpu().requires_grad_(True)
    fs = [leaves[f"weight.factors.factor_{i}"] for i in range(len(layer.weight.factors))]
    if fact == "cp":
        p = (leaves["weight.weights"], fs)
    elif fact == "tucker":
        p = (leaves["weight.core"], fs)
    else:
        p = fs
    yo = rp.factorized_linear(xo, fact, p, (out_ts, in_ts), bias=leaves["bias"], order="sane")
    go = torch.autograd.grad(yo, [xo] + list(leaves.values()), gy.double().cpu())
    tol = TOL[dtype]
    assert G.rel_err(y, yo) < tol
    for n, a, b in zip(["x"] + list(params), grads, go):
        assert G.rel_err(a, b) < tol, n


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("fact,impl", [("cp", "factorized"), ("cp", "mobilenet"), ("tucker", "factorized"),
                                       ("tt", "factorized"), ("cp", "reconstructed")])
def test_factorized_conv_layer_vs_oracle(fact, impl, dtype):
    import torch_b200 as tlt
    from oracle import reference_path as rp
    torch.manual_seed(5)
    conv = tlt.FactorizedConv(16, 24, 3, order=2, stride=2, padding=1, bias=True, factorization=fact, rank=0.5,
                              implementation=impl, device=dev())
    conv.reset_parameters(std=0.3)
    conv.bias.data.normal_()
    conv = conv.to(dtype)
    x = torch.randn(4, 16, 14, 14, device=dev(), dtype=dtype, requires_grad=True)
    y = conv(x)
    gy = torch.randn_like(y)
    params = dict(conv.named_parameters())
    grads = torch.autograd.grad(y, [x] + list(params.values()), gy)
    leaves = {n: p.detach().double().cpu().requires_grad_(True) for n, p in params.items()}
    xo = x.detach().double().cpu().requires_grad_(True)
    fs = [leaves[f"weight.factors.factor_{i}"] for i in range(len(conv.weight.factors))]
    kw = dict(bias=leaves["bias"], stride=[2, 2], padding=[1, 1], dilation=[1, 1])
    if impl == "reconstructed":
        yo = rp.reconstructed_conv(xo, fact, (leaves["weight.weights"], fs), **kw)
    elif fact == "cp":
        yo = (rp.cp_conv if impl == "factorized" else rp.cp_conv_mobilenet)(xo, leaves["weight.weights"], fs, **kw)
    elif fact == "tucker":
        yo = rp.tucker_conv(xo, leaves["weight.core"], fs, **kw)
    else:
        yo = rp.tt_conv(xo, fs, **kw)
    go = torch.autograd.grad(yo, [xo] + list(leaves.values()), gy.double().cpu())
    tol = TOL[dtype]
    assert G.rel_err(y, yo) < tol
    for n, a, b in zip(["x"] + list(params), grads, go):
        assert G.rel_err(a, b) < tol, n


def test_trl_and_dropout_vs_oracle():
    import torch_b200 as tlt
    from torch_b200.functional.tensor_regression import trl
    from oracle import reference_path as rp
    torch.manual_seed(9)
    layer = tlt.TRL((32, 7, 7), (10,), bias=True, factorization="tucker", rank=(8, 4, 4, 5), device=dev())
    x = torch.randn(16, 32, 7, 7, device=dev(), requires_grad=True)
    y = layer(x)
    core = layer.weight.core.detach().double().cpu()
    fs = [f.detach().double().cpu() for f in layer.weight.factors]
    yo = rp.tucker_trl(x.detach().double().cpu(), core, fs, bias=layer.bias.detach().double().cpu())
    assert G.rel_err(y, yo) < 1e-5
    # dropout through the functional (TRL.forward bypasses hooks, App. C.1); same seed -> same kept components
    tlt.tensor_dropout(layer.weight, p=0.5, min_dim=3)
    layer.weight.train()
    torch.manual_seed(77)
    dropped = layer.weight()
    torch.manual_seed(77)
    idx = [rp.sample_keep_indices(r, 0.5, 3, device=dev()).cpu() for r in layer.weight.rank]
    c2, f2 = rp.tucker_dropout_apply(core, fs, idx, 0.5, training=True)
    assert dropped.rank == tuple(c2.shape)
    yd = trl(x, dropped, bias=layer.bias)
    yo = rp.tucker_trl(x.detach().double().cpu(), c2, f2, bias=layer.bias.detach().double().cpu())
    assert G.rel_err(yd, yo) < 1e-5


# ------------------------------------------------------------------------------------------------- (c) full size
def test_config2_full_size_properties():
    """BASELINE config 2 (BERT-base FFN, BlockTT rank 16, 64x512 tokens, bf16): linearity in x, agreement of the
    factorized path with x @ to_matrix().T on a row sample checked in fp64, and gradient identities."""
    import torch_b200 as tlt
    torch.manual_seed(2)
    up = tlt.FactorizedLinear((8, 8, 12), (8, 12, 32), bias=True, factorization="blocktt", rank=16, device=dev()).to(torch.bfloat16)
    B = 64 * 512
    x = torch.randn(B, 768, device=dev(), dtype=torch.bfloat16, requires_grad=True)
    y = up(x)
    assert y.shape == (B, 3072)
    W = up.weight.to_matrix().double().cpu()
    rows = torch.randint(0, B, (64,))
    want = x.detach()[rows.to(dev())].double().cpu() @ W.t() + up.bias.double().cpu()
    assert G.rel_err(y[rows.to(dev())], want) < 2e-2
    # linearity: f(2x) - b == 2 (f(x) - b)   (exact scaling by 2 in bf16)
    y2 = up(2 * x.detach())
    lhs = (y2.float() - up.bias.float())
    rhs = 2 * (y.detach().float() - up.bias.float())
    assert G.rel_err(lhs, rhs.double()) < 2e-2
    # gradient identities on a column sample: dX = dY W ; dbias = sum_b dY
    gy = torch.randn_like(y)
    gx, gb = torch.autograd.grad(y, [x, up.bias], gy)
    want_gx = gy[rows.to(dev())].double().cpu() @ W
    assert G.rel_err(gx[rows.to(dev())], want_gx) < 2e-2
    assert G.rel_err(gb, gy.double().sum(0).cpu()) < 2e-2


# ------------------------------------------------------------------------------------------------- mmd16 / Tucker linear
def _mmd_ref(x, mats, flags):
    res = x
    for k in range(3):
        if mats[k] is None:
            continue
        m = mats[k].t() if flags[k] else mats[k]
        res = torch.movedim(torch.tensordot(res, m, dims=([1 + k], [1])), -1, 1 + k)
    return res


@pytest.mark.parametrize("B,dims_in,dims_out,flags,with_bias,skip", [
    (301, (16, 16, 16), (15, 15, 15), (True, True, True), False, None),       # config 5 project
    (301, (15, 15, 15), (16, 16, 16), (False, False, False), True, None),     # config 5 expand + bias (staged loads)
    (130, (8, 12, 16), (7, 5, 16), (True, False, True), True, None),          # ragged outer modes, full innermost
    (67, (16, 9, 11), (12, 16, 13), (False, True, False), False, None),       # ragged innermost on both sides
    (5, (16, 16, 16), (16, 16, 16), (False, False, False), False, 1),         # fewer samples than warps; identity mode
    (1000, (4, 16, 16), (16, 4, 16), (True, True, False), True, None),
])
def test_mmd16_kernels_vs_fp64(B, dims_in, dims_out, flags, with_bias, skip):
    """tlt_mmd16_fwd / tlt_mmd16_bwd (HMMA mode-product chain, one warp per sample) against an fp64 chain."""
    from torch_b200 import tucker_ops as to
    torch.manual_seed(B)
    dims_in, dims_out = list(dims_in), list(dims_out)
    if skip is not None:
        dims_out[skip] = dims_in[skip]
    n_in, n_out = math.prod(dims_in), math.prod(dims_out)
    ld = (n_in + 7) & ~7
    xbuf = torch.randn(B, ld, device=dev(), dtype=torch.bfloat16)
    xbuf[:, n_in:] = float("nan")                                             # pad columns must never be consumed
    x = xbuf.detach().requires_grad_(True)
    mats = []
    for k in range(3):
        if skip == k:
            mats.append(None)
            continue
        shape = (dims_in[k], dims_out[k]) if flags[k] else (dims_out[k], dims_in[k])
        mats.append((torch.randn(shape, device=dev()) / math.sqrt(dims_in[k])).to(torch.bfloat16).requires_grad_(True))
    bias = torch.randn(n_out, device=dev(), dtype=torch.bfloat16, requires_grad=True) if with_bias else None
    y = to.mmd16(x, dims_in, dims_out, mats, list(flags), bias)
    assert y.shape == (B, (n_out + 7) & ~7)
    assert bool((y[:, n_out:] == 0).all())
    gy = torch.randn_like(y)
    leaves = [x] + [m for m in mats if m is not None] + ([bias] if with_bias else [])
    grads = torch.autograd.grad(y, leaves, gy)
    # fp64 reference on the same bf16 values
    xo = xbuf[:, :n_in].double().cpu().reshape(B, *dims_in).requires_grad_(True)
    mo = [None if m is None else m.detach().double().cpu().requires_grad_(True) for m in mats]
    yo = _mmd_ref(xo, mo, flags).reshape(B, n_out)
    bo = None
    if with_bias:
        bo = bias.detach().double().cpu().requires_grad_(True)
        yo = yo + bo
    lo = [xo] + [m for m in mo if m is not None] + ([bo] if with_bias else [])
    go = torch.autograd.grad(yo, lo, gy[:, :n_out].double().cpu())
    assert G.rel_err(y[:, :n_out], yo) < 1e-2
    assert G.rel_err(grads[0][:, :n_in], go[0].reshape(B, n_in)) < 1.5e-2
    for a, b in zip(grads[1:], go[1:]):
        assert G.rel_err(a, b) < 1.5e-2


def test_tucker3_fused_linear_vs_oracle():
    """FactorizedLinear (Tucker, 16^3 -> 16^3, ranks 15^6 like config 5 at a reduced batch) through the fused route
    (mmd16 -> tcgen05 GEMM at pitch 3376 -> mmd16) vs the fp64 oracle."""
    import torch_b200 as tlt
    from torch_b200 import tucker_ops as to
    from oracle import reference_path as rp
    torch.manual_seed(5)
    ts = (16, 16, 16)
    layer = tlt.FactorizedLinear(ts, ts, bias=True, factorization="tucker", rank=0.75, device=dev()).to(torch.bfloat16)
    assert tuple(layer.weight.core.shape) == (15,) * 6
    x = torch.randn(520, 4096, device=dev(), dtype=torch.bfloat16, requires_grad=True)
    assert to.tucker3_eligible(x, layer.weight.core, list(layer.weight.factors), layer.bias, list(ts), list(ts))
    y = layer(x)
    gy = torch.randn_like(y)
    params = dict(layer.named_parameters())
    grads = torch.autograd.grad(y, [x] + list(params.values()), gy)
    leaves = {n: p.detach().double().cpu().requires_grad_(True) for n, p in params.items()}
    xo = x.detach().double().cpu().requires_grad_(True)
    fs = [leaves[f"weight.factors.factor_{i}"] for i in range(6)]
    yo = rp.factorized_linear(xo, "tucker", (leaves["weight.core"], fs), (ts, ts), bias=leaves["bias"], order="sane")
    go = torch.autograd.grad(yo, [xo] + list(leaves.values()), gy.double().cpu())
    assert G.rel_err(y, yo) < 2e-2
    _report(_tag(), {"y": G.rel_err(y, yo), **{n: G.rel_err(a, b) for n, a, b in zip(["x"] + list(params), grads, go)}}, 2e-2)


# ------------------------------------------------------------------------------------------------- rows (channels-last) conv path
@pytest.mark.parametrize("B,C,spatial,kernel,stride,padding,dilation", [
    (3, 40, (7, 7), (3, 3), (1, 1), (1, 1), (1, 1)),
    (2, 70, (14, 9), (3, 3), (2, 2), (1, 1), (1, 1)),
    (2, 64, (9, 11), (3, 2), (1, 2), (2, 0), (2, 1)),
    (4, 33, (19,), (5,), (2,), (2,), (1,)),
    (2, 16, (5, 6, 7), (3, 3, 3), (1, 2, 1), (1, 1, 1), (1, 1, 1)),
])
def test_rows_layout_and_unfold_kernels_vs_torch(B, C, spatial, kernel, stride, padding, dilation):
    """tlt_nchw_to_rows / rows_to_nchw / im2col_rows / col2im_rows against permute + F.unfold-style references (exact)."""
    import torch.nn.functional as F
    from torch_b200 import rows_ops as ro
    from torch_b200 import ops
    torch.manual_seed(C)
    x = torch.randn(B, C, *spatial, device=dev(), dtype=torch.bfloat16)
    rows = ro.nchw_to_rows(x)
    ld = (C + 7) & ~7
    S = math.prod(spatial)
    assert rows.shape == (B * S, ld)
    want = x.reshape(B, C, S).permute(0, 2, 1).reshape(B * S, C)
    assert torch.equal(rows[:, :C], want) and bool((rows[:, C:] == 0).all())
    assert torch.equal(ro.rows_to_nchw(rows, B, C, list(spatial)), x)
    # unfold: reference = conv with one-hot taps on every channel, in fp64 (values are copies: exact)
    nd = len(spatial)
    fn = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[nd]
    taps = math.prod(kernel)
    out_size = ops.conv_out_size(list(spatial), list(kernel), list(stride), list(padding), list(dilation))
    So = math.prod(out_size)
    cols = ro.im2col_rows(rows, B, spatial, kernel, stride, padding, dilation)
    assert cols.shape == (B * So, taps * ld)
    eye = torch.eye(taps, dtype=torch.float64).reshape(taps, 1, *kernel)
    img = rows.double().cpu().reshape(B, *spatial, ld).movedim(-1, 1).reshape(B * ld, 1, *spatial)
    ref = fn(img, eye, None, stride, padding, dilation).reshape(B, ld, taps, So).permute(0, 3, 2, 1).reshape(B * So, taps * ld)
    assert torch.equal(cols.double().cpu(), ref)
    # fold = adjoint of unfold: compare with the autograd transpose of the fp64 reference unfold
    c = torch.randn_like(cols)
    back = ro._fold_op(c, B, ld, list(spatial), list(kernel), list(stride), list(padding), list(dilation))
    img = img.requires_grad_(True)
    ref = fn(img, eye, None, stride, padding, dilation).reshape(B, ld, taps, So).permute(0, 3, 2, 1).reshape(B * So, taps * ld)
    (gimg,) = torch.autograd.grad(ref, img, c.double().cpu())
    want_back = gimg.reshape(B, ld, S).permute(0, 2, 1).reshape(B * S, ld)
    assert G.rel_err(back, want_back) < 5e-3


@pytest.mark.parametrize("cin,cout,hw,stride,padding,dilation,rank", [
    (64, 96, 7, 1, 1, 1, 0.5), (48, 64, 9, 2, 1, 1, 0.8), (96, 64, 8, 1, 2, 2, 0.6),
])
def test_tucker_conv_rows_path_vs_oracle(cin, cout, hw, stride, padding, dilation, rank):
    """Tucker FactorizedConv (bf16) through the channels-last row-GEMM route vs the fp64 oracle."""
    import torch_b200 as tlt
    from torch_b200 import rows_ops as ro
    from oracle import reference_path as rp
    torch.manual_seed(cin)
    conv = tlt.FactorizedConv(cin, cout, 3, order=2, stride=stride, padding=padding, dilation=dilation, bias=True,
                              factorization="tucker", rank=rank, device=dev())
    conv.reset_parameters(std=0.3)
    conv.bias.data.normal_()
    conv = conv.to(torch.bfloat16)
    assert min(cin, cout, *conv.weight.core.shape[:2]) >= ro.MIN_CHANNELS
    x = torch.randn(5, cin, hw, hw + 1, device=dev(), dtype=torch.bfloat16, requires_grad=True)
    y = conv(x)
    gy = torch.randn_like(y)
    params = dict(conv.named_parameters())
    grads = torch.autograd.grad(y, [x] + list(params.values()), gy)
    leaves = {n: p.detach().double().cpu().requires_grad_(True) for n, p in params.items()}
    xo = x.detach().double().cpu().requires_grad_(True)
    fs = [leaves[f"weight.factors.factor_{i}"] for i in range(4)]
    yo = rp.tucker_conv(xo, leaves["weight.core"], fs, bias=leaves["bias"], stride=[stride] * 2, padding=[padding] * 2,
                        dilation=[dilation] * 2)
    go = torch.autograd.grad(yo, [xo] + list(leaves.values()), gy.double().cpu())
    assert G.rel_err(y, yo) < 2e-2
    _report(_tag(), {"y": G.rel_err(y, yo), **{n: G.rel_err(a, b) for n, a, b in zip(["x"] + list(params), grads, go)}}, 2e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,I,J", [(150544, 4, 9), (1000, 49, 16), (77, 16, 49), (129, 64, 64), (5, 1, 3), (4096, 27, 8)])
def test_skinny_kernels_vs_fp64(rows, I, J, dtype):
    """tlt_skinny_fwd / tlt_skinny_bwd (core expansion, spatial-first TRL projection) vs fp64 matmuls."""
    from torch_b200 import rows_ops as ro
    torch.manual_seed(rows + I)
    a = torch.randn(rows, I, device=dev(), dtype=dtype, requires_grad=True)
    k = (torch.randn(J, I, device=dev()) / math.sqrt(I)).to(dtype).requires_grad_(True)
    y = ro.skinny_linear(a, k)
    gy = torch.randn_like(y)
    ga, gk = torch.autograd.grad(y, [a, k], gy)
    ad, kd, gd = a.detach().double().cpu(), k.detach().double().cpu(), gy.double().cpu()
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert G.rel_err(y, ad @ kd.t()) < tol
    assert G.rel_err(ga, gd @ kd) < tol
    assert G.rel_err(gk, gd.t() @ ad) < tol


def test_trl_rows_projection_vs_oracle():
    """Tucker TRL on (B, 2048, 7, 7) maps (config 4 shape, reduced batch / classes): map -> channels-last rows, channel
    mode as one tcgen05 GEMM, spatial modes on the 32x smaller result, then the regression GEMM -- against the fp64 einsum
    of the reconstructed weight."""
    import torch_b200 as tlt
    torch.manual_seed(3)
    trl = tlt.TRL((2048, 7, 7), (200,), bias=True, factorization="tucker", rank=(64, 4, 4, 10), device=dev())
    with torch.no_grad():
        for p in trl.parameters():
            p.mul_(0.3)
    trl = trl.to(torch.bfloat16)
    x = torch.randn(48, 2048, 7, 7, device=dev(), dtype=torch.bfloat16, requires_grad=True)
    y = trl(x)
    gy = torch.randn_like(y)
    params = dict(trl.named_parameters())
    grads = torch.autograd.grad(y, [x] + list(params.values()), gy)
    leaves = {n: p.detach().double().cpu().requires_grad_(True) for n, p in params.items()}
    xo = x.detach().double().cpu().requires_grad_(True)
    fs = [leaves[f"weight.factors.factor_{i}"] for i in range(4)]
    W = torch.einsum("abcd,ia,jb,kc,ld->ijkl", leaves["weight.core"], *fs)
    yo = torch.einsum("bijk,ijkl->bl", xo, W) + leaves["bias"]
    go = torch.autograd.grad(yo, [xo] + list(leaves.values()), gy.double().cpu())
    assert G.rel_err(y, yo) < 2e-2
    _report(_tag(), {"y": G.rel_err(y, yo), **{n: G.rel_err(a, b) for n, a, b in zip(["x"] + list(params), grads, go)}}, 2e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_contract_small_output_long_reduction(dtype):
    """tlt_contract's tiny-output path (M, N <= 8, K >= 8192): factor gradients of small spatial matrices."""
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert _contract_vs_einsum("bpqr,bpir->qi", (96, 4, 4, 64), (96, 4, 7, 64), dtype) < tol
    assert _contract_vs_einsum("bqpr,bipr->qi", (64, 7, 5, 40), (64, 3, 5, 40), dtype) < tol
    assert _contract_vs_einsum("km,kn->mn", (20000, 8), (20000, 8), dtype) < tol
    assert _contract_vs_einsum("mk,kn->mn", (2, 9000), (9000, 3), dtype, alpha=0.5) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_scale_modes_and_fixed_shape_dropout(dtype):
    """tlt_scale_modes vs broadcasting in fp64, and Tucker tensor dropout in its fixed-shape (sync-free) form vs the
    reference's gather form on the same RNG draws, through a Tucker convolution."""
    import torch_b200 as tlt
    from torch_b200 import rows_ops as ro
    from torch_b200.tensor_hooks import _tensor_dropout as td
    torch.manual_seed(1)
    x = torch.randn(37, 5, 3, 9, device=dev(), dtype=dtype, requires_grad=True)
    m0 = (torch.rand(37, device=dev()) > 0.3).to(dtype)
    m3 = torch.randn(9, device=dev()).to(dtype)
    y = ro.scale_modes(x, [m0, None, None, m3], 1.7)
    want = 1.7 * x.detach().double().cpu() * m0.double().cpu()[:, None, None, None] * m3.double().cpu()
    tol = 1e-6 if dtype == torch.float32 else 1e-2
    assert G.rel_err(y, want) < tol
    (gx,) = torch.autograd.grad(y, x, torch.ones_like(y))
    assert G.rel_err(gx, want / x.detach().double().cpu()) < tol
    conv = tlt.FactorizedConv(64, 64, 3, order=2, padding=1, factorization="tucker", rank=0.7, bias=False, device=dev())
    conv.reset_parameters(std=0.3)
    conv = conv.to(dtype)
    assert min(conv.weight.rank[:2]) >= 40
    conv.weight = tlt.tensor_dropout(conv.weight, p=0.15, min_dim=3)
    conv.train()
    xin = torch.randn(3, 64, 8, 8, device=dev(), dtype=dtype)
    torch.manual_seed(11)
    y_fixed = conv(xin)
    saved = td._empty_draw_impossible
    td._empty_draw_impossible = lambda r, p: False
    try:
        torch.manual_seed(11)
        y_gather = conv(xin)
    finally:
        td._empty_draw_impossible = saved
    assert G.rel_err(y_fixed, y_gather.double().cpu()) < (1e-5 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize("B,C,spatial,kernel,stride,padding,dilation", [
    (4, 288, (14, 14), (3, 3), (1, 1), (1, 1), (1, 1)),
    (3, 72, (14, 14), (3, 3), (2, 2), (1, 1), (1, 1)),
    (2, 576, (7, 7), (3, 3), (1, 1), (1, 1), (1, 1)),
    (2, 40, (9, 11), (3, 2), (1, 2), (2, 0), (2, 1)),
    (3, 24, (19,), (5,), (2,), (2,), (1,)),
    (2, 16, (5, 6, 7), (2, 2, 2), (1, 2, 1), (1, 1, 0), (1, 1, 1)),
])
def test_dwconv_rows_kernels_vs_torch(B, C, spatial, kernel, stride, padding, dilation):
    """tlt_dwconv_rows_{fwd,dgrad,wgrad} (channels-last depthwise conv) vs F.conv{1,2,3}d(groups=C) in fp64."""
    import torch.nn.functional as F
    from torch_b200 import rows_ops as ro
    torch.manual_seed(C)
    nd = len(spatial)
    fn = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[nd]
    taps = math.prod(kernel)
    z = torch.randn(B * math.prod(spatial), C, device=dev(), dtype=torch.bfloat16, requires_grad=True)
    w = (torch.randn(taps, C, device=dev()) / math.sqrt(taps)).to(torch.bfloat16).requires_grad_(True)
    y = ro.dwconv_rows(z, w, B, spatial, kernel, stride, padding, dilation)
    gy = torch.randn_like(y)
    gz, gw = torch.autograd.grad(y, [z, w], gy)
    zo = z.detach().double().cpu().reshape(B, *spatial, C).movedim(-1, 1).requires_grad_(True)
    wo = w.detach().double().cpu().requires_grad_(True)
    yo = fn(zo, wo.t().reshape(C, 1, *kernel), None, stride, padding, dilation, C)
    go = gy.double().cpu().reshape(B, *yo.shape[2:], C).movedim(-1, 1)
    gzo, gwo = torch.autograd.grad(yo, [zo, wo], go)
    assert G.rel_err(y, yo.movedim(1, -1).reshape(-1, C)) < 1e-2
    assert G.rel_err(gz, gzo.movedim(1, -1).reshape(-1, C)) < 1e-2
    assert G.rel_err(gw, gwo) < 1e-2


@pytest.mark.parametrize("cin,cout,hw,stride", [(256, 256, 14, 1), (128, 256, 28, 2), (512, 512, 7, 1), (256, 512, 14, 2)])
def test_cp_conv_rows_path_vs_oracle(cin, cout, hw, stride):
    """CP FactorizedConv (bf16, ResNet-18 layer3 / layer4 shapes at a small batch) through the channels-last route vs the
    fp64 oracle."""
    import torch_b200 as tlt
    from oracle import reference_path as rp
    torch.manual_seed(hw)
    conv = tlt.FactorizedConv(cin, cout, 3, order=2, stride=stride, padding=1, bias=False, factorization="cp", rank=0.25, device=dev())
    conv.reset_parameters(std=0.5)
    conv = conv.to(torch.bfloat16)
    x = torch.randn(4, cin, hw, hw, device=dev(), dtype=torch.bfloat16, requires_grad=True)
    y = conv(x)
    gy = torch.randn_like(y)
    params = dict(conv.named_parameters())
    grads = torch.autograd.grad(y, [x] + list(params.values()), gy)
    leaves = {n: p.detach().double().cpu().requires_grad_(True) for n, p in params.items()}
    xo = x.detach().double().cpu().requires_grad_(True)
    fs = [leaves[f"weight.factors.factor_{i}"] for i in range(4)]
    yo = rp.cp_conv(xo, leaves["weight.weights"], fs, bias=None, stride=[stride] * 2, padding=[1, 1], dilation=[1, 1])
    go = torch.autograd.grad(yo, [xo] + list(leaves.values()), gy.double().cpu())
    assert G.rel_err(y, yo) < 2e-2
    _report(_tag(), {"y": G.rel_err(y, yo), **{n: G.rel_err(a, b) for n, a, b in zip(["x"] + list(params), grads, go)}}, 2e-2)


@pytest.mark.parametrize("B,C,spatial,ranks", [(3, 2048, (7, 7), (4, 4)), (5, 128, (7, 7), (4, 4)), (2, 256, (4, 4), (2, 2)),
                                               (2, 384, (8, 8), (4, 4)), (7, 128, (3, 3), (2, 3)), (2, 256, (5,), (3,)),
                                               (600, 128, (7, 7), (3, 5)), (2, 128, (2, 3, 4), (2, 2, 2))])
def test_spatial_project_kernels_vs_fp64(B, C, spatial, ranks):
    """tlt_spatial_project_fwd / _bwd (spatial modes of tucker_trl's input projection, functional/tensor_regression.py:40,
    on channels-first maps) vs an fp64 einsum: t, dx and dk."""
    from torch_b200 import spatial_ops
    torch.manual_seed(C + len(spatial))
    S, J = math.prod(spatial), math.prod(ranks)
    x = torch.randn(B, C, S, device=dev(), dtype=torch.bfloat16, requires_grad=True)
    k = (torch.randn(S, J, device=dev()) * 0.5).to(torch.bfloat16).requires_grad_(True)
    assert spatial_ops.eligible(x, S, J)
    t = spatial_ops.spatial_project(x, k)
    assert t.shape == (B, J, C)
    g = torch.randn_like(t)
    gx, gk = torch.autograd.grad(t, [x, k], g)
    xo, ko = x.detach().double().cpu().requires_grad_(True), k.detach().double().cpu().requires_grad_(True)
    to = torch.einsum("bcs,sj->bjc", xo, ko)
    gxo, gko = torch.autograd.grad(to, [xo, ko], g.double().cpu())
    assert G.rel_err(t, to) < 1e-2
    assert G.rel_err(gx, gxo) < 1e-2
    assert G.rel_err(gk, gko) < 1e-2


@pytest.mark.parametrize("B,C,hw,ranks,out", [(6, 256, 7, (64, 4, 4, 10), 100), (4, 512, 4, (32, 2, 3, 5), 17)])
def test_trl_spatial_first_plan_vs_rows_plan(B, C, hw, ranks, out, monkeypatch):
    """Tucker TRL on (B, C, hw, hw) maps: the spatial-modes-first plan (tlt_spatial_project + channel GEMM) and the
    channels-last-rows plan give the same output and gradients (both stand in for tensor_regression.py:37-49)."""
    import torch_b200 as tlt
    torch.manual_seed(C)
    layer = tlt.TRL((C, hw, hw), (out,), bias=True, rank=ranks, factorization="tucker", device=dev())
    with torch.no_grad():
        for p_ in layer.parameters():
            p_.normal_(0, 0.3)
    layer = layer.to(torch.bfloat16)
    x = torch.randn(B, C, hw, hw, device=dev(), dtype=torch.bfloat16, requires_grad=True)
    res = {}
    for mode in ("on", "off"):
        monkeypatch.setenv("TLT_SPATIAL_FIRST", mode)
        y = layer(x)
        gy = torch.ones_like(y) * 0.5
        res[mode] = [y] + list(torch.autograd.grad(y, [x] + list(layer.parameters()), gy))
    for a, b_ in zip(res["on"], res["off"]):
        assert G.rel_err(a, b_.double().cpu()) < 2e-2


@pytest.mark.parametrize("stride,h,w,r,batch", [(1, 28, 28, 141, 3), (1, 7, 7, 573, 5), (1, 14, 14, 285, 2), (2, 28, 28, 189, 3),
                                                (2, 14, 14, 381, 2), (2, 56, 56, 93, 2), (1, 5, 9, 13, 3), (1, 3, 2, 8, 1),
                                                # staged kernels (dwsep_staged.cu): many bands per block (both stages of the ring, several
                                                # mbarrier phases), ragged last band / last segment, one-segment rows, wide rows
                                                (1, 14, 14, 24, 160), (1, 7, 7, 64, 400), (1, 28, 28, 141, 48), (1, 9, 30, 40, 70),
                                                (1, 6, 56, 72, 40), (1, 11, 3, 16, 200),
                                                # stride-2 weight gradient through the staged kernel, many bands per block
                                                (2, 28, 28, 24, 60), (2, 14, 14, 64, 150), (2, 10, 6, 16, 100), (2, 56, 56, 93, 12)])
def test_dwsep_rows_kernels_vs_fp64(stride, h, w, r, batch):
    """tlt_dwsep_rows_fwd / _dgrad / _wgrad / _finalize (separable 3x3 depthwise sweep on channels-last rows, config 3's rank
    widths and map sizes, odd maps, a ragged channel chunk) vs fp64 autograd of the two depthwise 1-D convolutions of cp_conv
    (functional/convolution.py:251-269)."""
    from torch_b200 import rows_ops
    torch.manual_seed(h * w + r)
    ld = (r + 7) // 8 * 8
    z = torch.zeros(batch * h * w, ld, device=dev(), dtype=torch.bfloat16)
    z[:, :r] = torch.randn(batch * h * w, r, device=dev()).to(torch.bfloat16)
    z.requires_grad_(True)
    kh = (torch.randn(3, r, device=dev()) * 0.6).to(torch.bfloat16).requires_grad_(True)
    kw = (torch.randn(3, r, device=dev()) * 0.6).to(torch.bfloat16).requires_grad_(True)
    lam = (1 + 0.2 * torch.randn(r, device=dev())).to(torch.bfloat16).requires_grad_(True)
    y = rows_ops._dwsep_op(z, kh, kw, lam, batch, h, w, stride)
    gy = torch.zeros_like(y)
    gy[:, :r] = torch.randn(y.shape[0], r, device=dev()).to(torch.bfloat16)
    grads = torch.autograd.grad(y, [z, kh, kw, lam], gy)
    zo, kho, kwo, lo = (t.detach().double().cpu().requires_grad_(True) for t in (z, kh, kw, lam))
    img = zo[:, :r].reshape(batch, h, w, r).permute(0, 3, 1, 2)
    u = torch.nn.functional.conv2d(img, kho.t().reshape(r, 1, 3, 1), stride=(stride, 1), padding=(1, 0), groups=r)
    v = torch.nn.functional.conv2d(u, (kwo * lo).t().reshape(r, 1, 1, 3), stride=(1, stride), padding=(0, 1), groups=r)
    yo = v.permute(0, 2, 3, 1).reshape(-1, r)
    go = torch.autograd.grad(yo, [zo, kho, kwo, lo], gy[:, :r].double().cpu())
    assert float(y[:, r:].abs().max()) == 0.0 if ld > r else True
    errs = {"y": G.rel_err(y[:, :r], yo), "dz": G.rel_err(grads[0][:, :r], go[0][:, :r]), "dkh": G.rel_err(grads[1], go[1]),
            "dkw": G.rel_err(grads[2], go[2]), "dlam": G.rel_err(grads[3], go[3])}
    _report(_tag(), errs, 1e-2)


@pytest.mark.parametrize("cin,cout,stride,height,rank,batch,bias,train", [
    (64, 64, 1, 56, 0.25, 3, False, True),      # config 3 layer1: rank 69, second-generation kernels (pass folded into the MMA)
    (64, 64, 1, 56, 0.25, 7, True, True),       # ranges that cross image boundaries, bias
    (64, 64, 1, 56, 0.4, 2, False, True),       # rank 110 > 80: first-generation kernels (TMEM-resident z1 with halo)
    (64, 64, 1, 24, 0.25, 5, True, True),       # H != 56: first-generation kernels
    (64, 128, 2, 56, 0.25, 5, False, False),    # config 3 layer2[0] (stride 2, rank 93): fused forward
    (64, 128, 2, 56, 0.25, 3, True, False),
])
def test_cp_conv_rows_fused_vs_oracle(cin, cout, stride, height, rank, batch, bias, train):
    """CP FactorizedConv (bf16, channels-last 56-wide maps) through the op-level fused kernels (tlt_cpconv_rows_fwd / _bwd:
    1x1 -> depthwise 3x3 -> 1x1 in one launch per direction, rank intermediates in TMEM / shared memory, every gradient
    accumulated on chip) vs the fp64 oracle of cp_conv (functional/convolution.py:231-283): y, dx and every parameter gradient."""
    import torch_b200 as tlt
    from torch_b200 import _cuda
    from oracle import reference_path as rp
    torch.manual_seed(height + batch)
    conv = tlt.FactorizedConv(cin, cout, 3, order=2, stride=stride, padding=1, bias=bias, factorization="cp", rank=rank, device=dev())
    conv.reset_parameters(std=0.5)
    if bias:
        with torch.no_grad():
            conv.bias.normal_()
    conv = conv.to(torch.bfloat16)
    x = torch.randn(batch, cin, height, 56, device=dev(), dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    x.requires_grad_(train)
    params = dict(conv.named_parameters())
    leaves = {n: p.detach().double().cpu().requires_grad_(True) for n, p in params.items()}
    xo = x.detach().double().cpu().requires_grad_(True)
    fs = [leaves[f"weight.factors.factor_{i}"] for i in range(4)]
    yo = rp.cp_conv(xo, leaves["weight.weights"], fs, bias=leaves.get("bias"), stride=[stride] * 2, padding=[1, 1], dilation=[1, 1])
    before = _cuda.launch_count()
    with torch.set_grad_enabled(train):
        y = conv(x)
    assert _cuda.launch_count() - before == 2, "operand packing + the fused chain, nothing else"
    assert y.is_contiguous(memory_format=torch.channels_last)
    errs = {"y": G.rel_err(y, yo)}
    if train:
        gy = torch.randn_like(y)
        before = _cuda.launch_count()
        grads = torch.autograd.grad(y, [x] + list(params.values()), gy)
        assert _cuda.launch_count() - before <= 3 + (1 if bias else 0)      # data + weight gradient kernels, finalisation
        go = torch.autograd.grad(yo, [xo] + list(leaves.values()), gy.double().cpu())
        errs.update({n: G.rel_err(a, b) for n, a, b in zip(["x"] + list(params), grads, go)})
    _report(_tag(), errs, 2e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("dims,R,with_w", [((4, 8, 16), 2341, True), ((4, 8, 16), 2341, False), ((3, 3), 69, True), ((7,), 33, True),
                                           ((2, 3, 4, 5), 17, True)])
def test_khatri_rao_kernels_vs_fp64(dims, R, with_w, dtype):
    """tlt_khatri_rao_fwd / _bwd (weighted Khatri-Rao product; CP linear's operands at config 1's rank) vs fp64 einsum."""
    from torch_b200 import kr_ops
    torch.manual_seed(R)
    fs = [(torch.randn(d, R, device=dev()) * 0.7).to(dtype).requires_grad_(True) for d in dims]
    w = torch.rand(R, device=dev()).add(0.5).to(dtype).requires_grad_(True) if with_w else None
    kr = kr_ops.khatri_rao(fs, w)
    g = torch.randn_like(kr)
    leaves = fs + ([w] if with_w else [])
    grads = torch.autograd.grad(kr, leaves, g)
    fo = [f.detach().double().cpu().requires_grad_(True) for f in fs]
    wo = w.detach().double().cpu().requires_grad_(True) if with_w else None
    L = "abcd"[:len(dims)]
    ref = torch.einsum(",".join(l + "r" for l in L) + "->" + L + "r", *fo)
    if with_w:
        ref = ref * wo
    ref = ref.reshape(-1, R)
    go = torch.autograd.grad(ref, fo + ([wo] if with_w else []), g.double().cpu())
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert G.rel_err(kr, ref) < tol
    for a, b in zip(grads, go):
        assert G.rel_err(a, b) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("dims,R,with_w,pad", [((4, 8, 16), 2341, True, 3), ((4, 8, 16), 2341, False, 0), ((5,), 130, True, 0), ((3, 20), 77, True, 1),
                                               ((2, 3, 4, 5), 17, True, 0), ((6, 40), 65, True, 0), ((7, 3, 32), 200, False, 8)])
def test_khatri_rao_all_gradients_one_launch_vs_fp64(dims, R, with_w, pad, dtype):
    """tlt_khatri_rao_bwd_all: every factor gradient and the weight gradient of one Khatri-Rao product in one launch, g at a
    leading dimension >= R -- the one-pass kernel (last dimension <= 16 / <= 32: register tiles of 16 / 32) and the per-factor kernel
    behind it (last dimension 40), against fp64 autograd of the einsum (factorized_tensordot.py:70-103 evaluates exactly this product)."""
    from torch_b200 import cplinear_tc_ops as ct
    torch.manual_seed(R + len(dims))
    fs = [(torch.randn(d, R, device=dev()) * 0.7).to(dtype) for d in dims]
    w = torch.rand(R, device=dev()).add(0.5).to(dtype) if with_w else None
    rows = 1
    for d in dims:
        rows *= d
    ldg = R + pad
    g = torch.randn(rows, ldg, device=dev()).to(dtype)
    dfs = [torch.full((d, R), 7.0, device=dev()) for d in dims]          # the call zeroes them itself
    dw = torch.zeros(R, device=dev()) if with_w else None
    ct._execute_kr_bwd_ld(fs, w, g, ldg, dfs, dw)
    fo = [f.double().cpu().requires_grad_(True) for f in fs]
    wo = w.double().cpu().requires_grad_(True) if with_w else None
    L = "abcd"[:len(dims)]
    ref = torch.einsum(",".join(l + "r" for l in L) + "->" + L + "r", *fo)
    if with_w:
        ref = ref * wo
    go = torch.autograd.grad(ref.reshape(rows, R), fo + ([wo] if with_w else []), g[:, :R].double().cpu())
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    errs = {f"dF{k}": G.rel_err(a, b) for k, (a, b) in enumerate(zip(dfs, go))}
    if with_w:
        errs["dw"] = G.rel_err(dw, go[-1])
    _report(_tag(), errs, tol)


# ------------------------------------------------------------------------------------------------- CP / TT tensor dropout on the GPU
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("kind,shape,rank,proba", [
    ("cp", (10, 11, 12), 9, 0.4),                      # gather form (an empty draw is possible: reference's index path)
    ("cp", (24, 16, 3, 3), 70, 0.12),                  # fixed-shape form (tlt_scale_modes on lambda), conv-kernel shaped
    ("tt", (10, 11, 12), (1, 6, 3, 1), 0.4),           # gather form
    ("tt", (9, 40, 40), (1, 36, 37, 1), 0.12),         # fixed-shape form (masks multiplied into the cores)
    ("tucker", (10, 11, 12), (7, 6, 3), 0.4),
])
def test_tensor_dropout_vs_oracle_on_gpu(kind, shape, rank, proba, training, dtype):
    """CPDropout / TTDropout / TuckerDropout through the real kernels against the oracle's restatement of
    tltorch/tensor_hooks/_tensor_dropout.py:56-142, the hook's RNG draws replayed on the same device generator: the dropped
    tensor's reconstruction AND the gradients reaching the un-dropped parameters through it."""
    import torch_b200 as tlt
    from torch_b200.tensor_hooks import _tensor_dropout as td
    from oracle import reference_path as rp
    torch.manual_seed(3)
    w = tlt.FactorizedTensor.new(shape, rank=rank, factorization=kind, device=dev())
    w.normal_(0, 0.5)
    if kind == "cp":
        w.weights.data.uniform_(0.5, 1.5)
    w = w.to(dtype)
    hook = {"tucker": td.TuckerDropout, "cp": td.CPDropout, "tt": td.TTDropout}[kind](proba, min_dim=3, min_values=1, drop_test=True)
    ranks = w.rank if kind != "cp" else (w.rank,)
    torch.manual_seed(1234)
    if kind == "tt":
        idx = [rp.sample_keep_indices(r, proba, 3, device=dev()).cpu() for r in ranks[1:]]
    else:
        idx = [rp.sample_keep_indices(r, proba, 3, device=dev()).cpu() for r in ranks]
    torch.manual_seed(1234)
    before = tlt.ops.launch_count()
    dropped = hook._apply_tensor_dropout(w, training=training)
    full = dropped.to_tensor()
    g = torch.randn_like(full)
    params = dict(w.named_parameters())
    grads = torch.autograd.grad(full, list(params.values()), g)
    assert tlt.ops.launch_count() > before
    leaves = {n: p.detach().double().cpu().requires_grad_(True) for n, p in params.items()}
    fs = [leaves[f"factors.factor_{i}"] for i in range(len(w.factors))]
    if kind == "cp":
        w2, f2 = rp.cp_dropout_apply(leaves["weights"], fs, idx[0], proba, training=training)
        fo = rp.cp_to_tensor(w2, f2)
    elif kind == "tucker":
        c2, f2 = rp.tucker_dropout_apply(leaves["core"], fs, idx, proba, training=training)
        fo = rp.tucker_to_tensor(c2, f2)
    else:
        fo = rp.tt_to_tensor(rp.tt_dropout_apply(fs, idx, proba, training=training))
    go = torch.autograd.grad(fo, list(leaves.values()), g.double().cpu())
    errs = {"full": G.rel_err(full, fo)}
    for n, a, b in zip(params, grads, go):
        errs[n] = G.rel_err(a, b)
    _report(f"dropout {kind} rank={rank} train={training} {dtype}", errs, TOL[dtype])


@pytest.mark.parametrize("in_shape,out_shape,rank,batch,bias", [
    ((4, 8, 16), (4, 8, 16), 2341, 128, True),     # BASELINE config 1
    ((4, 8, 16), (4, 8, 16), 0.5, 300, False),     # several batch tiles (128-row tiles of the projection, 32-row tiles of the expansion)
    ((2, 6), (3, 4), 9, 37, True),                 # everything ragged: 12 features, rank tile of 16 not full, rank % 4 = 1
    ((12,), (2, 2, 5, 1), 50, 3, True),            # one mode in, four out
    ((16, 16), (8, 64), 130, 64, True),            # rank % 4 = 2, 512 output features
])
def test_cp_linear_fused_vs_oracle(in_shape, out_shape, rank, batch, bias, monkeypatch):
    """fp32 CP FactorizedLinear through the fused op (tlt_cplinear_fwd / _bwd, csrc/cplinear.cu: 2 launches forward, 4 backward, no
    Khatri-Rao matrix materialised) vs the fp64 oracle of linear_cp -> tensor_dot_cp (functional/factorized_linear.py:23-36,
    factorized_tensordot.py:70-103): y, dx and every parameter gradient at north_star's fp32 tolerance 1e-5."""
    import math
    import torch_b200 as tlt
    from torch_b200 import _cuda
    from oracle import reference_path as rp
    monkeypatch.setenv("TLT_CPLINEAR_FUSED", "1")          # opt-in path (slower than the unfused chain at config 1's size)
    monkeypatch.setenv("TLT_CPLINEAR_TC", "0")             # (the tensor-core route would take the layer first)
    torch.manual_seed(batch)
    lin = tlt.FactorizedLinear(in_shape, out_shape, bias=bias, factorization="cp", rank=rank, implementation="factorized", device=dev(),
                               dtype=torch.float32)
    with torch.no_grad():
        for p in lin.parameters():
            p.normal_(0, 0.4)
    x = torch.randn(batch, math.prod(in_shape), device=dev(), requires_grad=True)
    gy = torch.randn(batch, math.prod(out_shape), device=dev())
    n0 = _cuda.lib().tlt_launch_count()
    y = lin(x)
    names = [n for n, _ in lin.named_parameters()]
    grads = torch.autograd.grad(y, [x] + list(lin.parameters()), gy)
    assert _cuda.lib().tlt_launch_count() - n0 == 6, "the fused op is 2 + 4 launches"
    leaves = {n: p.detach().double().cpu().requires_grad_(True) for n, p in lin.named_parameters()}
    xo = x.detach().double().cpu().requires_grad_(True)
    fs = [leaves[f"weight.factors.factor_{i}"] for i in range(len(in_shape) + len(out_shape))]
    yo = rp.factorized_linear(xo, "cp", (leaves["weight.weights"], fs), (out_shape, in_shape), bias=leaves.get("bias"), order="sane")
    go = torch.autograd.grad(yo, [xo] + [leaves[n] for n in names], gy.double().cpu())
    errs = {"y": G.rel_err(y, yo)}
    for a, b_, n in zip(grads, go, ["x"] + names):
        errs["d" + n] = G.rel_err(a, b_)
    _report(_tag(), errs, 1e-5)


@pytest.mark.parametrize("B,J,rc,ldz,ro,O,bias", [(512, 16, 64, 64, 10, 1000, True), (37, 12, 20, 24, 5, 30, True), (5, 1, 3, 8, 16, 7, False),
                                                  (300, 6, 33, 40, 1, 513, True)])
def test_trl_tail_op_vs_fp64(B, J, rc, ldz, ro, O, bias):
    """tlt_trl_tail_fwd / _bwd (core contraction + output expansion + bias of tucker_trl, tensor_regression.py:42-49, one launch forward,
    two backward) vs fp64 autograd: config 4's shape, ragged everything, the largest output rank, a batch that crosses the 256-row
    staging chunk of the parameter-gradient kernel."""
    from torch_b200 import trl_ops
    torch.manual_seed(B + O)
    zp = torch.zeros(B, J, ldz, device=dev(), dtype=torch.bfloat16)
    zp[:, :, :rc] = (torch.randn(B, J, rc, device=dev()) * 0.5).to(torch.bfloat16)
    zp.requires_grad_(True)
    core = (torch.randn(rc, J, ro, device=dev()) * 0.3).to(torch.bfloat16).requires_grad_(True)
    fo = (torch.randn(O, ro, device=dev()) * 0.5).to(torch.bfloat16).requires_grad_(True)
    bv = torch.randn(O, device=dev()).to(torch.bfloat16).requires_grad_(True) if bias else None
    y = trl_ops.trl_tail(zp, core, fo, bv)
    gy = torch.randn_like(y)
    leaves = [zp, core, fo] + ([bv] if bias else [])
    grads = torch.autograd.grad(y, leaves, gy)
    lo = [t.detach().double().cpu().requires_grad_(True) for t in leaves]
    vo = torch.einsum("bjc,cjr->br", lo[0][:, :, :rc], lo[1])
    yo = vo @ lo[2].t() + (lo[3] if bias else 0)
    go = torch.autograd.grad(yo, lo, gy.double().cpu())
    errs = {"y": G.rel_err(y, yo), "dzp": G.rel_err(grads[0][:, :, :rc], go[0][:, :, :rc]), "dcore": G.rel_err(grads[1], go[1]),
            "dfo": G.rel_err(grads[2], go[2])}
    if bias:
        errs["dbias"] = G.rel_err(grads[3], go[3])
    assert float(grads[0][:, :, rc:].abs().max()) == 0.0 if ldz > rc else True
    _report(_tag(), errs, 1e-2)


@pytest.mark.parametrize("in_shape,out_shape,rank,batch,bias", [
    ((4, 8, 16), (4, 8, 16), 2341, 128, True),     # BASELINE config 1
    ((4, 8, 16), (4, 8, 16), 0.5, 300, False),     # batch not a multiple of any tile
    ((2, 8), (4, 4), 67, 16, True),                # smallest eligible layer: 16 features, rank % 8 = 3
    ((16, 16), (8, 64), 130, 64, True),
])
def test_cp_linear_tensor_core_vs_oracle(in_shape, out_shape, rank, batch, bias):
    """fp32 CP FactorizedLinear with every product on tcgen05 through 3-way bf16-split operands (cplinear_tc_ops, csrc/split3.cu) vs the
    fp64 oracle of linear_cp -> tensor_dot_cp (functional/factorized_linear.py:23-36, factorized_tensordot.py:70-103): y, dx and every
    parameter gradient at north_star's fp32 tolerance 1e-5 (a plain bf16 product would sit at ~3e-3)."""
    import math
    import torch_b200 as tlt
    from oracle import reference_path as rp
    torch.manual_seed(batch)
    lin = tlt.FactorizedLinear(in_shape, out_shape, bias=bias, factorization="cp", rank=rank, implementation="factorized", device=dev(),
                               dtype=torch.float32)
    with torch.no_grad():
        for p in lin.parameters():
            p.normal_(0, 0.4)
    x = torch.randn(batch, math.prod(in_shape), device=dev(), requires_grad=True)
    gy = torch.randn(batch, math.prod(out_shape), device=dev())
    from torch_b200 import cplinear_tc_ops
    assert cplinear_tc_ops.eligible(x, lin.weight.weights, list(lin.weight.factors)[:len(out_shape)], list(lin.weight.factors)[len(out_shape):],
                                    lin.bias)
    y = lin(x)
    names = [n for n, _ in lin.named_parameters()]
    grads = torch.autograd.grad(y, [x] + list(lin.parameters()), gy)
    leaves = {n: p.detach().double().cpu().requires_grad_(True) for n, p in lin.named_parameters()}
    xo = x.detach().double().cpu().requires_grad_(True)
    fs = [leaves[f"weight.factors.factor_{i}"] for i in range(len(in_shape) + len(out_shape))]
    yo = rp.factorized_linear(xo, "cp", (leaves["weight.weights"], fs), (out_shape, in_shape), bias=leaves.get("bias"), order="sane")
    go = torch.autograd.grad(yo, [xo] + [leaves[n] for n in names], gy.double().cpu())
    errs = {"y": G.rel_err(y, yo)}
    for a, b_, n in zip(grads, go, ["x"] + names):
        errs["d" + n] = G.rel_err(a, b_)
    _report(_tag(), errs, 1e-5)


def _ws(nbytes):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev())


@pytest.mark.parametrize("transpose,bias", [(True, True), (False, False)])
def test_op_level_blocktt3_linear_matches_python_op(transpose, bias):
    """tlt_blocktt3_linear_fwd / _bwd (one C call per direction, caller-owned workspaces: the entry points a non-Python host binds,
    SURVEY 8(b)) give the same y / dx / core gradients / bias gradient as the torch_b200 op on BASELINE config 2's `up` layer shape
    (reduced batch), i.e. as linear_blocktt (functional/factorized_linear.py:39-60) within the bf16 tolerance."""
    import ctypes as C
    from torch_b200 import _cuda, blocktt_ops
    torch.manual_seed(3)
    rows, cols, r = (8, 12, 32), (8, 8, 12), 16
    c1 = (torch.randn(1, rows[0], cols[0], r, device=dev()) * 0.3).to(torch.bfloat16).requires_grad_(True)
    c2 = (torch.randn(r, rows[1], cols[1], r, device=dev()) * 0.3).to(torch.bfloat16).requires_grad_(True)
    c3 = (torch.randn(r, rows[2], cols[2], 1, device=dev()) * 0.3).to(torch.bfloat16).requires_grad_(True)
    R, Cn = 8 * 12 * 32, 8 * 8 * 12
    fin, fout = (Cn, R) if transpose else (R, Cn)
    B = 384
    x = torch.randn(B, fin, device=dev(), dtype=torch.bfloat16, requires_grad=True)
    bv = torch.randn(fout, device=dev(), dtype=torch.bfloat16, requires_grad=True) if bias else None
    y_ref = blocktt_ops.blocktt3_linear(x, c1, c2, c3, bv, transpose)
    gy = torch.randn_like(y_ref)
    g_ref = torch.autograd.grad(y_ref, [x, c1, c2, c3] + ([bv] if bias else []), gy)
    d = _cuda.BlockTT3LinearDesc()
    d.tt.dtype = _cuda.dtype_code(torch.bfloat16)
    for k in range(3):
        d.tt.rows[k], d.tt.cols[k] = rows[k], cols[k]
    d.tt.rank1 = d.tt.rank2 = r
    d.batch, d.transpose, d.dtype_bias = B, int(transpose), _cuda.dtype_code(torch.bfloat16)
    lib = _cuda.lib()
    ws, gws = _ws(lib.tlt_blocktt3_linear_workspace_size(C.byref(d))), _ws(lib.tlt_blocktt3_linear_grad_workspace_size(C.byref(d)))
    y = torch.empty_like(y_ref)
    st = _cuda.stream_ptr(x)
    _cuda.check(lib.tlt_blocktt3_linear_fwd(C.byref(d), _cuda.ptr(x), _cuda.ptr(c1), _cuda.ptr(c2), _cuda.ptr(c3), _cuda.ptr(bv), _cuda.ptr(ws),
                                            _cuda.ptr(y), st), "fwd")
    dx, d1, d2, d3 = torch.empty_like(x), torch.empty_like(c1), torch.empty_like(c2), torch.empty_like(c3)
    db = torch.empty(fout, device=dev(), dtype=torch.bfloat16) if bias else None
    _cuda.check(lib.tlt_blocktt3_linear_bwd(C.byref(d), _cuda.ptr(x), _cuda.ptr(gy), _cuda.ptr(ws), _cuda.ptr(c1), _cuda.ptr(c2), _cuda.ptr(c3),
                                            _cuda.ptr(gws), _cuda.ptr(dx), _cuda.ptr(d1), _cuda.ptr(d2), _cuda.ptr(d3), _cuda.ptr(db), st), "bwd")
    torch.cuda.synchronize()
    errs = {"y": G.rel_err(y, y_ref.double().cpu()), "dx": G.rel_err(dx, g_ref[0].double().cpu()), "d1": G.rel_err(d1, g_ref[1].double().cpu()),
            "d2": G.rel_err(d2, g_ref[2].double().cpu()), "d3": G.rel_err(d3, g_ref[3].double().cpu())}
    if bias:
        errs["dbias"] = G.rel_err(db, g_ref[4].double().cpu())
    _report(_tag(), errs, 1e-2)


@pytest.mark.parametrize("spatial,sranks,side", [((7, 7), (4, 4), True), ((7, 7), (4, 3), False), ((4, 3, 2), (2, 2, 2), True), ((49,), (12,), False)])
def test_op_level_tucker_trl_matches_python_op(spatial, sranks, side):
    """tlt_tucker_trl_fwd / _bwd (a whole Tucker TRL per C call, optional second stream for the parameter-gradient branch) vs the
    torch_b200 layer (tucker_trl, functional/tensor_regression.py:37-49): y, dx and every parameter gradient; 1, 2 and 3 spatial modes."""
    import ctypes as C
    import torch_b200 as tlt
    from torch_b200 import _cuda
    torch.manual_seed(len(spatial))
    B, Cc, rc, ro, O = 64, 256, 24, 5, 40
    layer = tlt.TRL((Cc,) + tuple(spatial), (O,), bias=True, factorization="tucker", rank=(rc,) + tuple(sranks) + (ro,), device=dev())
    with torch.no_grad():
        for p in layer.parameters():
            p.normal_(0, 0.3)
    layer = layer.to(torch.bfloat16)
    x = torch.randn(B, Cc, *spatial, device=dev(), dtype=torch.bfloat16, requires_grad=True)
    y_ref = layer(x)
    gy = torch.randn_like(y_ref)
    names = [n for n, _ in layer.named_parameters()]
    g_ref = dict(zip(["x"] + names, torch.autograd.grad(y_ref, [x] + list(layer.parameters()), gy)))
    core, factors, bv = layer.weight.core.detach(), [f.detach().contiguous() for f in layer.weight.factors], layer.bias.detach()
    d = _cuda.TuckerTrlDesc()
    d.batch, d.channels, d.n_spatial, d.rc, d.ro, d.out = B, Cc, len(spatial), rc, ro, O
    for k in range(len(spatial)):
        d.spatial[k], d.spatial_rank[k] = spatial[k], sranks[k]
    lib = _cuda.lib()
    nws = lib.tlt_tucker_trl_workspace_size(C.byref(d))
    assert nws > 0, "shape refused by the op-level entry point"
    ws, gws = _ws(nws), _ws(lib.tlt_tucker_trl_grad_workspace_size(C.byref(d)))
    fptr = (C.c_void_p * len(factors))(*[f.data_ptr() for f in factors])
    y = torch.empty_like(y_ref)
    st = _cuda.stream_ptr(x)
    _cuda.check(lib.tlt_tucker_trl_fwd(C.byref(d), _cuda.ptr(x), _cuda.ptr(core), fptr, _cuda.ptr(bv), _cuda.ptr(ws), _cuda.ptr(y), st), "fwd")
    dx, dcore, db = torch.empty_like(x), torch.empty_like(core), torch.empty_like(bv)
    dfs = [torch.empty_like(f) for f in factors]
    dptr = (C.c_void_p * len(dfs))(*[f.data_ptr() for f in dfs])
    s2 = torch.cuda.Stream() if side else None
    _cuda.check(lib.tlt_tucker_trl_bwd(C.byref(d), _cuda.ptr(x), _cuda.ptr(gy), _cuda.ptr(core), fptr, _cuda.ptr(ws), _cuda.ptr(gws), _cuda.ptr(dx),
                                       _cuda.ptr(dcore), dptr, _cuda.ptr(db), st, C.c_void_p(s2.cuda_stream) if side else None), "bwd")
    torch.cuda.synchronize()
    errs = {"y": G.rel_err(y, y_ref.double().cpu()), "dx": G.rel_err(dx, g_ref["x"].double().cpu()),
            "dcore": G.rel_err(dcore, g_ref["weight.core"].double().cpu()), "dbias": G.rel_err(db, g_ref["bias"].double().cpu())}
    for k, f in enumerate(dfs):
        errs[f"dfactor_{k}"] = G.rel_err(f, g_ref[f"weight.factors.factor_{k}"].double().cpu())
    _report(_tag(), errs, 1e-2)
