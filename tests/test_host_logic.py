"""CPU tests of the product's HOST logic with the CUDA kernels replaced by descriptor interpreters
(tests/_emulator.py): every functional path and its autograd must reproduce the real reference's golden outputs."""
import pytest
import torch

import torch_b200 as tlt
from torch_b200 import ops, tenalg
import _golden as G
import _cases
from _emulator import emulated_kernels

TOL = 1e-10
D = torch.float64


@pytest.mark.parametrize("name,sub", _cases.all_cases())
def test_functional_paths_match_reference_goldens(name, sub):
    g, sub2 = _cases.load_case(name, sub)
    with emulated_kernels() as log:
        y, grads, ey, eg = _cases.run_case(name, g, "cpu", D, sub2)
    assert log, "no kernel launch was recorded: the path bypassed the custom ops"
    assert G.rel_err(y, ey.reshape(y.shape)) < TOL
    assert set(grads) == set(eg)
    for n, gr in grads.items():
        assert gr is not None, n
        assert G.rel_err(gr, eg[n]) < TOL, n


def test_reconstruction_matches_reference():
    allg = G.load("reconstruction")
    with emulated_kernels():
        for key, g in allg.items():
            w, _ = _cases.build_weight(g["kind"], g["params"], "cpu", D, g.get("tensorized_shape"))
            full = w.to_matrix() if key.startswith("matrix_") else w.to_tensor()
            assert tuple(full.shape) == tuple(g["full"].shape), key
            assert G.rel_err(full, g["full"]) < TOL, key


def test_dropout_replays_reference_rng():
    from torch_b200.tensor_hooks._tensor_dropout import TuckerDropout, CPDropout, TTDropout
    allg = G.load("dropout")
    with emulated_kernels():
        for key, g in allg.items():
            kind = key.split("_")[0]
            w, _ = _cases.build_weight(kind, g["params"], "cpu", D)
            hook = {"tucker": TuckerDropout, "cp": CPDropout, "tt": TTDropout}[kind](g["proba"], min_dim=g["min_dim"],
                                                                                    min_values=1, drop_test=True)
            torch.manual_seed(g["seed"])
            dropped = hook._apply_tensor_dropout(w, training=g["training"])
            assert G.rel_err(dropped.to_tensor(), g["full"]) < TOL, key
            exp = g["dropped"]
            for a, b in zip(dropped.factors, exp["factors"]):
                assert a.shape == b.shape and G.rel_err(a, b) < 1e-14, key


def test_dropout_fixed_shape_form_equals_gather_form():
    """Large ranks (p ** rank < 1e-30) take the sync-free fixed-shape form: same RNG draws, same reconstructed tensor as
    the reference's gather form (reference: tensor_hooks/_tensor_dropout.py:56-142)."""
    import torch_b200 as tlt
    from torch_b200.tensor_hooks import _tensor_dropout as td
    with emulated_kernels():
        for kind, shape, rank in (("tucker", (40, 36, 3, 3), (34, 33, 2, 2)), ("cp", (12, 10, 9), 70), ("tt", (9, 40, 40), (1, 36, 37, 1))):
            torch.manual_seed(0)
            w = tlt.FactorizedTensor.new(shape, rank=rank, factorization=kind, dtype=D)
            w.normal_(0, 0.5)
            hook = {"tucker": td.TuckerDropout, "cp": td.CPDropout, "tt": td.TTDropout}[kind](0.12, min_dim=3, min_values=1)
            for training in (True, False):
                hook.drop_test = True
                torch.manual_seed(7)
                fixed = hook._apply_tensor_dropout(w, training=training)
                assert [tuple(f.shape) for f in fixed.factors] == [tuple(f.shape) for f in w.factors]      # shapes unchanged
                saved = td._empty_draw_impossible
                td._empty_draw_impossible = lambda r, p: False                                               # force the gather form
                try:
                    torch.manual_seed(7)
                    gathered = hook._apply_tensor_dropout(w, training=training)
                finally:
                    td._empty_draw_impossible = saved
                assert any(a.shape != b.shape for a, b in zip(gathered.factors, w.factors))                  # something was dropped
                assert G.rel_err(fixed.to_tensor(), gathered.to_tensor()) < TOL, (kind, training)


def test_dropout_hook_plumbing():
    # reference tests: tensor_hooks/tests/test_tensor_dropout.py:8-61 (shapes only)
    with emulated_kernels():
        for fact in ("cp", "tucker", "tt"):
            t = tlt.FactorizedTensor.new((10, 11, 12), rank=8 if fact != "tt" else (1, 8, 8, 1), factorization=fact).normal_()
            t = tlt.tensor_dropout(t, p=0.999, min_dim=1)
            t.train()
            r = t().rank
            r = (r,) if isinstance(r, int) else tuple(r)
            assert all(v == 1 for v in (r[1:-1] if fact == "tt" else r)), (fact, r)
            tlt.remove_tensor_dropout(t)
            assert not t._forward_hooks
            t = tlt.tensor_dropout(t, p=0)
            assert t().rank == t.rank
    with pytest.raises(KeyError):          # App. C.2: no dropout class for tensorized matrices
        tlt.tensor_dropout(tlt.TensorizedTensor.new(((2, 2), (2, 2)), rank=2, factorization="cp"), p=0.5)


def test_layers_match_dense_equivalents():
    """The reference's own identities: FactorizedLinear == nn.Linear(to_matrix) (factorized_layers/tests/
    test_factorized_linear.py:10-31), FactorizedConv == nn.Conv2d(to_tensor) (test_factorized_convolution.py:13-57)."""
    torch.manual_seed(0)
    with emulated_kernels():
        for fact in ("cp", "tucker", "blocktt"):
            for impl in ("factorized", "reconstructed"):
                layer = tlt.FactorizedLinear((3, 3), (4, 4), bias=True, factorization=fact, rank=0.5 if fact != "blocktt" else 2,
                                             implementation=impl, dtype=D)
                x = torch.randn(2, 9, dtype=D)
                want = torch.nn.functional.linear(x, layer.weight.to_matrix(), layer.bias)
                assert G.rel_err(layer(x), want) < TOL, (fact, impl)
        for fact, impls in (("cp", ("factorized", "reconstructed", "mobilenet")), ("tucker", ("factorized", "reconstructed")),
                            ("tt", ("factorized", "reconstructed"))):
            for impl in impls:
                conv = tlt.FactorizedConv(4, 5, 3, order=2, padding=1, bias=True, factorization=fact, rank=0.5,
                                          implementation=impl, dtype=D)
                conv.reset_parameters(std=0.5)
                conv.bias.data.normal_()
                x = torch.randn(1, 4, 8, 7, dtype=D)
                from torch_b200.factorized_layers.factorized_convolution import tensor_to_kernel
                k = tensor_to_kernel(fact, conv.weight.to_tensor())
                want = torch.nn.functional.conv2d(x, k, conv.bias, padding=1)
                assert G.rel_err(conv(x), want) < TOL, (fact, impl)
        # n_layers > 1: jointly factorized linears / convs index the leading mode in factorized form
        for fact in ("cp", "tucker", "blocktt"):
            layer = tlt.FactorizedLinear((3, 3), (4, 4), bias=True, factorization=fact, rank=3, n_layers=2, dtype=D)
            x = torch.randn(2, 9, dtype=D)
            full = layer.weight.to_matrix()
            for i in range(2):
                want = torch.nn.functional.linear(x, full[i], layer.bias[i])
                assert G.rel_err(layer(x, i), want) < TOL, fact
                assert G.rel_err(layer[i](x), want) < TOL, fact
        # TRL / TCL modules
        trl = tlt.TRL((3, 4), (2,), bias=True, factorization="tucker", rank=(2, 3, 2), dtype=D)
        x = torch.randn(5, 3, 4, dtype=D)
        want = torch.einsum("bij,ijo->bo", x, trl.weight.to_tensor()) + trl.bias
        assert G.rel_err(trl(x), want) < TOL
        tcl = tlt.TCL((4, 5, 6), (2, 3, 5), bias=True, dtype=D)
        x = torch.randn(2, 4, 5, 6, dtype=D)
        assert tuple(tcl(x).shape) == (2, 2, 3, 5)     # reference test pins the shape only
        want = torch.einsum("bijk,ri,sj,tk->brst", x, *tcl.factors) + tcl.bias
        assert G.rel_err(tcl(x), want) < TOL


def test_state_dict_keys_match_reference_names():
    lin = tlt.FactorizedLinear((2, 2), (2, 2), factorization="cp", rank=2)
    assert set(lin.state_dict()) == {"bias", "weight.weights", "weight.factors.factor_0", "weight.factors.factor_1",
                                     "weight.factors.factor_2", "weight.factors.factor_3"}
    conv = tlt.FactorizedConv(3, 4, 3, order=2, factorization="tucker", rank=0.5)
    assert {"weight.core", "weight.factors.factor_0"} <= set(conv.state_dict())
    tt = tlt.FactorizedLinear((2, 2), (2, 2), factorization="blocktt", rank=2, bias=False)
    assert set(tt.state_dict()) == {"weight.factors.factor_0", "weight.factors.factor_1"}


def test_planner_merges_and_classifies():
    # contiguous GEMM: everything merges to one sub-dim per class and is tensor-core eligible
    p = ops.plan_contract("bi,oi->bo", (64, 32), (32, 1), (16, 32), (32, 1), None, None)
    assert p.classes == dict(b=[], m=[64], n=[16], k=[32]) and p.tc is not None
    assert p.tc["a_major"] == 0 and p.tc["b_major"] == 0
    # dW = dY^T X: both operands MN-major
    p = ops.plan_contract("bo,bi->oi", (64, 16), (16, 1), (64, 32), (32, 1), None, None)
    assert p.tc is not None and p.tc["a_major"] == 1 and p.tc["b_major"] == 1 and p.tc["lda"] == 16 and not p.swap
    # same product written operand-swapped (what autograd emits for dW): planner swaps roles back -> row-major C
    p = ops.plan_contract("bi,bo->oi", (64, 32), (32, 1), (64, 16), (16, 1), None, None)
    assert p.swap and p.tc is not None and p.tc["M"] == 16 and p.tc["N"] == 32 and p.tc["ldc"] == 32
    # TT sweep stage: composite m and k
    p = ops.plan_contract("abcr,rodq->acodq"[:0] + "abcr,rodq->acodq".replace("d", "b"), (5, 4, 3, 2), (24, 6, 2, 1),
                          (2, 6, 4, 7), (168, 28, 7, 1), None, None)
    assert p.classes["k"] and p.tc is None
    # size-1 dims disappear, batch index detected
    p = ops.plan_contract("zar,zbr->zab", (3, 4, 1), (4, 1, 1), (3, 5, 1), (5, 1, 1), None, None)
    assert p.classes["b"] == [3] and p.classes["k"] == []
    with pytest.raises(ValueError):
        ops.plan_contract("ab,bc->c", (2, 3), (3, 1), (3, 4), (4, 1), None, None)   # 'a' summed alone


def test_blocktt_cost_model_picks_dense_for_bert_ffn():
    from torch_b200.functional.factorized_linear import blocktt_plan, blocktt_sweep_macs
    assert blocktt_sweep_macs((8, 8, 12), (8, 12, 32), (1, 16, 16, 1)) == 3047424          # 3.05 M MAC
    assert blocktt_plan(32768, (8, 8, 12), (8, 12, 32), (1, 16, 16, 1)) == "dense"
    assert blocktt_plan(32768, (8, 12, 32), (8, 8, 12), (1, 16, 16, 1)) == "dense"
    assert blocktt_plan(32768, (16, 16, 16), (16, 16, 16), (1, 2, 2, 1)) == "sweep"


def test_cpu_tensors_are_rejected_without_fallback():
    lin = tlt.FactorizedLinear((2, 2), (2, 2), factorization="cp", rank=2)
    with pytest.raises(NotImplementedError):
        lin(torch.randn(3, 4))
    with pytest.raises(AssertionError):
        tlt.functional.factorized_linear(torch.randn(3, 4), lin.weight, implementation="nope")
    with pytest.raises(ValueError):
        tlt.FactorizedTensor.new((2, 2), rank=1, factorization="nope")


def test_tensor_core_routes_host_logic_bf16():
    """bf16 inputs take the tcgen05 routes (fused BlockTT linear, NCHW pointwise); with the launch funnels emulated on the
    CPU the op plumbing (operand majors, transposed weight views, gradient shapes) must reproduce fp64 autograd."""
    from torch_b200.functional.convolution import cp_conv, tucker_conv
    torch.manual_seed(5)
    bf = torch.bfloat16
    with emulated_kernels() as log:
        # CP conv 16 -> 24 channels on 8x8 maps (S = 64, a multiple of 8 -> pointwise_tc eligible once work is large enough)
        from torch_b200 import pointwise_ops
        old = pointwise_ops.MIN_WORK
        pointwise_ops.MIN_WORK = 1
        try:
            w = tlt.FactorizedTensor.new((24, 16, 3, 3), rank=5, factorization="cp").normal_(0, 0.3).to(bf)
            x = torch.randn(3, 16, 8, 8).to(bf).requires_grad_(True)
            bias = torch.randn(24).to(bf).requires_grad_(True)
            y = cp_conv(x, w, bias=bias, stride=1, padding=1)
            leaves = [x, bias, w.weights] + list(w.factors)
            gy = torch.randn_like(y)
            got = torch.autograd.grad(y, leaves, gy)
            assert any(e[0] == "pointwise_tc" for e in log) and any(e[0] == "pointwise_wgrad_tc" for e in log)
        finally:
            pointwise_ops.MIN_WORK = old
    # fp64 reference with torch's own conv on the reconstructed kernel
    ref_leaves = [t.detach().double().requires_grad_(True) for t in leaves]
    xr, br, lam = ref_leaves[0], ref_leaves[1], ref_leaves[2]
    f = ref_leaves[3:]
    kern = torch.einsum("r,or,cr,hr,wr->ochw", lam, f[0], f[1], f[2], f[3])
    yr = torch.nn.functional.conv2d(xr, kern, br, stride=1, padding=1)
    want = torch.autograd.grad(yr, ref_leaves, gy.double())
    assert G.rel_err(y, yr) < 2e-2
    for a, b in zip(got, want):
        assert G.rel_err(a, b) < 2e-2


def test_fused_multi_mode_dot_host_logic():
    """tenalg.multi_mode_dot -> ONE fused launch when the touched modes trail a contiguous tensor: plan, transposed data
    gradient, per-matrix gradients (fused op with one mode skipped + long reduction), fused bias; checked against fp64
    autograd with the launch funnels emulated."""
    from torch_b200 import modedot_ops
    torch.manual_seed(7)
    old = modedot_ops.MIN_BATCH, modedot_ops.MIN_SAMPLE_ELEMS
    modedot_ops.MIN_BATCH, modedot_ops.MIN_SAMPLE_ELEMS = 1, 1
    try:
        for shape, ranks, modes, transpose, with_bias in [((5, 6, 4, 3), (2, 5, 4), [1, 2, 3], False, True),
                                                           ((7, 3, 6, 4), (5, 2), [2, 3], True, False),
                                                           ((4, 2, 3, 5, 6), (4, 3), [2, 4], False, False),
                                                           ((9, 8), (3,), [1], True, True)]:
            x = torch.randn(shape, dtype=D, requires_grad=True)
            mats = []
            for r, m in zip(ranks, modes):
                mats.append(torch.randn((shape[m], r) if transpose else (r, shape[m]), dtype=D, requires_grad=True))
            out_tail = list(shape[min(modes):])
            for r, m in zip(ranks, modes):
                out_tail[m - min(modes)] = r
            bias = torch.randn(out_tail, dtype=D, requires_grad=True) if with_bias else None
            with emulated_kernels() as log:
                y = tenalg.multi_mode_dot(x, mats, modes, transpose=transpose, addend=bias)
                gy = torch.randn_like(y)
                leaves = [x] + mats + ([bias] if with_bias else [])
                got = torch.autograd.grad(y, leaves, gy)
            assert sum(e[0] == "multi_mode_dot" for e in log) >= 2, log
            ref = x
            for m, mode in zip(mats, modes):
                mm = m.t() if transpose else m
                ref = torch.movedim(torch.tensordot(ref, mm, dims=([mode], [1])), -1, mode)
            if with_bias:
                ref = ref + bias
            want = torch.autograd.grad(ref, leaves, gy)
            assert G.rel_err(y, ref) < TOL
            for a, b in zip(got, want):
                assert G.rel_err(a, b) < TOL
    finally:
        modedot_ops.MIN_BATCH, modedot_ops.MIN_SAMPLE_ELEMS = old


def test_spatial_first_projection_host_logic(monkeypatch):
    """tenalg.multi_mode_dot on a (B, C, h, w) map with a large channel mode -> spatial modes first (tltorch::spatial_project with
    the Kronecker product of the spatial factors, then the channel mode as one padded GEMM): plan selection, output layout and the
    autograd of x and of every factor, with the launch funnels emulated (tucker_trl's input projection, tensor_regression.py:40)."""
    from torch_b200 import spatial_ops
    torch.manual_seed(11)
    monkeypatch.setattr(spatial_ops, "eligible", lambda x, S, J: x.shape[1] % 128 == 0 and S <= 64 and J <= 16)
    B, C, h, w = 24, 256, 7, 7
    ranks = (32, 4, 3)
    x = (torch.randn(B, C, h, w) * 0.5).to(torch.bfloat16).requires_grad_(True)
    mats = [(torch.randn(d, r) * 0.3).to(torch.bfloat16).requires_grad_(True) for d, r in zip((C, h, w), ranks)]
    with emulated_kernels() as log:
        y = tenalg.multi_mode_dot(x, mats, [1, 2, 3], transpose=True)
        gy = torch.randn_like(y)
        got = torch.autograd.grad(y, [x] + mats, gy)
    assert any(e[0] == "spatial_project_fwd" for e in log) and any(e[0] == "spatial_project_bwd" for e in log), log
    assert tuple(y.shape) == (B, *ranks)
    xr = x.detach().double().requires_grad_(True)
    mr = [m.detach().double().requires_grad_(True) for m in mats]
    ref = torch.einsum("bchw,cp,hq,wr->bpqr", xr, *mr)
    want = torch.autograd.grad(ref, [xr] + mr, gy.double())
    assert G.rel_err(y, ref) < 2e-2
    for a, b in zip(got, want):
        assert G.rel_err(a, b) < 2e-2


def test_cp_conv_rows_fused_op_host_logic():
    """tltorch::cpconv_rows_fwd / _bwd (the op-level fused CP convolution on channels-last rows): routing from FactorizedConv
    (channels-last 56 x 56 map, 64 -> 64, CP rank 0.25 -> the fused op, two launches forward, two backward, no Khatri-Rao and no
    transposition launch), autograd plumbing and the bias gradient, against fp64 autograd of 1x1 -> depthwise -> 1x1 (cp_conv,
    convolution.py:231-283), with the launch funnels emulated."""
    torch.manual_seed(12)
    conv = tlt.FactorizedConv(64, 64, 3, order=2, padding=1, bias=True, factorization="cp", rank=0.25)
    conv.reset_parameters(std=0.4)
    conv.bias.data.normal_()
    conv = conv.to(torch.bfloat16)
    x = (torch.randn(2, 64, 56, 56) * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    params = list(conv.parameters())
    with emulated_kernels() as log:
        y = conv(x)
        gy = torch.randn_like(y)
        got = torch.autograd.grad(y, [x] + params, gy)
        names = [e[0] for e in log]
    assert names == ["cpconv_rows_pack", "cpconv_rows_fwd", "cpconv_rows_bwd", "cpconv_rows_finalize"], names
    assert y.shape == (2, 64, 56, 56) and y.is_contiguous(memory_format=torch.channels_last)
    xr = x.detach().double().requires_grad_(True)
    pr = {n: p.detach().double().requires_grad_(True) for n, p in conv.named_parameters()}
    fs = [pr[f"weight.factors.factor_{i}"] for i in range(4)]
    w = torch.einsum("r,or,cr,hr,wr->ochw", pr["weight.weights"], *fs)
    ref = torch.nn.functional.conv2d(xr, w, pr["bias"], padding=1)
    want = torch.autograd.grad(ref, [xr] + [pr[n] for n, _ in conv.named_parameters()], gy.double())
    assert G.rel_err(y, ref) < 2e-2
    for (n, _), g, wv in zip([("x", None)] + list(conv.named_parameters()), got, want):
        assert G.rel_err(g, wv) < 2e-2, n
    # an NCHW-contiguous input, another map size or stride-2 training keep the unfused routes
    with emulated_kernels() as log:
        conv(x.detach().contiguous())
        assert "cpconv_rows_fwd" not in [e[0] for e in log]


@pytest.mark.parametrize("stride,hw", [(1, (7, 7)), (2, (14, 14)), (1, (6, 10))])
def test_cp_conv_rows_separable_route_host_logic(stride, hw):
    """Channels-last CP convolution outside the fused shapes: GEMM -> tltorch::dwsep_rows (both depthwise passes in one sweep,
    the CP taps unmerged: no Khatri-Rao launch forward or backward) -> GEMM, against fp64 autograd of the dense convolution with
    the reconstructed kernel (cp_conv, convolution.py:231-283); launch funnels emulated."""
    torch.manual_seed(5 + stride)
    conv = tlt.FactorizedConv(64, 72, 3, order=2, padding=1, stride=stride, bias=True, factorization="cp", rank=0.3)
    conv.reset_parameters(std=0.4)
    conv.bias.data.normal_()
    conv = conv.to(torch.bfloat16)
    x = (torch.randn(2, 64, *hw) * 0.5).to(torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    with emulated_kernels() as log:
        y = conv(x)
        gy = torch.randn_like(y)
        got = torch.autograd.grad(y, [x] + list(conv.parameters()), gy)
        names = [e[0] for e in log]
    assert names == ["cpchain_prep", "gemm", "dwsep_rows_fwd", "gemm", "gemm", "dwsep_rows_dgrad", "gemm", "dwsep_rows_wgrad", "gemm",
                     "gemm", "cpchain_post"] + (["contract"] if "contract" in names else []), names
    xr = x.detach().double().requires_grad_(True)
    pr = {n: p.detach().double().requires_grad_(True) for n, p in conv.named_parameters()}
    w = torch.einsum("r,or,cr,hr,wr->ochw", pr["weight.weights"], *[pr[f"weight.factors.factor_{i}"] for i in range(4)])
    ref = torch.nn.functional.conv2d(xr, w, pr["bias"], stride=stride, padding=1)
    want = torch.autograd.grad(ref, [xr] + [pr[n] for n, _ in conv.named_parameters()], gy.double())
    assert G.rel_err(y, ref) < 2e-2
    for (n, _), g, wv in zip([("x", None)] + list(conv.named_parameters()), got, want):
        assert G.rel_err(g, wv) < 2e-2, n


@pytest.mark.parametrize("name", _cases.module_cases())
def test_module_cases_match_reference_goldens(name):
    """n_layers > 1 FactorizedLinear / FactorizedConv and FactorizedEmbedding, built from the REAL reference's state_dict
    (same key names), against its outputs and gradients -- kernels emulated."""
    g = G.load(name)
    with emulated_kernels() as log:
        y, grads, ey, eg = _cases.run_module_case(name, g, "cpu", D)
    assert log, "no kernel launch was recorded: the path bypassed the custom ops"
    assert G.rel_err(y, ey.reshape(y.shape)) < TOL
    assert set(grads) == set(eg)
    for n, gr in grads.items():
        assert gr is not None, n
        assert G.rel_err(gr, eg[n]) < TOL, n


def test_tltorch_import_alias_is_the_product():
    """north_star: the classes "work as drop-ins" -- ``import tltorch`` resolves to this implementation, submodule by submodule."""
    import tltorch
    import tltorch.functional.factorized_linear as fl
    from tltorch.factorized_tensors import CPTensor, BlockTT
    from tltorch.tensor_hooks import tensor_dropout
    import torch_b200
    assert tltorch.FactorizedLinear is torch_b200.FactorizedLinear and tltorch.TRL is torch_b200.TRL
    assert fl is torch_b200.functional.factorized_linear and CPTensor is torch_b200.CPTensor and BlockTT is torch_b200.BlockTT
    assert tensor_dropout is torch_b200.tensor_dropout
    assert tltorch.__file__.startswith(_cases.__file__.rsplit("/tests/", 1)[0])


def test_conv_transduct_starts_as_framewise_conv():
    """FactorizedConv.transduct (reference: factorized_convolution.py:346-393): a CP conv transducted 2-D -> 3-D with the
    centre-delta factor applies the old 2-D conv to every frame."""
    torch.manual_seed(0)
    with emulated_kernels():
        conv = tlt.FactorizedConv(4, 5, 3, order=2, padding=1, bias=True, factorization="cp", rank=0.5, dtype=D)
        conv.reset_parameters(std=0.5)
        conv.bias.data.normal_()
        x = torch.randn(2, 4, 3, 6, 5, dtype=D)
        y2 = torch.stack([conv(x[:, :, t]) for t in range(3)], dim=2)
        conv.transduct(3, mode=0, padding=1)
        assert conv.order == 3 and tuple(conv.weight.shape[2:]) == (3, 3, 3)
        y3 = conv(x)
    assert G.rel_err(y3, y2) < TOL


def test_channels_last_input_takes_the_rows_route_without_transposes():
    """A torch.channels_last activation is the (B H W, C) rows matrix already: the CP / Tucker conv rows routes view it (no
    nchw_to_rows / rows_to_nchw launch), return a channels_last output, and agree with the NCHW-contiguous result."""
    torch.manual_seed(3)
    for fact, rank in (("cp", 0.25), ("tucker", 0.8)):
        conv = tlt.FactorizedConv(64, 72, 3, order=2, padding=1, stride=2, bias=True, factorization=fact, rank=rank)
        conv.reset_parameters(std=0.3)
        conv.bias.data.normal_()
        conv = conv.to(torch.bfloat16)
        x = torch.randn(2, 64, 9, 10).to(torch.bfloat16)
        xcl = x.contiguous(memory_format=torch.channels_last).requires_grad_(True)
        xn = x.clone().requires_grad_(True)
        with emulated_kernels() as log_n:
            yn = conv(xn)
            gn = torch.autograd.grad(yn, [xn] + list(conv.parameters()), torch.ones_like(yn))
            log_n = list(log_n)
        with emulated_kernels() as log_c:
            yc = conv(xcl)
            gc_ = torch.autograd.grad(yc, [xcl] + list(conv.parameters()), torch.ones_like(yc))
        assert any(e[0] in ("nchw_to_rows", "rows_to_nchw") for e in log_n)
        assert not any(e[0] in ("nchw_to_rows", "rows_to_nchw") for e in log_c), [e[0] for e in log_c]
        assert yc.is_contiguous(memory_format=torch.channels_last) and yc.shape == yn.shape
        assert G.rel_err(yc, yn.double()) < 1e-2
        for a, b in zip(gc_, gn):
            assert G.rel_err(a, b.double()) < 2e-2


@pytest.mark.parametrize("in_shape,out_shape,rank,batch,bias", [((4, 8, 16), (4, 8, 16), 37, 5, True), ((2, 6), (3, 4), 9, 3, False),
                                                              ((12,), (2, 2, 5, 1), 6, 4, True)])
def test_cp_linear_fused_op_host_logic(in_shape, out_shape, rank, batch, bias, monkeypatch):
    """fp32 CP FactorizedLinear takes the fused op (cplinear_ops: 2 + 4 launches); descriptor dims / factor order / autograd plumbing
    checked on the emulator against the oracle of linear_cp (factorized_linear.py:23-36)."""
    from oracle import reference_path as rp
    import math
    monkeypatch.setenv("TLT_CPLINEAR_FUSED", "1")
    monkeypatch.setenv("TLT_CPLINEAR_TC", "0")
    torch.manual_seed(rank)
    lin = tlt.FactorizedLinear(in_shape, out_shape, bias=bias, factorization="cp", rank=rank, implementation="factorized", dtype=torch.float32)
    with torch.no_grad():
        for p in lin.parameters():
            p.normal_(0, 0.5)
    x = torch.randn(batch, math.prod(in_shape), requires_grad=True)
    gy = torch.randn(batch, math.prod(out_shape))
    with emulated_kernels() as log:
        y = lin(x)
        names = [n for n, _ in lin.named_parameters()]
        grads = torch.autograd.grad(y, [x] + list(lin.parameters()), gy)
    assert [e[0] for e in log] == ["cplinear_fwd", "cplinear_bwd"]
    leaves = {n: p.detach().double().requires_grad_(True) for n, p in lin.named_parameters()}
    xo = x.detach().double().requires_grad_(True)
    fs = [leaves[f"weight.factors.factor_{i}"] for i in range(len(in_shape) + len(out_shape))]
    yo = rp.factorized_linear(xo, "cp", (leaves["weight.weights"], fs), (out_shape, in_shape), bias=leaves.get("bias"), order="sane")
    go = torch.autograd.grad(yo, [xo] + [leaves[n] for n in names], gy.double())
    assert G.rel_err(y, yo) < 1e-5
    for a, b_, n in zip(grads, go, ["x"] + names):
        assert G.rel_err(a, b_) < 1e-5, n


def test_tucker_trl_fused_route_host_logic(monkeypatch):
    """bf16 Tucker TRL with one output mode on a (B, C, h, w) map -> spatial_project, channel GEMM, tltorch::trl_tail (core
    contraction + output expansion + bias as one op): routing, descriptor, autograd of x and of every parameter against the fp64
    einsum of tucker_trl (tensor_regression.py:37-49), launch funnels emulated."""
    from torch_b200 import spatial_ops
    torch.manual_seed(21)
    monkeypatch.setattr(spatial_ops, "eligible", lambda x, S, J: x.shape[1] % 128 == 0 and S <= 64 and J <= 16)
    B, C, h, w, O = 24, 256, 7, 7, 30
    trl_layer = tlt.TRL((C, h, w), (O,), bias=True, factorization="tucker", rank=(20, 4, 3, 5))
    with torch.no_grad():
        for p in trl_layer.parameters():
            p.normal_(0, 0.3)
    trl_layer = trl_layer.to(torch.bfloat16)
    x = (torch.randn(B, C, h, w) * 0.5).to(torch.bfloat16).requires_grad_(True)
    params = dict(trl_layer.named_parameters())
    with emulated_kernels() as log:
        y = trl_layer(x)
        gy = torch.randn_like(y)
        got = torch.autograd.grad(y, [x] + list(params.values()), gy)
    kinds = [e[0] for e in log]
    assert "trl_tail_fwd" in kinds and "trl_tail_bwd" in kinds and "spatial_project_fwd" in kinds, kinds
    leaves = {n: p.detach().double().requires_grad_(True) for n, p in params.items()}
    xo = x.detach().double().requires_grad_(True)
    fs = [leaves[f"weight.factors.factor_{i}"] for i in range(4)]
    Wt = torch.einsum("abcd,ia,jb,kc,ld->ijkl", leaves["weight.core"], *fs)
    yo = torch.einsum("bijk,ijkl->bl", xo, Wt) + leaves["bias"]
    want = torch.autograd.grad(yo, [xo] + list(leaves.values()), gy.double())
    assert G.rel_err(y, yo) < 2e-2
    for n, a, b in zip(["x"] + list(params), got, want):
        assert G.rel_err(a, b) < 2e-2, n


@pytest.mark.parametrize("in_shape,out_shape,rank,batch,bias", [((4, 8, 16), (4, 8, 16), 70, 24, True), ((2, 8), (4, 4), 67, 16, False)])
def test_cp_linear_tc_route_host_logic(in_shape, out_shape, rank, batch, bias, monkeypatch):
    """fp32 CP FactorizedLinear on the tensor cores (cplinear_tc_ops: Khatri-Rao operands, 3-way bf16 split in the six-block patterns,
    one GEMM per product, Khatri-Rao backward at a padded leading dimension): operand layouts, paddings and autograd plumbing on
    the emulator against the oracle of linear_cp (factorized_linear.py:23-36).  The emulated split keeps the real bf16 pieces, so
    the 1e-5 agreement also checks that the six blocks pair up as hi.hi + hi.mid + mid.hi + hi.lo + mid.mid + lo.hi."""
    from oracle import reference_path as rp
    import math
    torch.manual_seed(rank)
    lin = tlt.FactorizedLinear(in_shape, out_shape, bias=bias, factorization="cp", rank=rank, implementation="factorized", dtype=torch.float32)
    with torch.no_grad():
        for p in lin.parameters():
            p.normal_(0, 0.5)
    x = torch.randn(batch, math.prod(in_shape), requires_grad=True)
    gy = torch.randn(batch, math.prod(out_shape))
    with emulated_kernels() as log:
        y = lin(x)
        names = [n for n, _ in lin.named_parameters()]
        grads = torch.autograd.grad(y, [x] + list(lin.parameters()), gy)
    kinds = [e[0] for e in log]
    assert kinds.count("gemm") == 6 and "split3" in kinds, kinds
    leaves = {n: p.detach().double().requires_grad_(True) for n, p in lin.named_parameters()}
    xo = x.detach().double().requires_grad_(True)
    fs = [leaves[f"weight.factors.factor_{i}"] for i in range(len(in_shape) + len(out_shape))]
    yo = rp.factorized_linear(xo, "cp", (leaves["weight.weights"], fs), (out_shape, in_shape), bias=leaves.get("bias"), order="sane")
    go = torch.autograd.grad(yo, [xo] + [leaves[n] for n in names], gy.double())
    assert G.rel_err(y, yo) < 1e-5
    for a, b_, n in zip(grads, go, ["x"] + names):
        assert G.rel_err(a, b_) < 1e-5, n
