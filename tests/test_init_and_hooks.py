"""init-time decompositions, TensorLasso and the Complex* tensors.  Each test restates the check the
reference's own test makes (file:line in the docstring).  CPU runs swap the kernel funnels for the descriptor emulators
(tests/_emulator.py); the GPU variants call the real kernels."""
import numpy as np
import pytest
import torch
from torch import nn

import torch_b200 as tlt
from torch_b200 import decomposition
from torch_b200.factorized_layers.factorized_convolution import tensor_to_kernel
from _emulator import emulated_kernels


def _rel(a, b):
    return float((a - b).norm() / b.norm())


# ------------------------------------------------------------------------------------------------------------------- decompositions
def test_parafac_recovers_low_rank_tensor():
    """A rank-3 tensor is recovered by CP-ALS to solver precision (tensorly's own parafac test strategy)."""
    torch.manual_seed(0)
    factors = [torch.randn(s, 3, dtype=torch.float64) for s in (6, 7, 8)]
    full = torch.einsum("ar,br,cr->abc", *factors)
    w, f = decomposition.parafac(full, 3, n_iter_max=500, tol=1e-14)
    rec = torch.einsum("r,ar,br,cr->abc", w, *f)
    assert _rel(rec, full) < 1e-6


def test_tucker_and_tt_exact_at_full_rank():
    torch.manual_seed(1)
    full = torch.randn(4, 5, 6, dtype=torch.float64)
    core, f = decomposition.tucker(full, (4, 5, 6))
    assert _rel(torch.einsum("xyz,ax,by,cz->abc", core, *f), full) < 1e-10
    for fac in f:                                            # HOOI factors are orthonormal
        assert torch.allclose(fac.t() @ fac, torch.eye(fac.shape[1], dtype=torch.float64), atol=1e-10)
    cores = decomposition.tensor_train(full, (1, 4, 6, 1))
    assert _rel(torch.einsum("xay,ybz,zcw->abc", *cores), full) < 1e-10
    # truncated: the TT-SVD error is bounded by the discarded singular values (quasi-optimality)
    cores = decomposition.tensor_train(full, (1, 3, 3, 1))
    assert [tuple(c.shape) for c in cores] == [(1, 4, 3), (3, 5, 3), (3, 6, 1)]


def test_tensor_train_matrix_layout():
    """Cores come back as (r_k, rows_k, cols_k, r_{k+1}) and multiply out to the tensorized matrix (tensorized_matrices.py:517-524)."""
    torch.manual_seed(2)
    full = torch.randn(2, 3, 4, 3, 2, 2, dtype=torch.float64)
    cores = decomposition.tensor_train_matrix(full, (1, 6, 8, 1))
    assert [tuple(c.shape[1:3]) for c in cores] == [(2, 3), (3, 2), (4, 2)]
    rec = torch.einsum("wadx,xbey,ycfz->abcdef", *cores)
    assert _rel(rec, full) < 1e-10


@pytest.mark.parametrize("factorization", ["CP", "Tucker", "BlockTT"])
def test_from_linear_decomposes_the_weight(factorization):
    """reference: factorized_layers/tests/test_factorized_linear.py:31-44 (rank 34 on a 9 -> 16 layer, outputs equal to 2 decimals)."""
    torch.manual_seed(3)
    data = torch.randn(4, 9, dtype=torch.float64)
    fc = nn.Linear(9, 16, bias=True).double()
    with emulated_kernels():
        tfc = tlt.FactorizedLinear.from_linear(fc, auto_tensorize=False, in_tensorized_features=(3, 3), out_tensorized_features=(4, 4),
                                               rank=34, bias=True, factorization=factorization)
        np.testing.assert_array_almost_equal(fc(data).detach(), tfc(data).detach(), decimal=2)
        tfc = tlt.FactorizedLinear.from_linear(fc, auto_tensorize=True, n_tensorized_modes=2, rank=34, bias=True, factorization=factorization)
        np.testing.assert_array_almost_equal(fc(data).detach(), tfc(data).detach(), decimal=2)


def test_from_linear_list_multi_layer():
    """reference: test_factorized_linear.py:46-58 (two layers jointly factorized, n_layers = 2)."""
    torch.manual_seed(4)
    data = torch.randn(4, 9, dtype=torch.float64)
    fc1, fc2 = nn.Linear(9, 16).double(), nn.Linear(9, 16).double()
    with emulated_kernels():
        tfc = tlt.FactorizedLinear.from_linear_list([fc1, fc2], in_tensorized_features=(3, 3), out_tensorized_features=(4, 4), rank=38,
                                                    bias=True, decomposition_kwargs=dict(init="random", random_state=0, n_iter_max=400))
        np.testing.assert_array_almost_equal(fc1(data).detach(), tfc[0](data).detach(), decimal=2)
        np.testing.assert_array_almost_equal(fc2(data).detach(), tfc[1](data).detach(), decimal=2)


@pytest.mark.parametrize("factorization", ["CP", "Tucker", "TT"])
def test_from_conv_decomposes_the_kernel(factorization):
    """reference: factorized_layers/tests/test_factorized_convolution.py:66-74 (rank 30 on a 4 -> 5 channel 3x3 conv)."""
    torch.manual_seed(5)
    conv = nn.Conv2d(4, 5, 3, bias=True, padding=1).double()
    conv.weight.data.uniform_(-1, 1)
    data = torch.randn(2, 4, 8, 8, dtype=torch.float64)
    with emulated_kernels():
        fact = tlt.FactorizedConv.from_conv(conv, rank=30, factorization=factorization)
        np.testing.assert_array_almost_equal(conv(data).detach(), fact(data).detach(), decimal=2)
        rec = tensor_to_kernel(factorization, fact.weight.to_tensor())
        assert _rel(rec.detach(), conv.weight.data) < 1e-2


@pytest.mark.parametrize("unsqueezed_init", ["average", 1.2])
def test_tucker_init_unsqueezed_modes(unsqueezed_init):
    """reference: factorized_tensors/tests/test_factorizations.py:108-125."""
    torch.manual_seed(6)
    with emulated_kernels():
        tensor = tlt.FactorizedTensor.new((4, 4, 4), rank=(4, 1, 4), factorization="tucker")
        mat = torch.randn(4, 4)
        tensor.init_from_tensor(mat, unsqueezed_modes=[1], unsqueezed_init=unsqueezed_init)
        rec = tensor.to_tensor()
    coef = 0.25 if unsqueezed_init == "average" else unsqueezed_init
    for i in range(4):
        torch.testing.assert_close(rec[:, i], mat * coef, rtol=1e-4, atol=1e-5)
    with pytest.raises(ValueError):
        tensor.init_from_tensor(mat, unsqueezed_modes=[0])


def test_trl_init_from_linear():
    """reference: factorized_layers/tests/test_trl.py:150-190 (TRL initialised from a fully connected layer reproduces it)."""
    torch.manual_seed(7)
    fc = nn.Linear(3 * 4 * 5, 6).double()
    x = torch.randn(2, 3, 4, 5, dtype=torch.float64)
    with emulated_kernels():
        trl = tlt.TRL((3, 4, 5), (6,), rank="same", bias=True, factorization="tucker", dtype=torch.float64)
        trl.init_from_linear(fc)
        np.testing.assert_array_almost_equal(fc(x.reshape(2, -1)).detach(), trl(x).detach(), decimal=5)


# ------------------------------------------------------------------------------------------------------------------- lasso
@pytest.mark.parametrize("factorization", ["cp", "tucker", "tt"])
def test_tensor_lasso(factorization):
    """reference: tensor_hooks/tests/test_tensor_lasso.py:12-78, statement by statement."""
    with emulated_kernels():
        shape, rank = (5, 5, 6), 3
        tensor1 = tlt.FactorizedTensor.new(shape, rank, factorization=factorization).normal_()
        tensor2 = tlt.FactorizedTensor.new(shape, rank, factorization=factorization).normal_()
        lasso = tlt.tensor_lasso(factorization, penalty=1, clamp_weights=False, normalize_loss=False)
        lasso.apply(tensor1)
        lasso.apply(tensor2)
        value = 1.5
        lasso.set_weights(tensor1, value)
        lasso.set_weights(tensor2, value)
        torch.sum(tensor1())
        l1 = lasso.loss
        torch.sum(tensor2())
        l2 = lasso.loss
        count = {"tt": lambda t: sum(t.rank[1:-1]), "tucker": lambda t: sum(t.rank), "cp": lambda t: t.rank}[factorization]
        assert l1 == count(tensor1) * value
        assert l2 == (count(tensor1) + count(tensor2)) * value
        assert tensor1._forward_hooks and tensor2._forward_hooks
        lasso.reset()
        lasso.set_weights(tensor1, 0)
        lasso.set_weights(tensor2, 0)
        torch.sum(tensor1()) + torch.sum(tensor2())
        assert lasso.loss == 0
        tlt.remove_tensor_lasso(tensor1)
        assert not tensor1._forward_hooks and tensor2._forward_hooks
        tlt.remove_tensor_lasso(tensor2)
        assert not tensor2._forward_hooks
        lasso = tlt.tensor_lasso(factorization, penalty=1, normalize_loss=True, clamp_weights=False)
        lasso.apply(tensor1)
        tensor1()
        assert abs(lasso.loss - 1) < 1e-5
        tlt.remove_tensor_lasso(tensor1)


@pytest.mark.parametrize("factorization", ["tucker", "tt"])
def test_tensor_lasso_scales_the_reconstruction_and_has_gradients(factorization):
    """Gates of 0.5 on every rank scale the reconstruction by 0.5 ** n_gated_modes; the gates receive gradients through the
    contraction kernel (apply_lasso, _tensor_lasso.py:240-246 / :314-320)."""
    with emulated_kernels():
        t = tlt.FactorizedTensor.new((4, 5, 6), 3, factorization=factorization, dtype=torch.float64).normal_()
        ref = t.to_tensor().detach()
        lasso = tlt.tensor_lasso(factorization, penalty=0.1, clamp_weights=True, threshold=1e-6)
        lasso.apply(t)
        lasso.set_weights(t, 0.5)
        out = t().to_tensor()
        n_gates = len(list(t.lasso_weights))
        assert _rel(out.detach(), ref * 0.5 ** n_gates) < 1e-12
        (out.sum() + lasso.loss).backward()
        for g in t.lasso_weights:
            assert g.grad is not None and torch.isfinite(g.grad).all()
        # clamp + threshold: 3.0 -> 1.0, 1e-9 -> 0
        lasso.reset()
        with torch.no_grad():
            first = list(t.lasso_weights)[0]
            first.data.reshape(-1)[0] = 3.0
            first.data.reshape(-1)[1] = 1e-9
        t()
        vals = list(t.lasso_weights)[0].data.reshape(-1)
        assert vals[0] == 1.0 and vals[1] == 0.0


# ------------------------------------------------------------------------------------------------------------------- complex
@pytest.mark.parametrize("factorization", ["ComplexCP", "ComplexTucker", "ComplexTT", "ComplexDense"])
def test_complex_factorized_tensor(factorization):
    """reference: factorized_tensors/tests/test_factorizations.py:128-136, plus: parameters are stored real, the reconstruction
    equals the complex einsum of the factors."""
    shape = (4, 3, 2, 5)
    with emulated_kernels():
        t = tlt.FactorizedTensor.new(shape=shape, rank="same", factorization=factorization)
        t.normal_()
        full = t.to_tensor()
    assert tuple(full.shape) == shape and full.dtype == torch.cfloat
    assert all(p.dtype == torch.float32 and p.shape[-1] == 2 for p in t.parameters())
    if factorization == "ComplexCP":
        ref = torch.einsum("r,ar,br,cr,dr->abcd", t.weights, *list(t.factors))
    elif factorization == "ComplexTucker":
        ref = torch.einsum("wxyz,aw,bx,cy,dz->abcd", t.core, *list(t.factors))
    elif factorization == "ComplexTT":
        ref = torch.einsum("pax,xby,ycz,zdq->abcd", *list(t.factors))
    else:
        ref = t.tensor
    assert _rel(torch.view_as_real(full), torch.view_as_real(ref)) < 1e-5


@pytest.mark.parametrize("factorization", ["ComplexCP", "ComplexTucker", "ComplexTT", "ComplexDense"])
def test_complex_tensorized_matrix(factorization):
    """The tensorized counterparts (complex_tensorized_matrices.py:10-52): dtype cfloat out, float32 parameters, as
    test_factorizations.py:128-138 checks for the plain ones; the dense class stores (and returns) the matrix form."""
    row, col = (4, 3), (2, 5)
    with emulated_kernels():
        t = tlt.TensorizedTensor.new((row, col), rank=0.5, factorization=factorization)
        t.normal_()
        full = t.to_tensor()
        assert tuple(full.shape) == ((12, 10) if factorization == "ComplexDense" else row + col) and full.dtype == torch.cfloat
        m = t.to_matrix()
    assert tuple(m.shape) == (12, 10) and m.is_complex()
    assert all(p.dtype == torch.float32 for p in t.parameters())


def test_complex_contract_matches_einsum():
    from torch_b200 import ops
    torch.manual_seed(8)
    a = torch.randn(3, 4, 5, dtype=torch.complex64)
    b = torch.randn(5, 6, dtype=torch.complex64)
    d = torch.randn(6, dtype=torch.complex64)
    with emulated_kernels():
        out = ops.contract("abk,kn->abn", a, b, addend=d, addend_letters="n", alpha=0.5)
        out_r = ops.contract("abk,kn->abn", a, b.real.contiguous())
    ref = 0.5 * torch.einsum("abk,kn->abn", a, b) + d
    assert _rel(torch.view_as_real(out), torch.view_as_real(ref)) < 1e-5
    assert _rel(torch.view_as_real(out_r), torch.view_as_real(torch.einsum("abk,kn->abn", a, b.real.to(a.dtype)))) < 1e-5


# ------------------------------------------------------------------------------------------------------------------- same, real kernels
@pytest.mark.gpu
@pytest.mark.parametrize("factorization", ["cp", "tucker", "tt"])
def test_gpu_tensor_lasso(factorization):
    dev = torch.device("cuda:0")
    t = tlt.FactorizedTensor.new((6, 5, 7), 4, factorization=factorization, device=dev, dtype=torch.float32).normal_()
    ref = t.to_tensor().detach()
    lasso = tlt.tensor_lasso(factorization, penalty=0.1)
    lasso.apply(t)
    lasso.set_weights(t, 0.5)
    out = t().to_tensor()
    n_gates = 0 if factorization == "cp" else len(list(t.lasso_weights))
    assert _rel(out.detach(), ref * 0.5 ** n_gates) < 1e-5
    loss = out.sum() + lasso.loss
    loss.backward()
    gates = [t.lasso_weights] if factorization == "cp" else list(t.lasso_weights)
    for g in gates:
        assert g.grad is not None and torch.isfinite(g.grad).all()
    tlt.remove_tensor_lasso(t)
    assert not t._forward_hooks


@pytest.mark.gpu
@pytest.mark.parametrize("factorization", ["ComplexCP", "ComplexTucker", "ComplexTT"])
def test_gpu_complex_reconstruction(factorization):
    dev = torch.device("cuda:0")
    t = tlt.FactorizedTensor.new(shape=(4, 3, 2, 5), rank="same", factorization=factorization, device=dev)
    t.normal_()
    full = t.to_tensor()
    if factorization == "ComplexCP":
        ref = torch.einsum("r,ar,br,cr,dr->abcd", t.weights, *list(t.factors))
    elif factorization == "ComplexTucker":
        ref = torch.einsum("wxyz,aw,bx,cy,dz->abcd", t.core, *list(t.factors))
    else:
        ref = torch.einsum("pax,xby,ycz,zdq->abcd", *list(t.factors))
    assert full.dtype == torch.cfloat
    assert _rel(torch.view_as_real(full).detach(), torch.view_as_real(ref).detach()) < 1e-5
    full.real.sum().backward()
    assert all(p.grad is not None for p in t.parameters())


@pytest.mark.gpu
@pytest.mark.parametrize("factorization", ["CP", "Tucker", "TT"])
def test_gpu_from_conv(factorization):
    """from_conv on the device: decomposition through torch.linalg there, forward through the CUDA kernels."""
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    conv = nn.Conv2d(4, 5, 3, bias=True, padding=1).to(dev)
    conv.weight.data.uniform_(-1, 1)
    data = torch.randn(2, 4, 8, 8, device=dev)
    fact = tlt.FactorizedConv.from_conv(conv, rank=30, factorization=factorization)
    np.testing.assert_array_almost_equal(conv(data).detach().cpu(), fact(data).detach().cpu(), decimal=2)


@pytest.mark.gpu
@pytest.mark.parametrize("factorization", ["CP", "Tucker", "BlockTT"])
def test_gpu_from_linear(factorization):
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    data = torch.randn(4, 9, device=dev)
    fc = nn.Linear(9, 16, bias=True).to(dev)
    tfc = tlt.FactorizedLinear.from_linear(fc, auto_tensorize=False, in_tensorized_features=(3, 3), out_tensorized_features=(4, 4),
                                           rank=34, bias=True, factorization=factorization)
    np.testing.assert_array_almost_equal(fc(data).detach().cpu(), tfc(data).detach().cpu(), decimal=2)
