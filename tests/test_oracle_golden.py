"""Pins the CPU oracle (oracle/reference_path.py) against outputs of the REAL reference (tests/golden/*.pt)."""
import pytest
import torch

from oracle import reference_path as rp
from oracle import tensorly_semantics as tls
import _golden as G

TOL = 1e-12


def _check(fn, leaves, names, g, extra_names=()):
    y, grads = rp.fwd_bwd(fn, leaves, g["gy"])
    assert G.rel_err(y, g["y"]) < TOL
    for n, gr in zip(names, grads):
        if n in g["grads"]:
            assert G.rel_err(gr, g["grads"][n]) < TOL, n


@pytest.mark.parametrize("name", G.names("linear_"))
def test_linear(name):
    g = G.load(name)
    kind = g["kind"]
    leaves, rebuild, pnames = G.params_list(kind, g["params"])
    in_shape = tuple(g["in_shape"])
    B = g["x"].shape[0]
    for order in (("reference", "sane") if kind != "blocktt" else ("reference",)):
        def fn(x, *ps):
            p = rebuild(ps)
            if kind == "cp":
                return rp.cp_linear(x, p[0], p[1], in_shape, order=order).reshape(B, -1)
            if kind == "tucker":
                return rp.tucker_linear(x, p[0], p[1], in_shape, order=order).reshape(B, -1)
            return rp.blocktt_linear(x, p, in_shape)
        _check(fn, [g["x"]] + leaves, ["x"] + pnames, g)
    # the reference's own identity: factorized == x @ to_matrix().T  (functional/tests/test_factorized_linear.py:15-32)
    W = rp.to_matrix(rp.reconstruct(kind, rebuild(leaves), (tuple(g["out_shape"]), in_shape)), g["W"].shape)
    assert G.rel_err(W, g["W"]) < TOL
    assert G.rel_err(g["x"] @ W.T, g["y"]) < 1e-10


def test_factorized_linear_dispatch():
    allg = G.load("factorized_linear_dispatch")
    for key, g in allg.items():
        kind, impl = g["kind"], g["impl"]
        leaves, rebuild, pnames = G.params_list(kind, g["params"])

        def fn(x, bias, *ps):
            return rp.factorized_linear(x, kind, rebuild(ps), g["tensorized_shape"], bias=bias, implementation=impl)
        _check(fn, [g["x"], g["bias"]] + leaves, ["x", "extra0"] + pnames, g)


@pytest.mark.parametrize("name", G.names("conv"))
def test_conv(name):
    g = G.load(name)
    kind, impl = g["kind"], g["impl"]
    leaves, rebuild, pnames = G.params_list(kind, g["params"])
    has_bias = g["bias"] is not None
    kw = dict(stride=g["stride"], padding=g["padding"], dilation=g["dilation"])

    def fn(x, *rest):
        bias = rest[0] if has_bias else None
        p = rebuild(rest[1:] if has_bias else rest)
        if impl == "reconstructed":
            return rp.reconstructed_conv(x, kind, p, bias=bias, **kw)
        if kind == "cp":
            f = rp.cp_conv if impl == "factorized" else rp.cp_conv_mobilenet
            return f(x, p[0], p[1], bias=bias, **kw)
        if kind == "tucker":
            return rp.tucker_conv(x, p[0], p[1], bias=bias, **kw)
        return rp.tt_conv(x, p, bias=bias, **kw)
    lv = [g["x"]] + ([g["bias"]] if has_bias else []) + leaves
    nm = ["x"] + (["extra0"] if has_bias else []) + pnames
    _check(fn, lv, nm, g)
    # reference identity: factorized conv == dense conv with the reconstructed kernel (test_factorized_convolution.py:13-57)
    dense = rp.convolve(g["x"], g["kernel"], bias=g["bias"], **kw)
    assert G.rel_err(dense, g["y"]) < 1e-10


@pytest.mark.parametrize("name", G.names("trl_"))
def test_trl(name):
    g = G.load(name)
    kind = g["kind"]
    leaves, rebuild, pnames = G.params_list(kind, g["params"])
    has_bias = g["bias"] is not None

    def fn(x, *rest):
        bias = rest[0] if has_bias else None
        p = rebuild(rest[1:] if has_bias else rest)
        return rp.trl(x, kind, p, bias=bias, project_input=g["project_input"])
    lv = [g["x"]] + ([g["bias"]] if has_bias else []) + leaves
    nm = ["x"] + (["extra0"] if has_bias else []) + pnames
    _check(fn, lv, nm, g)


def test_tcl():
    g = G.load("tcl")
    n = len(g["factors"])
    _check(lambda x, *f: rp.tcl(x, list(f)), [g["x"]] + list(g["factors"]), ["x"] + [f"factor_{i}" for i in range(n)], g)


def test_reconstruction():
    allg = G.load("reconstruction")
    for key, g in allg.items():
        kind = g["kind"]
        leaves, rebuild, _ = G.params_list(kind if not (key.startswith("tensor_tt")) else "tt", g["params"])
        full = rp.reconstruct(kind, rebuild(leaves), g.get("tensorized_shape"))
        assert G.rel_err(full.reshape(g["full"].shape), g["full"]) < TOL, key


def test_dropout_matches_reference_rng_and_scaling():
    allg = G.load("dropout")
    for key, g in allg.items():
        kind = key.split("_")[0]
        p = g["params"]
        torch.manual_seed(g["seed"])
        if kind == "tucker":
            idx = [rp.sample_keep_indices(r, g["proba"], g["min_dim"]) for r in g["rank"]]
            core, factors = rp.tucker_dropout_apply(p["core"], p["factors"], idx, g["proba"], g["training"])
            assert torch.equal(core, g["dropped"]["core"])
            for a, b in zip(factors, g["dropped"]["factors"]):
                assert torch.equal(a, b)
        elif kind == "cp":
            idx = rp.sample_keep_indices(g["rank"], g["proba"], g["min_dim"])
            w, factors = rp.cp_dropout_apply(p["weights"], p["factors"], idx, g["proba"], g["training"])
            assert torch.equal(w, g["dropped"]["weights"])
            for a, b in zip(factors, g["dropped"]["factors"]):
                assert torch.equal(a, b)
        else:
            idx = [rp.sample_keep_indices(r, g["proba"], g["min_dim"]) for r in g["rank"][1:]]
            cores = rp.tt_dropout_apply(p["factors"], idx, g["proba"], g["training"])
            for a, b in zip(cores, g["dropped"]["factors"]):
                assert torch.equal(a, b)


def test_rank_validators_match_survey_configs():
    # config-1 CP rank, config-2 BlockTT rank, config-4 Tucker conv rank, config-5 ranks
    assert tls.validate_cp_rank((4, 8, 16, 4, 8, 16), 0.5) == 2341
    assert tls.validate_tt_rank((64, 96, 384), 16) == (1, 16, 16, 1)
    assert tls.validate_tucker_rank((512, 512, 3, 3), 0.5) == (388, 388, 2, 2)
    assert tls.validate_tucker_rank((512, 512, 3, 3), 0.5, fixed_modes=[2, 3]) == (310, 310, 3, 3)
    assert tls.validate_tucker_rank((16,) * 6, 0.75) == (15,) * 6
    assert tls.validate_tt_rank((256, 256, 256), 0.75) == (1, 221, 221, 1)
    for cout, cin, r in ((64, 64, 69), (128, 64, 93), (128, 128, 141), (256, 128, 189), (256, 256, 285),
                         (512, 256, 381), (512, 512, 573)):
        assert tls.validate_cp_rank((cout, cin, 3, 3), 0.25) == r
