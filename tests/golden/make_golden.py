"""Generate tests/golden/*.pt by executing the REAL reference package (/root/reference/tltorch).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

tensorly, the reference's un-pinned third-party dependency, is not installable offline, so the reference is
imported on top of ``oracle/_shim/tensorly`` (an adapter over oracle/tensorly_semantics.py).  Everything that
lives in the reference itself -- einsum-string construction, conv chains, BlockTT reconstruction, layer
dispatch, dropout hooks -- is executed as shipped.  Inputs (activations, every parameter, upstream gradients) are
bf16-representable values evaluated in fp64, so the stored outputs are exact to ~1e-15 and the fp64 oracle, the fp32 and
the bf16 CUDA runs all consume EXACTLY the same numbers: the bf16 parity gate (2e-2) measures the kernels' arithmetic only.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(REPO, "oracle", "_shim"))
sys.path.insert(0, "/root/reference")          # the REAL tltorch must win over the repo's drop-in alias package of the same name
sys.path.append(REPO)

import warnings  # noqa: E402
warnings.simplefilter("ignore")

import tltorch  # noqa: E402  (the real reference)
from tltorch.functional.factorized_linear import linear_blocktt, linear_cp, linear_tucker  # noqa: E402
from tltorch.functional import factorized_linear  # noqa: E402
from tltorch.functional.convolution import (cp_conv, cp_conv_mobilenet, tucker_conv, tt_conv, convolve)  # noqa: E402
from tltorch.functional.tensor_regression import trl  # noqa: E402
from tltorch.tensor_hooks._tensor_dropout import TuckerDropout, CPDropout, TTDropout  # noqa: E402

assert tltorch.__file__.startswith("/root/reference"), tltorch.__file__
D = torch.float64


def rnd(*shape):
    """bf16-representable values held in fp64."""
    return torch.randn(*shape, dtype=torch.float32).to(torch.bfloat16).to(D)


def init(w, std=1.0):
    w.normal_(0, std)
    with torch.no_grad():
        for p in w.parameters():
            p.copy_(p.to(torch.bfloat16).to(D))
        if hasattr(w, "weights") and w.weights is not None:   # make CP lambda non-trivial
            w.weights.copy_((1 + 0.25 * torch.randn_like(w.weights)).to(torch.bfloat16).to(D))
    return w


def params_of(w):
    out = {}
    if hasattr(w, "weights"):
        out["weights"] = w.weights.detach().clone()
    if hasattr(w, "core"):
        out["core"] = w.core.detach().clone()
    if hasattr(w, "factors"):
        out["factors"] = [f.detach().clone() for f in w.factors]
    if hasattr(w, "tensor") and torch.is_tensor(getattr(w, "tensor")):
        out["tensor"] = w.tensor.detach().clone()
    return out


def run(fn, x, w, extra=()):
    """fwd + bwd through the reference; returns y and grads wrt x, every parameter of w and extras (bias)."""
    x = x.detach().clone().requires_grad_(True)
    leaves = [x] + [p for p in w.parameters()] + [e for e in extra if e is not None]
    for e in extra:
        if e is not None:
            e.requires_grad_(True)
    y = fn(x)
    gy = rnd(*y.shape)
    grads = torch.autograd.grad(y, leaves, gy, allow_unused=True)
    names = ["x"] + [n for n, _ in w.named_parameters()] + [f"extra{i}" for i, e in enumerate(extra) if e is not None]
    return {"y": y.detach(), "gy": gy, "grads": {n: g.detach() for n, g in zip(names, grads) if g is not None}}


CASES = {}


def case(name):
    def deco(f):
        CASES[name] = f
        return f
    return deco


# ------------------------------------------------------------------ linear
def _linear_case(fact, fn, out_shape, in_shape, rank, batch, seed):
    torch.manual_seed(seed)
    w = init(tltorch.TensorizedTensor.new((out_shape, in_shape), rank=rank, factorization=fact, dtype=D))
    x = rnd(batch, int(torch.tensor(in_shape).prod()))
    res = run(lambda t: fn(t, w, transpose=True).reshape(batch, -1), x, w)
    res.update(x=x, params=params_of(w), out_shape=out_shape, in_shape=in_shape, rank=w.rank,
               W=w.to_matrix().detach(), kind=fact)
    return res


for _f, _fn in (("blocktt", linear_blocktt), ("cp", linear_cp), ("tucker", linear_tucker)):
    CASES[f"linear_{_f}_ref_test"] = (lambda f=_f, fn=_fn: _linear_case(f, fn, (6, 2), (4, 5), 3, 2, 11))
    CASES[f"linear_{_f}_order3"] = (lambda f=_f, fn=_fn: _linear_case(f, fn, (3, 2, 4), (2, 3, 4), 0.6, 7, 12))
CASES["linear_blocktt_order4"] = lambda: _linear_case("blocktt", linear_blocktt, (2, 3, 2, 2), (3, 2, 2, 3), 4, 5, 13)


@case("factorized_linear_dispatch")
def _():
    out = {}
    for fact in ("cp", "tucker", "blocktt"):
        for impl in ("factorized", "reconstructed"):
            torch.manual_seed(21)
            w = init(tltorch.TensorizedTensor.new(((3, 4), (4, 2)), rank=0.7, factorization=fact, dtype=D))
            bias = rnd(12)
            x = rnd(3, 5, 8)        # extra leading dims
            r = run(lambda t: factorized_linear(t, w, bias=bias, in_features=8, implementation=impl), x, w, extra=(bias,))
            r.update(x=x, params=params_of(w), bias=bias.detach().clone(), kind=fact, impl=impl,
                     tensorized_shape=((3, 4), (4, 2)))
            out[f"{fact}_{impl}"] = r
    return out


# ------------------------------------------------------------------ conv
def _conv_case(fact, impl, order, cin, cout, k, spatial, rank, stride, padding, dilation, batch, seed, bias=True):
    torch.manual_seed(seed)
    kshape = (cout, cin) + (k,) * order
    fshape = kshape[1:] + (cout,) if fact == "tt" else kshape
    w = init(tltorch.FactorizedTensor.new(fshape, rank=rank, factorization=fact, dtype=D))
    b = rnd(cout) if bias else None
    x = rnd(batch, cin, *spatial)
    fn = {("cp", "factorized"): cp_conv, ("cp", "mobilenet"): cp_conv_mobilenet, ("tucker", "factorized"): tucker_conv,
          ("tt", "factorized"): tt_conv}.get((fact, impl), convolve)
    r = run(lambda t: fn(t, w, bias=b, stride=stride, padding=padding, dilation=dilation), x, w, extra=(b,))
    full = w.to_tensor().detach()
    if fact == "tt":
        full = torch.movedim(full, -1, 0)
    r.update(x=x, params=params_of(w), bias=b, kind=fact, impl=impl, stride=stride, padding=padding,
             dilation=dilation, kernel=full, rank=w.rank)
    return r


for _fact, _impl in (("cp", "factorized"), ("cp", "mobilenet"), ("cp", "reconstructed"), ("tucker", "factorized"),
                     ("tucker", "reconstructed"), ("tt", "factorized"), ("tt", "reconstructed")):
    CASES[f"conv2d_{_fact}_{_impl}_ref_test"] = (
        lambda f=_fact, i=_impl: _conv_case(f, i, 2, 4, 5, 3, (8, 7), 0.5, 1, 1, 1, 1, 31))
for _fact, _impl in (("cp", "factorized"), ("cp", "mobilenet"), ("tucker", "factorized"), ("tt", "factorized")):
    CASES[f"conv2d_{_fact}_{_impl}_strided"] = (
        lambda f=_fact, i=_impl: _conv_case(f, i, 2, 6, 8, 3, (9, 10), 0.5, [2, 1] if i != "mobilenet" else (2, 1),
                                             [1, 2] if i != "mobilenet" else (1, 2), [1, 2] if i != "mobilenet" else (1, 2), 3, 32))
    CASES[f"conv1d_{_fact}_{_impl}"] = (
        lambda f=_fact, i=_impl: _conv_case(f, i, 1, 5, 4, 3, (11,), 0.8, 2, 1, 1, 2, 33))
    CASES[f"conv3d_{_fact}_{_impl}"] = (
        lambda f=_fact, i=_impl: _conv_case(f, i, 3, 3, 4, 3, (5, 6, 4), 0.8, 1, 1, 1, 2, 34, bias=False))


# ------------------------------------------------------------------ TRL / TCL
def _trl_case(fact, in_shape, out_shape, rank, batch, seed, project_input=None, bias=True):
    torch.manual_seed(seed)
    w = init(tltorch.FactorizedTensor.new(in_shape + out_shape, rank=rank, factorization=fact, dtype=D))
    b = rnd(*out_shape) if bias else None
    x = rnd(batch, *in_shape)
    kw = {} if project_input is None else {"project_input": project_input}
    r = run(lambda t: trl(t, w, bias=b, **kw), x, w, extra=(b,))
    r.update(x=x, params=params_of(w), bias=b, kind=fact, project_input=bool(project_input), rank=w.rank,
             in_shape=in_shape, out_shape=out_shape)
    return r


CASES["trl_tucker"] = lambda: _trl_case("tucker", (3, 4), (2,), (2, 3, 2), 5, 41)
CASES["trl_tucker_project"] = lambda: _trl_case("tucker", (3, 4, 2), (2, 3), (2, 3, 2, 2, 2), 5, 42, project_input=True)
CASES["trl_cp"] = lambda: _trl_case("cp", (3, 4), (2,), 4, 5, 43)
CASES["trl_tt"] = lambda: _trl_case("tt", (3, 4), (2, 2), (1, 2, 3, 2, 1), 5, 44, bias=False)


@case("tcl")
def _():
    torch.manual_seed(51)
    layer = tltorch.TCL((4, 5, 6), (2, 3, 5), dtype=D)
    with torch.no_grad():
        for p in layer.parameters():
            p.copy_(p.to(torch.bfloat16).to(D))
    x = rnd(2, 4, 5, 6)
    r = run(lambda t: layer(t), x, layer)
    r.update(x=x, factors=[f.detach().clone() for f in layer.factors])
    return r


# ------------------------------------------------------------------ reconstruction
@case("reconstruction")
def _():
    out = {}
    torch.manual_seed(61)
    for fact, rank in (("cp", 5), ("tucker", (3, 2, 2, 4)), ("tt", (1, 3, 4, 2, 1))):
        w = init(tltorch.FactorizedTensor.new((4, 3, 2, 5), rank=rank, factorization=fact, dtype=D))
        out[f"tensor_{fact}"] = dict(params=params_of(w), full=w.to_tensor().detach(), kind=fact)
    for fact in ("cp", "tucker", "blocktt"):
        ts = ((4, 3, 2), (5, 3, 2))
        w = init(tltorch.TensorizedTensor.new(ts, rank=0.5, factorization=fact, dtype=D))
        out[f"matrix_{fact}"] = dict(params=params_of(w), full=w.to_matrix().detach(), kind=fact, tensorized_shape=ts,
                                     rank=w.rank)
    ts = (3, (4, 2), (2, 5))        # batched BlockTT (n_layers-style leading int mode)
    w = init(tltorch.TensorizedTensor.new(ts, rank=3, factorization="blocktt", dtype=D))
    out["matrix_blocktt_batched"] = dict(params=params_of(w), full=w.to_matrix().detach(), kind="blocktt",
                                         tensorized_shape=ts, rank=w.rank)
    return out


# ------------------------------------------------------------------ dropout hooks
@case("dropout")
def _():
    out = {}
    for name, fact, shape, rank, hook_cls in (("tucker", "tucker", (10, 11, 12), (7, 6, 3), TuckerDropout),
                                              ("cp", "cp", (10, 11, 12), 9, CPDropout),
                                              ("tt", "tt", (10, 11, 12), (1, 6, 3, 1), TTDropout)):
        for training in (True, False):
            torch.manual_seed(71)
            w = init(tltorch.FactorizedTensor.new(shape, rank=rank, factorization=fact, dtype=D))
            hook = hook_cls(0.4, min_dim=3, min_values=1, drop_test=True)
            torch.manual_seed(1234)
            dropped = hook._apply_tensor_dropout(w, training=training)
            out[f"{name}_train{int(training)}"] = dict(params=params_of(w), dropped=params_of(dropped), proba=0.4,
                                                       min_dim=3, seed=1234, rank=w.rank, training=training,
                                                       full=dropped.to_tensor().detach())
    return out


# ------------------------------------------------------------------ n_layers > 1 (jointly factorized layers) and embeddings
def _round_module(m):
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(p.to(torch.bfloat16).to(D))
    return m


def _module_case(layer, x, call):
    """fwd + bwd of a reference nn.Module; gradients keyed by state_dict name."""
    x = x.detach().clone()
    if x.is_floating_point():
        x.requires_grad_(True)
    names = [n for n, _ in layer.named_parameters()]
    leaves = [p for _, p in layer.named_parameters()]
    y = call(layer, x)
    gy = rnd(*y.shape)
    grads = torch.autograd.grad(y, ([x] if x.requires_grad else []) + leaves, gy, allow_unused=True)
    gnames = (["x"] if x.requires_grad else []) + names
    return {"x": x.detach(), "y": y.detach(), "gy": gy, "state_dict": {k: v.detach().clone() for k, v in layer.state_dict().items()},
            "grads": {n: g.detach() for n, g in zip(gnames, grads) if g is not None}}


def _nlayers_linear(fact, seed):
    torch.manual_seed(seed)
    layer = _round_module(tltorch.FactorizedLinear((4, 3), (3, 2), bias=True, factorization=fact, rank=0.6, n_layers=2, dtype=D))
    with torch.no_grad():
        layer.bias.copy_(rnd(*layer.bias.shape))
    r = _module_case(layer, rnd(5, 12), lambda l, t: l(t, 1))
    r.update(kind=fact, index=1, ctor=dict(in_tensorized_features=(4, 3), out_tensorized_features=(3, 2), bias=True,
                                          factorization=fact, rank=0.6, n_layers=2))
    return r


def _nlayers_conv(fact, seed):
    torch.manual_seed(seed)
    layer = tltorch.FactorizedConv(4, 5, 3, order=2, padding=1, n_layers=2, factorization=fact, rank=0.5, bias=True, dtype=D)
    layer.reset_parameters(std=0.5)
    _round_module(layer)
    with torch.no_grad():
        layer.bias.copy_(rnd(*layer.bias.shape))
    r = _module_case(layer, rnd(2, 4, 6, 5), lambda l, t: l(t, 1))
    r.update(kind=fact, index=1, ctor=dict(in_channels=4, out_channels=5, kernel_size=3, order=2, padding=1, n_layers=2,
                                          factorization=fact, rank=0.5, bias=True))
    return r


def _embedding(fact, seed, n_layers=1):
    torch.manual_seed(seed)
    ctor = dict(num_embeddings=24, embedding_dim=8, auto_tensorize=False, tensorized_num_embeddings=(2, 3, 4),
                tensorized_embedding_dim=(2, 2, 2), factorization=fact, rank=3 if fact != "tucker" else 0.8, n_layers=n_layers)
    layer = tltorch.FactorizedEmbedding(dtype=D, **ctor)
    with torch.no_grad():
        for p in layer.parameters():
            p.copy_((3 * p).to(torch.bfloat16).to(D))
    idx = torch.tensor([[1, 5, 23], [0, 7, 7]])
    r = _module_case(layer, idx, (lambda l, t: l(t)) if n_layers == 1 else (lambda l, t: l(t, 1)))
    r.update(kind=fact, ctor=ctor, index=0 if n_layers == 1 else 1)
    return r


for _f in ("cp", "blocktt"):
    CASES[f"nlayers_linear_{_f}"] = (lambda f=_f: _nlayers_linear(f, 81))
for _f in ("cp", "tucker", "tt"):
    CASES[f"nlayers_conv_{_f}"] = (lambda f=_f: _nlayers_conv(f, 82))
for _f in ("blocktt", "cp", "tucker"):
    CASES[f"embedding_{_f}"] = (lambda f=_f: _embedding(f, 83))
CASES["embedding_blocktt_nlayers"] = lambda: _embedding("blocktt", 84, n_layers=2)


def main():
    total = 0
    for name, fn in CASES.items():
        data = fn()
        path = os.path.join(HERE, f"{name}.pt")
        torch.save(data, path)
        total += os.path.getsize(path)
        print(f"{name:45s} {os.path.getsize(path):8d} B")
    print("total bytes", total)


if __name__ == "__main__":
    main()
