"""Maps every golden case (tests/golden/*.pt, produced by the real reference) onto the PRODUCT's public API
(torch_b200).  Used by the CPU host-logic tests (kernels emulated) and by the GPU parity tests (real kernels)."""
import torch

import torch_b200 as tlt
from torch_b200.functional.factorized_linear import linear_blocktt, linear_cp, linear_tucker
from torch_b200.functional.convolution import cp_conv, cp_conv_mobilenet, tucker_conv, tt_conv, convolve
from torch_b200.functional.tensor_regression import trl
from torch_b200.functional import factorized_linear

import _golden as G


def _leaf(t, device, dtype):
    return t.detach().to(device=device, dtype=dtype).clone().requires_grad_(True)


def build_weight(kind, params, device, dtype, tensorized_shape=None):
    """Golden param dict -> (product weight object, {reference parameter name: leaf tensor})."""
    P = torch.nn.Parameter
    leaves = {}
    fs = [P(_leaf(f, device, dtype)) for f in params["factors"]]
    for i, f in enumerate(fs):
        leaves[f"factors.factor_{i}"] = f
    if kind == "cp":
        w = P(_leaf(params["weights"], device, dtype))
        leaves["weights"] = w
        obj = tlt.CPTensorized(w, fs, tensorized_shape) if tensorized_shape else tlt.CPTensor(w, fs)
    elif kind == "tucker":
        c = P(_leaf(params["core"], device, dtype))
        leaves["core"] = c
        obj = tlt.TuckerTensorized(c, fs, tensorized_shape) if tensorized_shape else tlt.TuckerTensor(c, fs)
    elif kind == "tt":
        obj = tlt.TTTensor(fs)
    elif kind == "blocktt":
        rank = tuple(f.shape[0] for f in fs) + (1,)
        obj = tlt.BlockTT(fs, tensorized_shape, rank)
    else:
        raise ValueError(kind)
    return obj, leaves


def run_case(name, g, device, dtype, sub=None):
    """Run one golden case through the product; returns (y, grads{name}, expected_y, expected_grads)."""
    extra = {}
    x = _leaf(g["x"], device, dtype)
    if name.startswith("linear_"):
        ts = (tuple(g["out_shape"]), tuple(g["in_shape"]))
        w, leaves = build_weight(g["kind"], g["params"], device, dtype, ts)
        fn = {"cp": linear_cp, "tucker": linear_tucker, "blocktt": linear_blocktt}[g["kind"]]
        kw = {"plan": sub} if (g["kind"] == "blocktt" and sub) else {}
        y = fn(x, w, transpose=True, **kw).reshape(x.shape[0], -1)
    elif name == "factorized_linear_dispatch":
        w, leaves = build_weight(g["kind"], g["params"], device, dtype, g["tensorized_shape"])
        bias = _leaf(g["bias"], device, dtype)
        extra["extra0"] = bias
        y = factorized_linear(x, w, bias=bias, in_features=8, implementation=g["impl"])
    elif name.startswith("conv"):
        w, leaves = build_weight(g["kind"], g["params"], device, dtype)
        bias = None
        if g["bias"] is not None:
            bias = _leaf(g["bias"], device, dtype)
            extra["extra0"] = bias
        fn = {("cp", "factorized"): cp_conv, ("cp", "mobilenet"): cp_conv_mobilenet,
              ("tucker", "factorized"): tucker_conv, ("tt", "factorized"): tt_conv}.get((g["kind"], g["impl"]), convolve)
        y = fn(x, w, bias=bias, stride=g["stride"], padding=g["padding"], dilation=g["dilation"])
    elif name.startswith("trl_"):
        w, leaves = build_weight(g["kind"], g["params"], device, dtype)
        bias = None
        if g["bias"] is not None:
            bias = _leaf(g["bias"], device, dtype)
            extra["extra0"] = bias
        kw = {"project_input": True} if g["project_input"] else {}
        y = trl(x, w, bias=bias, **kw)
    elif name == "tcl":
        layer = tlt.TCL(tuple(g["x"].shape[1:]), tuple(f.shape[0] for f in g["factors"]))
        leaves = {}
        for i, f in enumerate(g["factors"]):
            p = torch.nn.Parameter(_leaf(f, device, dtype))
            setattr(layer, f"factor_{i}", p)
            leaves[f"factor_{i}"] = p
        y = layer(x)
    else:
        raise ValueError(name)
    leaves = {"x": x, **leaves, **extra}
    names = [n for n in leaves if n in g["grads"]]
    gy = g["gy"].to(device=device, dtype=dtype)
    grads = torch.autograd.grad(y, [leaves[n] for n in names], gy, allow_unused=True)
    return y, dict(zip(names, grads)), g["y"], g["grads"]


def all_cases():
    out = []
    for name in G.names():
        if name in ("reconstruction", "dropout") or name.startswith(("nlayers_", "embedding_")):
            continue
        if name == "factorized_linear_dispatch":
            out += [(name, k) for k in G.load(name)]
        elif name.startswith("linear_blocktt"):
            out += [(name, "dense"), (name, "sweep")]
        else:
            out.append((name, None))
    return out


def load_case(name, sub):
    g = G.load(name)
    if name == "factorized_linear_dispatch":
        return g[sub], None
    return g, sub


def module_cases():
    return [n for n in G.names() if n.startswith(("nlayers_", "embedding_"))]


def run_module_case(name, g, device, dtype):
    """Golden cases recorded from the reference's nn.Modules (n_layers > 1 linears / convs, embeddings): build the product's
    module with the same constructor arguments, load the reference's state_dict (same key names) and run fwd + bwd.
    Returns (y, grads{state_dict name}, expected_y, expected_grads)."""
    ctor = dict(g["ctor"])
    if name.startswith("nlayers_linear"):
        layer = tlt.FactorizedLinear(ctor.pop("in_tensorized_features"), ctor.pop("out_tensorized_features"), dtype=torch.float64, **ctor)
    elif name.startswith("nlayers_conv"):
        layer = tlt.FactorizedConv(ctor.pop("in_channels"), ctor.pop("out_channels"), ctor.pop("kernel_size"), dtype=torch.float64, **ctor)
    else:
        layer = tlt.FactorizedEmbedding(ctor.pop("num_embeddings"), ctor.pop("embedding_dim"), dtype=torch.float64, **ctor)
    missing = layer.load_state_dict(g["state_dict"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    layer = layer.to(device=device, dtype=dtype)
    x = g["x"].to(device)
    if x.is_floating_point():
        x = x.to(dtype).requires_grad_(True)
    y = layer(x, g["index"]) if g["index"] else layer(x)
    params = dict(layer.named_parameters())
    leaves = {"x": x, **params} if x.is_floating_point() else dict(params)
    names = [n for n in leaves if n in g["grads"]]
    gy = g["gy"].to(device=device, dtype=dtype)
    grads = torch.autograd.grad(y, [leaves[n] for n in names], gy, allow_unused=True)
    return y, dict(zip(names, grads)), g["y"], g["grads"]
