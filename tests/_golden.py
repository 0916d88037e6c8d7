"""Helpers shared by the oracle (CPU) and CUDA (GPU) parity tests: load golden fixtures produced by the real
reference (tests/golden/make_golden.py) and evaluate the oracle on them."""
import glob
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(prefix=""):
    return sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.pt")))


def load(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def rel_err(a, b):
    """relative Frobenius error ||a-b|| / ||b|| (the metric north_star states tolerances in)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.norm()
    if denom == 0:
        return (a - b).norm().item()
    return ((a - b).norm() / denom).item()


def params_list(kind, p):
    """golden params dict -> (leaf list, rebuild(list) -> oracle params)"""
    if kind == "cp":
        n = len(p["factors"])
        return [p["weights"], *p["factors"]], (lambda l: (l[0], list(l[1:1 + n]))), ["weights"] + [f"factors.factor_{i}" for i in range(n)]
    if kind == "tucker":
        n = len(p["factors"])
        return [p["core"], *p["factors"]], (lambda l: (l[0], list(l[1:1 + n]))), ["core"] + [f"factors.factor_{i}" for i in range(n)]
    if kind in ("tt", "blocktt"):
        n = len(p["factors"])
        return list(p["factors"]), (lambda l: list(l[:n])), [f"factors.factor_{i}" for i in range(n)]
    raise ValueError(kind)
