"""Side streams for ops that issue two independent branches of launches (a data-gradient chain beside a weight-gradient chain).

``fork(device)`` returns ``(side, main)``: ``side`` already waits for everything queued on the current stream ``main``; the caller
issues one branch under ``torch.cuda.stream(side)``, the other on ``main``, and ends with ``join(side, main)``.  Every buffer
both branches touch must be allocated before the fork (on ``main``) and outlive the join.  Captured into a CUDA graph the fork / join
are plain graph edges.  One side stream per device; ``None`` off the GPU or when the switch named by ``env`` is "0"."""
import os

import torch

_SIDE = {}


def side_stream(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


def fork(device, env=None):
    if device.type != "cuda" or (env is not None and os.environ.get(env, "1") == "0"):
        return None, None
    side, main = side_stream(device), torch.cuda.current_stream(device)
    side.wait_stream(main)
    return side, main


def join(side, main):
    if side is not None:
        main.wait_stream(side)
