"""Host-buffer pipeline (torch_b200.pipeline.HostPipelinedStep) on the GPU: results of every submitted step, in order, equal the
same layer run directly -- with device staging buffers and with two captured replicas of the step (no staging copies), the form
bench.py's e2e leg times."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("replicated", [False, True])
def test_host_pipelined_step_matches_direct_steps(replicated):
    import torch_b200 as tlt
    from torch_b200.graphs import GraphedStep
    from torch_b200.pipeline import HostPipelinedStep
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    layer = tlt.FactorizedLinear((4, 8, 8), (8, 8, 4), bias=True, factorization="blocktt", rank=8, device=dev).to(torch.bfloat16)
    params = list(layer.parameters())

    def make(x, gy):
        def step():
            x.grad = None
            for p in params:
                p.grad = None
            layer(x).backward(gy)
        return step

    def capture(x, gy):
        g = GraphedStep(make(x, gy), results=lambda: [x.grad] + [p.grad for p in params])
        res = g.results
        flat = torch.cat([t.reshape(-1) for t in res[1:]])

        def outs():
            torch.cat([t.reshape(-1) for t in res[1:]], out=flat)
            return [res[0], flat]
        return g, outs

    B = 96
    ins = [(torch.randn(B, 256, device=dev, dtype=torch.bfloat16).requires_grad_(True), torch.randn(B, 256, device=dev, dtype=torch.bfloat16))
           for _ in range(2 if replicated else 1)]
    caps = [capture(x, gy) for x, gy in ins]
    if replicated:
        pipe = HostPipelinedStep(list(ins[0]), caps[0][0].replay, caps[0][1],
                                 replicas=[(list(i), c[0].replay, c[1]) for i, c in zip(ins, caps)])
    else:
        pipe = HostPipelinedStep(list(ins[0]), caps[0][0].replay, caps[0][1])
    steps = 7
    host_in = [[torch.randn(B, 256, dtype=torch.bfloat16).pin_memory(), torch.randn(B, 256, dtype=torch.bfloat16).pin_memory()] for _ in range(steps)]
    n_par = sum(p.numel() for p in params)
    host_out = [[torch.empty(B, 256, dtype=torch.bfloat16).pin_memory(), torch.empty(n_par, dtype=torch.bfloat16).pin_memory()] for _ in range(steps)]
    for i in range(steps):
        pipe.submit(host_in[i], host_out[i])
    pipe.drain()
    for i in range(steps):
        x = host_in[i][0].to(dev).requires_grad_(True)
        grads = torch.autograd.grad(layer(x), [x] + params, host_in[i][1].to(dev))
        want = [grads[0], torch.cat([g.reshape(-1) for g in grads[1:]])]
        for got, w in zip(host_out[i], want):
            err = float((got.double() - w.cpu().double()).norm() / w.cpu().double().norm().clamp_min(1e-30))
            assert err < 5e-3, f"step {i}: the pipeline's results differ from the direct step (rel err {err:.2e})"
