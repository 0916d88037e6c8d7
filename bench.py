#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native TensorLy-Torch factorized-layer hot path.

Workload (BASELINE.json configs[1]): FactorizedLinear BlockTT (TT-matrix) layers replacing the BERT-base FFN
768 -> 3072 -> 768 (tensorized (8,8,12) / (8,12,32), TT rank 16), 64 x 512 = 32768 tokens PER GPU, bf16,
forward + backward (gradients w.r.t. the input, every TT core and both biases).  A "step" is one fwd+bwd pass of the
two layers over that batch; with N > 1 GPUs each rank owns its own 32768 tokens (weak scaling) and the core / bias
gradients are all-reduced (one flat NCCL bucket) inside the step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (contract in the task statement): value = whole-job tokens/s with inputs resident in HBM;
e2e = same metric with HOST (pinned) inputs, H2D/D2H copies inside the timed region; roofline = tcgen05 GEMM kernel
against the measured bf16 peak; cpu_baseline = the oracle port of the reference's op sequence on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

IN_TS, HID_TS = (8, 8, 12), (8, 12, 32)
IN_F, HID_F = 768, 3072
TT_RANK = 16
TOKENS_PER_GPU = 64 * 512
METRIC = "factorized_layer_fwd_bwd_samples_per_s"
UNIT = "samples/s"
WORKLOAD = ("BERT-base FFN 768->3072->768 as two FactorizedLinear(blocktt, tensorized (8,8,12)/(8,12,32), rank 16, "
            "bias) layers, 64x512 tokens per GPU, bf16, fwd+bwd")


def _config(n_gpus, extra=None):
    cfg = {"workload": WORKLOAD, "tokens_per_gpu": TOKENS_PER_GPU, "global_tokens": TOKENS_PER_GPU * n_gpus,
           "parallelism": f"dp{n_gpus}", "l2": "working set per step (~0.9 GB) >> 126 MB L2, no explicit flush",
           "nonlinearity": "none between the two layers (not part of the factorized-layer path)"}
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle restatement of the reference's op sequence (linear_blocktt einsum + bias, autograd backward)
# ---------------------------------------------------------------------------------------------------------------------
def _cpu_step_fn(tokens, dtype):
    import torch
    from oracle import reference_path as rp
    torch.manual_seed(1234)
    import math

    def cores(in_ts, out_ts):
        rank = (1, TT_RANK, TT_RANK, 1)
        std = (math.sqrt(5) / math.sqrt(math.prod(in_ts)) / math.prod(rank)) ** 0.5
        return [(torch.randn(rank[k], out_ts[k], in_ts[k], rank[k + 1]) * std).to(dtype).requires_grad_(True) for k in range(3)]
    up, down = cores(IN_TS, HID_TS), cores(HID_TS, IN_TS)
    b_up = torch.zeros(HID_F, dtype=dtype, requires_grad=True)
    b_down = torch.zeros(IN_F, dtype=dtype, requires_grad=True)
    x = torch.randn(tokens, IN_F).to(dtype).requires_grad_(True)
    gy = torch.randn(tokens, IN_F).to(dtype)
    leaves = [x, b_up, b_down] + up + down

    def step():
        h = rp.blocktt_linear(x, up, IN_TS) + b_up              # factorized_linear.py:39-60 + linear.py:42-43
        y = rp.blocktt_linear(h, down, HID_TS) + b_down
        return torch.autograd.grad(y, leaves, gy)
    return step


def cpu_baseline(sample_tokens=2048, reps=3):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    best, best_dt = 0.0, None
    for dtype in (torch.float32, torch.bfloat16):
        step = _cpu_step_fn(sample_tokens, dtype)
        step()
        t = []
        for _ in range(reps):
            t0 = time.perf_counter()
            step()
            t.append(time.perf_counter() - t0)
        thr = sample_tokens / min(t)
        if thr > best:
            best, best_dt = thr, str(dtype).replace("torch.", "")
    return {"value": best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample_tokens} tokens fwd+bwd, oracle port of the reference op sequence (linear_blocktt einsum, "
                      f"left-to-right), best of {reps} after 1 warm-up, best dtype of fp32/bf16 = {best_dt}, "
                      f"torch.set_num_threads({cores})"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: tensorly is not installable, so the
    real package cannot run on the GPU box), all host threads, bounded sample per step."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample = 2048
    step = _cpu_step_fn(sample, torch.float32)
    for _ in range(max(args.warmup, 1)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    cb = {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
          "sample": f"{sample} tokens per step, fp32, oracle port (reference needs tensorly: not installable offline)"}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": _config(args.gpus, {"sample_tokens": sample}),
            "cpu_baseline": cb, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampler (NVML, polled during the timed region)
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.010)       # NVML queries contend with graph launches on the driver: keep them sparse

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _finish(world):
    """Multi-rank exit: every rank has printed / synchronised; leave without tearing the NCCL communicator down.
    (destroy_process_group() with captured CUDA graphs that hold NCCL kernels still alive was seen to hang a 2-GPU run for
    the whole box timeout after the JSON line had been printed.)"""
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import torch_b200 as tlt
    from torch_b200 import ops
    from torch_b200.parallel import GradientBucket

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch with torch.distributed.run", file=sys.stderr)

    torch.manual_seed(1234)        # identical parameters on every rank (replicated)
    up = tlt.FactorizedLinear(IN_TS, HID_TS, bias=True, factorization="blocktt", rank=TT_RANK, device=dev).to(torch.bfloat16)
    down = tlt.FactorizedLinear(HID_TS, IN_TS, bias=True, factorization="blocktt", rank=TT_RANK, device=dev).to(torch.bfloat16)
    params = list(up.parameters()) + list(down.parameters())
    torch.manual_seed(1234 + rank)  # each rank's own shard of the batch
    B = TOKENS_PER_GPU
    x = torch.randn(B, IN_F, device=dev, dtype=torch.bfloat16).requires_grad_(True)
    gy = torch.randn(B, IN_F, device=dev, dtype=torch.bfloat16)
    bucket = GradientBucket(params) if world > 1 else None

    def step():
        x.grad = None
        for p in params:
            p.grad = None
        y = down(up(x))
        y.backward(gy)
        if bucket is not None:
            bucket.allreduce()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # One step = ~26 short kernels; issued eagerly the host (Python dispatch, ~45 us per op) is slower than the GPU, so
    # the step is captured once into a CUDA graph and replayed (torch_b200.graphs.GraphedStep).  --eager disables it.
    launches0 = ops.launch_count()
    step()
    launches_per_step = ops.launch_count() - launches0
    run_step, mode = step, "eager"
    if not args.eager:
        try:
            from torch_b200.graphs import GraphedStep
            run_step, mode = GraphedStep(step).replay, "cuda_graph"
        except Exception as exc:            # e.g. a collective that cannot be captured on this stack
            if rank == 0:
                print(f"bench.py: CUDA-graph capture failed ({exc!r}); running eagerly", file=sys.stderr)
            torch.cuda.synchronize()
    for _ in range(3):
        run_step()
    barrier()

    # ------------------------------------------------------------------ timed region (device-resident inputs)
    # clocks are sampled by rank 0 only (its own GPU): eight pollers hammering NVML slowed an 8-GPU step from 0.91 to 1.11 ms
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if rank == 0:
        sampler.start()
    e0.record()
    for _ in range(args.steps):
        run_step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = launches_per_step * args.steps
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    value = world * B * args.steps / (ms / 1e3)

    # per-launch CUDA events cannot be recorded inside a graph replay: the dominant kernel is timed launch by launch in
    # an eager pass over the same steps (kernel durations do not depend on how the launch was issued)
    ops.PROFILE = []
    for _ in range(args.steps):
        step()
    barrier()
    prof, ops.PROFILE = ops.PROFILE, None

    # roofline of the dominant kernel (tcgen05 GEMM): algorithmic FLOPs / CUDA-event duration, per launch
    gemm_ms = sum(a.elapsed_time(b) for _, a, b, _, _ in prof)
    gemm_flops = sum(f for *_, f, _ in prof)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained")
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
    if peak is None:
        peak, peak_src = 1590.0, "fallback 1.59 PFLOP/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    # dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this command (profiles/)
    traffic, traffic_src = None, None
    try:
        cap = json.load(open(os.path.join(REPO, "profiles", "r01_gemm_tc3_traffic.json")))
        traffic, traffic_src = cap["avg_dram_mbytes_per_launch"] * 1e6, "profiles/r01_gemm_tc3_traffic.json: " + cap["source"]
    except Exception:
        pass
    gemm_bytes = sum(b for *_, b in prof)
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_tc3_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write)",
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": gemm_bytes / max(len(prof), 1),
                "peak_source": peak_src, "launches": len(prof),
                "avg_launch_ms": gemm_ms / max(len(prof), 1), "kernel_share_of_step": gemm_ms / ms, "step_mode": mode,
                "algorithmic_flops_per_step": gemm_flops / max(args.steps, 1)}
    per_shape = {}
    for name, a, b, f, _ in prof:
        t = per_shape.setdefault(name, [0.0, 0, f])
        t[0] += a.elapsed_time(b)
        t[1] += 1
    roofline["per_gemm"] = {k: {"avg_ms": round(v[0] / v[1], 4), "tflops": round(v[2] / (v[0] / v[1] / 1e3) / 1e12, 1)}
                            for k, v in per_shape.items()}

    if args.skip_extras:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                              "ms_per_step": ms / args.steps, "gpu_launches": launches, "roofline": roofline, "clocks": clocks}),
                  flush=True)
        _finish(world)
        return

    # ------------------------------------------------------------------ e2e: host (pinned) inputs, copies inside the timed region
    x_h = torch.randn(B, IN_F, dtype=torch.bfloat16).pin_memory()
    gy_h = torch.randn(B, IN_F, dtype=torch.bfloat16).pin_memory()
    dx_h = torch.empty(B, IN_F, dtype=torch.bfloat16).pin_memory()
    n_param = sum(p.numel() for p in params)
    gp_h = torch.empty(n_param, dtype=torch.bfloat16).pin_memory()

    # Host buffers in, host buffers out, through the public API: the three-stream front end (torch_b200.pipeline) overlaps
    # the H2D copy of step i+1 and the D2H copy of step i-1 with the kernels of step i; every copy is inside the timed region.
    from torch_b200.pipeline import HostPipelinedStep

    def outputs():
        return [x.grad, torch.cat([p.grad.reshape(-1) for p in params])]
    pipe = HostPipelinedStep([x, gy], run_step, outputs)
    for _ in range(3):
        pipe.submit([x_h, gy_h], [dx_h, gp_h])
    pipe.drain()
    barrier()
    e2e_steps = max(10, min(args.steps, 30))
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(pipe.h2d)                      # the first timed operation is an H2D copy on this stream
    for _ in range(e2e_steps):
        pipe.submit([x_h, gy_h], [dx_h, gp_h])
    pipe.d2h.wait_stream(pipe.compute)
    pipe.d2h.wait_stream(pipe.h2d)
    t1.record(pipe.d2h)                      # ... and the last one a D2H copy on this one
    pipe.drain()
    barrier()
    ms2 = torch.tensor([t0.elapsed_time(t1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e = {"value": world * B * e2e_steps / (ms2.item() / 1e3), "unit": UNIT, "steps": e2e_steps,
           "h2d_bytes_per_step": x_h.numel() * 2 + gy_h.numel() * 2, "d2h_bytes_per_step": dx_h.numel() * 2 + gp_h.numel() * 2,
           "pipeline": "3 streams: H2D(i+1) | kernels(i) | D2H(i-1); pinned host buffers"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": _config(world, {"step_mode": mode}), "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline}
    if rank == 0:
        if world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    _finish(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--eager", action="store_true", help="issue every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--skip-extras", action="store_true", help="profiling runs: skip the e2e and cpu_baseline legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
