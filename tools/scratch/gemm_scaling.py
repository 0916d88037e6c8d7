"""GEMM time against M at fixed (N, K): fixed cost of a launch vs cost per tile (development aid)."""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")))
sys.argv = [sys.argv[0]]
import torch
import importlib.util
spec = importlib.util.spec_from_file_location("gemm_bench", os.path.join(os.path.dirname(__file__), "..", "gemm_bench.py"))
gb = importlib.util.module_from_spec(spec)
sys.argv = [sys.argv[0], "--none"]
spec.loader.exec_module(gb)
bf, f32 = torch.bfloat16, torch.float32
for N, K, base in ((576, 512, 12544), (144, 128, 200704), (288, 256, 50176)):
    for div in (16, 8, 4, 2, 1):
        gb.run(base // div, N, K, 0, 0, bf, 0, iters=10)
for M, N, base in ((512, 576, 12544), (128, 144, 200704), (256, 288, 50176)):
    for div in (16, 4, 1):
        gb.run(M, N, base // div, 1, 1, f32, 0, iters=10)
