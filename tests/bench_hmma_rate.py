"""Microbenchmark: legacy warp-level tensor path, HMMA.1688.TF32 vs HMMA.16816.BF16 (gemm.cu, mma.sync)."""
import os, sys, time
os.environ["EESEN_B200_GEMM_ENGINE"] = "legacy"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eesen_b200 import binding
ctx = binding.Context(0)
M = N = K = 4096
A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.zeros(M, N, device="cuda")
torch.cuda.synchronize()
st = torch.cuda.ExternalStream(ctx.stream)
for prec in ("tf32", "bf16", "fp32x3"):
    ctx.set_precision(prec, "fp32x3")
    for _ in range(3):
        ctx.gemm(0, 1, M, N, K, 1.0, A, K, B, K, 0.0, C, N)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.synchronize()
    e0.record(st)
    for _ in range(10):
        ctx.gemm(0, 1, M, N, K, 1.0, A, K, B, K, 0.0, C, N)
    e1.record(st)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"legacy mma.sync {prec:7s}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s algorithmic")
