"""Microbenchmark of the tcgen05 GEMM on the dense contractions of the C2 step (plain python, one GPU):

    python tests/bench_gemm_shapes.py [precision]          # all tile widths, one subprocess each

Per shape and tile width (EESEN_B200_GEMM_BN): max error against torch fp64, ms per call, algorithmic TFLOP/s.
Used to set the width heuristic in gemm_tc.cu:pick_bn."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, transA, transB, M, N, K, beta)   rows of a C2 minibatch: 600 x 64 = 38400
SHAPES = [
    ("fwd  x*Wx^T  L2-4", 0, 1, 38400, 1280, 640, 0.0),
    ("fwd  x*Wx^T  L1  ", 0, 1, 38400, 1280, 40, 0.0),
    ("dX   DG*Wx       ", 0, 0, 38400, 640, 1280, 1.0),
    ("dWx  DG^T*x      ", 1, 0, 1280, 640, 38400, 0.0),
    ("dWm  DG^T*m      ", 1, 0, 1280, 320, 38336, 0.0),
    ("dX as NT (Wx^T)  ", 0, 1, 38400, 640, 1280, 1.0),
    ("square 4096      ", 0, 1, 4096, 4096, 4096, 0.0),
]


def child(prec):
    import torch
    from eesen_b200 import binding
    ctx = binding.Context(0)
    ctx.set_precision(prec, "fp32x3")
    st = torch.cuda.ExternalStream(ctx.stream)
    g = torch.Generator(device="cuda").manual_seed(0)
    only = os.environ.get("BENCH_ONLY", "")
    for name, ta, tb, M, N, K, beta in SHAPES:
        if only and only not in name:
            continue
        A = torch.randn((K, M) if ta else (M, K), device="cuda", generator=g)
        B = torch.randn((N, K) if tb else (K, N), device="cuda", generator=g)
        C0 = torch.randn(M, N, device="cuda", generator=g)
        Cm = C0.clone()
        torch.cuda.synchronize()
        ctx.gemm(ta, tb, M, N, K, 1.0, A, A.shape[1], B, B.shape[1], beta, Cm, N)
        ctx.synchronize()
        # spot-check 64 random rows in fp64
        rows = torch.randint(0, M, (64,), device="cuda")
        a = (A.t() if ta else A)[rows].double()
        b = (B.t() if tb else B).double() if tb else B.double()
        ref = a @ (B.double().t() if tb else B.double()) + beta * C0[rows].double()
        err = (Cm[rows].double() - ref).abs().max().item() / max(1.0, ref.abs().max().item())
        for _ in range(3):
            ctx.gemm(ta, tb, M, N, K, 1.0, A, A.shape[1], B, B.shape[1], 0.0, Cm, N)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.synchronize()
        e0.record(st)
        reps = 10
        for _ in range(reps):
            ctx.gemm(ta, tb, M, N, K, 1.0, A, A.shape[1], B, B.shape[1], beta, Cm, N)
        e1.record(st)
        ctx.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"  {name} {M:6d}x{N:5d}x{K:6d}  rel.err {err:.2e}  {ms:8.3f} ms  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(sys.argv[2])
        sys.exit(0)
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32x3"
    if len(sys.argv) > 2 and sys.argv[2] == "splits":      # split-K factor sweep on the weight-gradient shapes
        for sp in ("0", "6", "9", "12", "14", "17", "20", "26"):
            env = dict(os.environ, BENCH_ONLY="dW")
            if sp != "0":
                env["EESEN_B200_GEMM_SPLITS"] = sp
            print(f"== split-K factor {sp if sp != '0' else 'auto'}, arithmetic {prec}", flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", prec], env=env, check=False, timeout=600)
        sys.exit(0)
    for bn in ("128", "256", "auto"):
        env = dict(os.environ)
        if bn != "auto":
            env["EESEN_B200_GEMM_BN"] = bn
        else:
            env.pop("EESEN_B200_GEMM_BN", None)
        print(f"== tile width {bn}, arithmetic {prec}", flush=True)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", prec], env=env, check=False, timeout=600)
