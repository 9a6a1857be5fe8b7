"""Times bess_dot_gemm alone (CUDA events) at the cfg-2 contraction shapes."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from besskge_b200 import _lib as L, kernels as K  # noqa: E402


def operands(x, dtype):
    R, W = x.shape
    hi = torch.empty(R, W, dtype=dtype, device="cuda")
    lo = torch.empty(R, W, dtype=dtype, device="cuda") if dtype == torch.float32 else None
    K.split_operand(L.F32, L.rows(x), R, W, None, L.dtype_code(dtype), hi, lo, W, None, None, 0, x.device)
    return hi, lo


def run(name, dtype, M, N, Kd, iters=20):
    a = torch.randn(M, Kd, device="cuda")
    b = torch.randn(N, Kd, device="cuda")
    a_hi, a_lo = operands(a, dtype)
    b_hi, b_lo = operands(b, dtype)
    out = torch.empty(M, N, device="cuda")
    ws = torch.empty(max(K.dot_gemm_workspace(M, N, Kd) // 4, 1), device="cuda")
    dt = L.dtype_code(dtype)
    for _ in range(3):
        K.dot_gemm(dt, a_hi, a_lo, Kd, b_hi, b_lo, Kd, M, N, Kd, out, L.IDENT, N, 0, False, ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        K.dot_gemm(dt, a_hi, a_lo, Kd, b_hi, b_lo, Kd, M, N, Kd, out, L.IDENT, N, 0, False, ws)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    tf = 2.0 * M * N * Kd / us / 1e6
    ref = (a.double() @ b.double().t()) if M * N <= 1 << 24 else None
    err = float((out.double() - ref).abs().max()) if ref is not None else float("nan")
    print(f"{name:28s} M={M:6d} N={N:5d} K={Kd:6d} {dtype}: {us:9.1f} us  {tf:8.1f} TFLOP/s (algorithmic)  maxerr {err:.2e}",
          flush=True)


if __name__ == "__main__":
    S = 16384
    if "--one" in sys.argv:  # short run for an ncu capture
        run("fwd  scores = Q C^T", torch.float32, S, 2048, 256, iters=2)
        run("bwd  dC = dS^T Q", torch.float32, 2048, 256, S, iters=2)
        sys.exit(0)
    for dtype in (torch.float32, torch.bfloat16):
        run("fwd  scores = Q C^T", dtype, S, 2048, 256)
        run("bwd  dQ = dS C", dtype, S, 256, 2048)
        run("bwd  dC = dS^T Q", dtype, 2048, 256, S)
        run("small", dtype, 1024, 1024, 256)
