"""Times bess_dot_gemm alone (CUDA events) at the cfg-2 contraction shapes, for the three
operand formats of fp32 tables / half tables: tf32 pairs (3xTF32), scaled fp16 pairs (3xFP16),
plain bf16."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from besskge_b200 import _lib as L, kernels as K  # noqa: E402
from besskge_b200.bess import _TcOperand  # noqa: E402

FORMATS = {"tf32x3": (torch.float32, L.F32), "f16x3": (torch.float32, L.F16X3),
           "bf16": (torch.bfloat16, L.BF16)}


def run(name, fmt_name, M, N, Kd, iters=20, mn=False):
    dtype, fmt = FORMATS[fmt_name]
    ws = K.Workspace(torch.device("cuda"))
    a = torch.randn(M, Kd, device="cuda")
    b = torch.randn(N, Kd, device="cuda")
    a_src = a.t().contiguous() if mn else a
    a_op = _TcOperand(ws, "a", a_src.shape[0], a_src.shape[1], dtype, False, fmt)
    a_op.fill(L.F32, L.rows(a_src), 0, None, a.device)
    b_op = _TcOperand(ws, "b", N, Kd, dtype, False, fmt)
    b_op.fill(L.F32, L.rows(b), 0, None, a.device)
    out = torch.empty(M, N, device="cuda")
    gws = torch.empty(max(K.dot_gemm_workspace(M, N, Kd) // 4, 1), device="cuda")

    def go():
        K.dot_gemm(fmt, a_op.hi, a_op.lo, a_op.ld, b_op.hi, b_op.lo, b_op.ld, M, N, Kd, out, L.IDENT, N, 0,
                   False, gws, a_mn_major=mn, a_scale=a_op.scale, b_scale=b_op.scale)
    for _ in range(3):
        go()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        go()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    tf = 2.0 * M * N * Kd / us / 1e6
    ref = (a.double() @ b.double().t()) if M * N <= 1 << 24 else None
    err = float((out.double() - ref).abs().max()) if ref is not None else float("nan")
    print(f"{name:22s} {fmt_name:7s} M={M:6d} N={N:5d} K={Kd:6d}: {us:9.1f} us  {tf:8.1f} TFLOP/s (algorithmic)"
          f"  maxerr {err:.2e}", flush=True)


if __name__ == "__main__":
    S = 16384
    if "--one" in sys.argv:  # short run for an ncu capture
        fmt = sys.argv[sys.argv.index("--one") + 1] if len(sys.argv) > sys.argv.index("--one") + 1 else "f16x3"
        run("fwd  scores = Q C^T", fmt, S, 2048, 256, iters=2)
        run("bwd  dQ = dS C", fmt, S, 256, 2048, iters=2)
        run("bwd  dC = dS^T Q", fmt, 2048, 256, S, iters=2, mn=True)
        sys.exit(0)
    for fmt in FORMATS:
        run("fwd  scores = Q C^T", fmt, S, 2048, 256)
        run("bwd  dQ = dS C", fmt, S, 256, 2048)
        run("bwd  dC = dS^T Q", fmt, 2048, 256, S, mn=True)
        run("small", fmt, 1024, 1024, 256)
