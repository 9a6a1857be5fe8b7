"""-m gpu: the tcgen05 contraction (bess_dot_gemm) and its operand pre-pass
(bess_split_operand) through the C-ABI, against an fp64 restatement of
`torch.matmul(v1, v2.T)` (reference scoring.py:252).

Tolerances: fp32 tables run 3xTF32 (dropped product term ~2^-22): error is
bounded relative to sum_k |a_k b_k|, tested at 4e-6 of that bound (the
north_star's fp32 bar is 1e-5).  bf16 / fp16 operands are exact products of
the ROUNDED operands accumulated in fp32: compared against the fp64 product
of the same rounded operands at 1e-5, i.e. the only error left is fp32
accumulation.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

DTYPES = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}


def _imports():
    from besskge_b200 import _lib as L, kernels as K
    return L, K


def _operands(K_, L, x, gemm_dtype, transpose=False):
    """Run the pre-pass on a dense fp32 matrix x [R, W]."""
    R, W = x.shape
    dt = L.dtype_code(gemm_dtype)
    n_arr = 2 if gemm_dtype == torch.float32 else 1
    ld = (W + 7) // 8 * 8
    ldt = (R + 7) // 8 * 8
    hi = torch.zeros(R, ld, dtype=gemm_dtype, device="cuda")
    lo = torch.zeros(R, ld, dtype=gemm_dtype, device="cuda") if n_arr == 2 else None
    hit = lot = None
    if transpose:
        hit = torch.zeros(W, ldt, dtype=gemm_dtype, device="cuda")
        lot = torch.zeros(W, ldt, dtype=gemm_dtype, device="cuda") if n_arr == 2 else None
    K_.split_operand(L.F32, L.rows(x), R, W, None, dt, hi, lo, ld, hit, lot, ldt, x.device)
    return hi, lo, ld, hit, lot, ldt


def _effective(hi, lo, W):
    v = hi[:, :W].double()
    if lo is not None:
        v = v + lo[:, :W].double()
    return v


@pytest.mark.parametrize("name", ["f32", "bf16", "f16"])
def test_split_operand(name):
    L, K = _imports()
    dtype = DTYPES[name]
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(70, 100, generator=g) * torch.logspace(-3, 3, 100)).cuda()
    hi, lo, ld, hit, lot, ldt = _operands(K, L, x, dtype, transpose=True)
    torch.cuda.synchronize()
    if dtype == torch.float32:
        # both parts are exactly representable in tf32 (low 13 mantissa bits clear)
        assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0
        assert int((lo.view(torch.int32) & 0x1FFF).abs().max()) == 0
        err = (_effective(hi, lo, 100) - x.double()).abs()
        assert float((err / x.double().abs()).max()) < 2.0 ** -21
    else:
        assert torch.equal(hi[:, :100], x.to(dtype))
    assert torch.equal(hit[:, :70], hi[:, :100].t())
    if lo is not None:
        assert torch.equal(lot[:, :70], lo[:, :100].t())


def test_split_operand_rowmap_scale_and_half_source():
    L, K = _imports()
    g = torch.Generator().manual_seed(2)
    table = torch.randn(50, 24, generator=g).to(torch.bfloat16).cuda()
    idx = torch.randint(50, (30,), generator=g, dtype=torch.int32).cuda()
    scale = torch.rand(12, generator=g).cuda() + 0.5
    # logical row x -> idx[(x // 4) * 10 + x % 4 + 3]
    src = L.rows(table, idx=idx, rmap=L.rowmap(4, 10, 3))
    hi = torch.zeros(12, 24, device="cuda")
    lo = torch.zeros(12, 24, device="cuda")
    K.split_operand(L.BF16, src, 12, 24, scale, L.F32, hi, lo, 24, None, None, 0, table.device)
    torch.cuda.synchronize()
    rows = [int(idx[(x // 4) * 10 + x % 4 + 3]) for x in range(12)]
    ref = table[rows].float() * scale[:, None]
    assert float((hi + lo - ref).abs().max()) <= 2.0 ** -21 * float(ref.abs().max())


SHAPES = [
    # M, N, K
    (128, 256, 64),      # one exact tile
    (1, 1, 4),           # degenerate
    (200, 300, 100),     # tails everywhere
    (1000, 2048, 256),   # cfg-2 forward shape (reduced S)
    (2048, 256, 4096),   # few tiles, long K -> split-K
    (333, 512, 1000),
]


@pytest.mark.parametrize("aligned", [False, True])  # True: TMA-store epilogue; False: direct stores
@pytest.mark.parametrize("name", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("M,N,Kd", SHAPES)
def test_dot_gemm_matches_fp64(name, M, N, Kd, aligned):
    L, K = _imports()
    dtype = DTYPES[name]
    g = torch.Generator().manual_seed(M * 7 + N * 3 + Kd)
    a = torch.randn(M, Kd, generator=g).cuda()
    b = torch.randn(N, Kd, generator=g).cuda()
    a_hi, a_lo, lda, *_ = _operands(K, L, a, dtype)
    b_hi, b_lo, ldb, *_ = _operands(K, L, b, dtype)
    c0 = 4 if aligned else 1
    ld_out = (N + 11) // 4 * 4 if aligned else N + 5
    out = torch.full((M, ld_out), 7.0, device="cuda")
    ws_bytes = K.dot_gemm_workspace(M, N, Kd)
    ws = torch.empty(max(ws_bytes // 4, 1), device="cuda")
    K.dot_gemm(L.dtype_code(dtype), a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, Kd, out, L.IDENT, ld_out,
               c0, False, ws)
    # the same product with A handed over transposed ([K, M], MN-major UMMA descriptors)
    at_hi, at_lo, ldat, *_ = _operands(K, L, a.t().contiguous(), dtype)
    out_mn = torch.full((M, ld_out), 7.0, device="cuda")
    K.dot_gemm(L.dtype_code(dtype), at_hi, at_lo, ldat, b_hi, b_lo, ldb, M, N, Kd, out_mn, L.IDENT,
               ld_out, c0, False, ws, a_mn_major=True)
    torch.cuda.synchronize()
    assert torch.all(out[:, :c0] == 7.0) and torch.all(out[:, N + c0:] == 7.0)  # untouched columns
    assert torch.equal(out, out_mn)  # same operand values, same accumulation order
    got = out[:, c0:N + c0].double()
    if dtype == torch.float32:
        ref = a.double() @ b.double().t()
        bound = a.double().abs() @ b.double().abs().t()
        assert float(((got - ref).abs() / bound).max()) < 4e-6
    else:
        ea, eb = _effective(a_hi, None, Kd), _effective(b_hi, None, Kd)
        ref = ea @ eb.t()
        bound = ea.abs() @ eb.abs().t()
        assert float(((got - ref).abs() / bound).max()) < 1e-5


def test_dot_gemm_rowmap_accumulate_and_determinism():
    L, K = _imports()
    g = torch.Generator().manual_seed(5)
    M, N, Kd = 96, 40, 3000
    a = torch.randn(M, Kd, generator=g).cuda()
    b = torch.randn(N, Kd, generator=g).cuda()
    a_hi, a_lo, lda, *_ = _operands(K, L, a, torch.float32)
    b_hi, b_lo, ldb, *_ = _operands(K, L, b, torch.float32)
    # logical row m -> storage row (m // 32) * 50 + m % 32 + 2
    omap = L.rowmap(32, 50, 2)
    base = torch.randn(160, N, generator=g).cuda()
    ws = torch.empty(max(K.dot_gemm_workspace(M, N, Kd) // 4, 1), device="cuda")
    outs = []
    for _ in range(2):
        out = base.clone()
        K.dot_gemm(L.F32, a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, Kd, out, omap, N, 0, True, ws)
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.equal(outs[0], outs[1])  # split-K partials are reduced in a fixed order
    rows = torch.tensor([(m // 32) * 50 + m % 32 + 2 for m in range(M)], device="cuda")
    ref = base.double()
    ref[rows] += a.double() @ b.double().t()
    bound = (a.double().abs() @ b.double().abs().t()).max()
    assert float((outs[0].double() - ref).abs().max() / bound) < 4e-6
    untouched = torch.ones(160, dtype=torch.bool, device="cuda")
    untouched[rows] = False
    assert torch.equal(outs[0][untouched], base[untouched])
