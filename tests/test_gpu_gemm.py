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


# ----------------------------------------------------------------------------
# 3xFP16 (BESS_F16X3): fp32 operands as scaled fp16 (hi, lo) pairs
# ----------------------------------------------------------------------------
def _operands_x3(K_, L, x, transpose=False, row_scale=None):
    R, W = x.shape
    ld, ldt = (W + 7) // 8 * 8, (R + 7) // 8 * 8
    hi = torch.zeros(R, ld, dtype=torch.float16, device="cuda")
    lo = torch.zeros_like(hi)
    hit = lot = None
    if transpose:
        hit = torch.zeros(W, ldt, dtype=torch.float16, device="cuda")
        lot = torch.zeros_like(hit)
    scale = torch.zeros(2, device="cuda")
    state = torch.zeros(2, dtype=torch.int32, device="cuda")
    K_.operand_scale(L.F32, L.rows(x), R, W, row_scale, 1.0, scale, state)
    K_.split_operand(L.F32, L.rows(x), R, W, row_scale, L.F16X3, hi, lo, ld, hit, lot, ldt, x.device,
                     scale)
    return hi, lo, ld, hit, lot, ldt, scale, state


@pytest.mark.parametrize("magnitude", [1.0, 1.0 / 256, 3e-5, 4e4])
def test_split_operand_f16x3_scale_and_precision(magnitude):
    """scale = power of two with the largest scaled element in [2^13, 2^14]; hi + lo
    reproduces x * s to 2^-21 relative for everything within ~2^20 of the largest element —
    whatever the operand's own magnitude (default-initialised tables are ~1/256, products of
    two of them ~1e-5: far below fp16's normal range without the scale)."""
    L, K = _imports()
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(70, 100, generator=g) * magnitude).cuda()
    x[5, 7] = 0.0
    hi, lo, ld, hit, lot, ldt, scale, state = _operands_x3(K, L, x, transpose=True)
    torch.cuda.synchronize()
    s, inv = float(scale[0]), float(scale[1])
    assert s * inv == 1.0 and s == 2.0 ** round(torch.log2(scale[0]).item())
    top = float(x.abs().max()) * s
    assert 2.0 ** 13 <= top <= 2.0 ** 14
    assert int(state.abs().max()) == 0  # the kernel leaves its state words zero
    eff = (hi[:, :100].double() + lo[:, :100].double()) * inv
    big = x.abs() > float(x.abs().max()) * 2.0 ** -18
    assert float(((eff - x.double()).abs() / x.double().abs().clamp_min(1e-300))[big].max()) < 2.0 ** -21
    # every element, however small: 2^-21 of itself + ~2^-37 of the largest one (fp16 subnormal
    # spacing 2^-24 against a largest scaled element >= 2^13)
    err = (eff - x.double()).abs()
    assert bool((err <= x.double().abs() * 2.0 ** -21 + float(x.abs().max()) * 2.0 ** -36).all())
    assert torch.equal(hit[:, :70], hi[:, :100].t()) and torch.equal(lot[:, :70], lo[:, :100].t())


@pytest.mark.parametrize("aligned", [False, True])
@pytest.mark.parametrize("scales", [(1.0, 1.0), (1.0 / 256, 1.0 / 256), (2e-5, 0.5), (3e3, 1e-3)])
@pytest.mark.parametrize("M,N,Kd", SHAPES)
def test_dot_gemm_f16x3_matches_fp64(M, N, Kd, scales, aligned):
    """Same bound as 3xTF32 (4e-6 of sum |a_k b_k|), for operands of very different magnitudes,
    K-major and MN-major A, TMA-store and direct-store epilogues, split-K shapes."""
    L, K = _imports()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + Kd)
    a = (torch.randn(M, Kd, generator=g) * scales[0]).cuda()
    b = (torch.randn(N, Kd, generator=g) * scales[1]).cuda()
    # a wide dynamic range inside one operand as well (rows spread over 4 decades)
    a *= torch.logspace(-2, 2, M).cuda()[:, None]
    a_hi, a_lo, lda, _, _, _, sa, _ = _operands_x3(K, L, a)
    b_hi, b_lo, ldb, _, _, _, sb, _ = _operands_x3(K, L, b)
    c0 = 4 if aligned else 1
    ld_out = (N + 11) // 4 * 4 if aligned else N + 5
    out = torch.full((M, ld_out), 7.0, device="cuda")
    ws = torch.empty(max(K.dot_gemm_workspace(M, N, Kd) // 4, 1), device="cuda")
    K.dot_gemm(L.F16X3, a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, Kd, out, L.IDENT, ld_out, c0, False,
               ws, a_scale=sa, b_scale=sb)
    at_hi, at_lo, ldat, _, _, _, sat, _ = _operands_x3(K, L, a.t().contiguous())
    out_mn = torch.full((M, ld_out), 7.0, device="cuda")
    K.dot_gemm(L.F16X3, at_hi, at_lo, ldat, b_hi, b_lo, ldb, M, N, Kd, out_mn, L.IDENT, ld_out, c0,
               False, ws, a_mn_major=True, a_scale=sat, b_scale=sb)
    torch.cuda.synchronize()
    assert torch.all(out[:, :c0] == 7.0) and torch.all(out[:, N + c0:] == 7.0)
    assert torch.equal(out, out_mn)
    got = out[:, c0:N + c0].double()
    ref = a.double() @ b.double().t()
    bound = a.double().abs() @ b.double().abs().t()
    assert float(((got - ref).abs() / bound).max()) < 4e-6


def test_dot_gemm_f16x3_vs_tf32x3_same_grade():
    """The two split formats deliver the same accuracy class on the cfg-2 shape."""
    L, K = _imports()
    g = torch.Generator().manual_seed(11)
    M, N, Kd = 2048, 2048, 256
    a = (torch.randn(M, Kd, generator=g) * 0.3).cuda()
    b = (torch.randn(N, Kd, generator=g) * 0.3).cuda()
    ref = a.double() @ b.double().t()
    bound = a.double().abs() @ b.double().abs().t()
    ws = torch.empty(max(K.dot_gemm_workspace(M, N, Kd) // 4, 1), device="cuda")
    errs = {}
    a_hi, a_lo, lda, *_ = _operands(K, L, a, torch.float32)
    b_hi, b_lo, ldb, *_ = _operands(K, L, b, torch.float32)
    out = torch.empty(M, N, device="cuda")
    K.dot_gemm(L.F32, a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, Kd, out, L.IDENT, N, 0, False, ws)
    errs["tf32x3"] = float(((out.double() - ref).abs() / bound).max())
    a_hi, a_lo, lda, _, _, _, sa, _ = _operands_x3(K, L, a)
    b_hi, b_lo, ldb, _, _, _, sb, _ = _operands_x3(K, L, b)
    K.dot_gemm(L.F16X3, a_hi, a_lo, lda, b_hi, b_lo, ldb, M, N, Kd, out, L.IDENT, N, 0, False, ws,
               a_scale=sa, b_scale=sb)
    errs["f16x3"] = float(((out.double() - ref).abs() / bound).max())
    print(errs)
    assert errs["f16x3"] < 4e-6 and errs["tf32x3"] < 4e-6
    assert errs["f16x3"] < 4 * errs["tf32x3"] + 1e-7
