"""-m gpu: individual CUDA kernels through the C-ABI vs the oracle / reference goldens."""
import numpy as np
import pytest
import torch
from numpy.testing import assert_array_equal
from torch.testing import assert_close

from oracle import besskge_oracle as O

from .conftest import load_golden

pytestmark = pytest.mark.gpu

FAMS = ["TransE", "RotatE", "DistMult", "ComplEx", "PairRE", "BoxE", "TripleRE", "InterHT", "TranS"]


def _imports():
    from besskge_b200 import _lib as L, kernels as K
    from . import gpu_helpers as H
    return L, K, H


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("W", [16, 64, 256, 1000])
def test_gather_route_bit_exact(dtype, W):
    L, K, H = _imports()
    if (W * torch.finfo(dtype).bits // 8) % 16:
        W = W + 8 - W % 8
    g = torch.Generator().manual_seed(0)
    Es, n, n_local, per = 3000, 4, 37, 53
    table = torch.randn(Es, W, generator=g).to(dtype).cuda()
    idx = torch.randint(Es, (n_local + n * per,), generator=g, dtype=torch.int32).cuda()
    local = torch.zeros(n_local, W, dtype=dtype, device="cuda")
    recv = torch.zeros(n, 3, per, W, dtype=dtype, device="cuda")  # [dst][slot][per]
    slot = 2
    K.gather_route(table, idx, n_local, per, local, [recv[j].data_ptr() for j in range(n)], slot)
    torch.cuda.synchronize()
    ref = table[idx.long()]
    assert torch.equal(local, ref[:n_local])
    assert torch.equal(recv[:, slot], ref[n_local:].view(n, per, W))
    assert torch.count_nonzero(recv[:, 0]) == 0 and torch.count_nonzero(recv[:, 1]) == 0
    out = torch.empty(idx.numel(), W, dtype=dtype, device="cuda")
    K.gather_rows(table, idx, out)
    assert torch.equal(out, ref)


def test_gather_empty_and_alignment_error():
    L, K, H = _imports()
    table = torch.randn(10, 16, device="cuda")
    out = torch.empty(0, 16, device="cuda")
    K.gather_rows(table, torch.empty(0, dtype=torch.int32, device="cuda"), out)  # no-op
    bad = torch.randn(10, 6, device="cuda")  # 24-byte rows
    with pytest.raises(L.BessLibraryError):
        K.gather_rows(bad, torch.zeros(2, dtype=torch.int32, device="cuda"),
                      torch.empty(2, 6, device="cuda"))


@pytest.mark.parametrize("n,bits", [(1, 5), (31, 3), (2048, 11), (2049, 17), (100_003, 22),
                                    (300_000, 8)])
def test_radix_sort_stable(n, bits):
    L, K, H = _imports()
    rng = np.random.default_rng(n)
    keys = rng.integers(1 << bits, size=n).astype(np.int32)
    if n > 1000:
        keys[: n // 3] = keys[0]  # a heavy hitter
    kd = torch.from_numpy(keys).cuda()
    ko = torch.empty_like(kd)
    po = torch.empty_like(kd)
    ws = torch.empty(K.sort_workspace(n) // 4 + 64, dtype=torch.int32, device="cuda")
    K.sort_keys(kd, n, bits, ko, po, ws)
    torch.cuda.synchronize()
    order = np.argsort(keys, kind="stable")
    assert_array_equal(po.cpu().numpy(), order.astype(np.int32))
    assert_array_equal(ko.cpu().numpy(), keys[order])


@pytest.mark.parametrize("fam", FAMS)
def test_score_functions_vs_reference_golden(fam):
    L, K, H = _imports()
    from besskge_b200.sharding import Sharding
    cfg, g = load_golden(f"scores_{fam}")
    d = cfg["d"]
    sh = Sharding.create(60, 1, seed=1234)
    ent, rel = H.T(g["ent"]), H.T(g["rel"])
    h, t, r = H.T(g["h"]).cuda(), H.T(g["t"]).cuda(), H.T(g["r"]).cuda()
    for vi, v in enumerate(cfg["variants"]):
        kw = {}
        if "normalize_entities" in v:
            kw["normalize_entities"] = v["normalize_entities"]
        if "apply_tanh" in v:
            kw["apply_tanh"] = v["apply_tanh"]
        if "dist_func_per_dim" in v:
            kw["dist_func_per_dim"] = v["dist_func_per_dim"]
        if "u" in v:
            kw["u"] = v["u"]
        if "offset" in v:
            kw["offset"] = v["offset"]
        for sharing in (True, False):
            sf = H.make_score_fn(fam, sharing, v["p"], sh, cfg["n_rel"], d, ent, rel, **kw)
            assert_close(sf.score_triple(h, r, t).cpu(), H.T(g[f"v{vi}_triple"]), rtol=1e-5,
                         atol=2e-5)
            key = f"v{vi}_s{int(sharing)}_heads"
            if key not in g:
                continue
            cand = H.T(g["c_shared"] if sharing else g["c_per"]).cuda()
            assert_close(sf.score_heads(cand, r, t).cpu(), H.T(g[key]), rtol=1e-5, atol=2e-5)
            assert_close(sf.score_tails(h, r, cand).cpu(), H.T(g[f"v{vi}_s{int(sharing)}_tails"]),
                         rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("fam,p", [("TransE", 1), ("TransE", 2), ("RotatE", 1), ("DistMult", 2),
                                   ("ComplEx", 2), ("PairRE", 2), ("BoxE", 1), ("TripleRE", 1),
                                   ("InterHT", 1), ("InterHT", 2), ("TranS", 1), ("TranS", 2)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_score_shared_tiles_vs_oracle(fam, p, dtype):
    """multi-tile shapes (ragged edges) in every table dtype; 1e-5 (fp32) / 1e-2 (half)."""
    L, K, H = _imports()
    from besskge_b200.sharding import Sharding
    d, nq, nc, n_rel = 48, 300, 517, 9
    sh = Sharding.create(64, 1, seed=1)
    g = torch.Generator().manual_seed(5)
    ew = 2 if fam in ("RotatE", "ComplEx", "BoxE", "InterHT", "TranS") else 1
    rw = {"TransE": d, "RotatE": d, "DistMult": d, "ComplEx": 2 * d, "PairRE": 2 * d,
          "BoxE": 4 * d + 2, "TripleRE": 3 * d, "InterHT": d, "TranS": 3 * d}[fam]
    ent = torch.randn(1, 64, ew * d, generator=g)
    rel = (torch.randn(n_rel, rw, generator=g)).to(dtype).float()
    h = torch.randn(nq, ew * d, generator=g).to(dtype)
    cand = torch.randn(1, nc, ew * d, generator=g).to(dtype)
    r = torch.randint(n_rel, (nq,), generator=g)
    sf = H.make_score_fn(fam, True, p, sh, n_rel, d, ent, rel, dtype=dtype)
    oc = H.score_cfg(fam, d, p)
    tol = dict(rtol=1e-5, atol=1e-4) if dtype == torch.float32 else dict(rtol=1e-2, atol=5e-2)
    for mode in ("t", "h"):
        want = O.score_candidates(oc, mode, h.float(), rel, r, cand.float(), True)
        got = (sf.score_tails(h.cuda(), r.cuda(), cand.cuda()) if mode == "t"
               else sf.score_heads(cand.cuda(), r.cuda(), h.cuda()))
        assert got.dtype == dtype
        assert_close(got.float().cpu(), want, **tol)


def test_loss_vs_reference_golden():
    L, K, H = _imports()
    cfg, g = load_golden("loss")
    pos, neg = H.T(g["pos"]).cuda(), H.T(g["neg"]).cuda()
    for i, case in enumerate(cfg["cases"]):
        fn = H.make_loss(case)
        for wi, w in enumerate((H.T(g["w"]).cuda(), torch.tensor([1.0], device="cuda"))):
            lossv, dpos, dneg = fn.fwd_bwd(pos.clone(), neg.clone(), w)
            assert_close(lossv.cpu(), H.T(g[f"c{i}_w{wi}_loss"]), rtol=1e-5, atol=1e-5)
            assert_close(dpos.cpu(), H.T(g[f"c{i}_w{wi}_dpos"]), rtol=1e-5, atol=1e-6)
            assert_close(dneg.cpu(), H.T(g[f"c{i}_w{wi}_dneg"]), rtol=1e-5, atol=1e-6)


def test_loss_wide_rows_vs_oracle():
    L, K, H = _imports()
    g = torch.Generator().manual_seed(9)
    S, N = 70, 3001
    pos = torch.randn(S, generator=g) * 4
    neg = torch.randn(S, N, generator=g) * 4
    w = torch.rand(S, generator=g)
    for case in (dict(kind="logsigmoid", margin=6.0, negative_adversarial_sampling=True),
                 dict(kind="margin_ranking", margin=1.0, negative_adversarial_sampling=True,
                      negative_adversarial_scale=0.3),
                 dict(kind="softmax_ce", n_entity=93773)):
        p_, n_ = pos.clone().requires_grad_(True), neg.clone().requires_grad_(True)
        want = O.loss_value(H.oracle_loss_cfg(case), p_, n_, w)
        want.backward()
        lossv, dpos, dneg = H.make_loss(case).fwd_bwd(pos.cuda(), neg.clone().cuda(), w.cuda())
        assert_close(lossv.cpu(), want.detach(), rtol=1e-5, atol=1e-4)
        assert_close(dpos.cpu(), p_.grad, rtol=1e-4, atol=1e-6)
        assert_close(dneg.cpu(), n_.grad, rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("N", [1024, 2048, 4096])
@pytest.mark.parametrize("adversarial", [True, False])
def test_loss_full_rows_fast_path_vs_oracle(N, adversarial):
    """Rows of exactly 16 scores per thread take the predicate-free LogSigmoid path
    (approx SFU ops, folded constants): same tolerances as the generic path."""
    L, K, H = _imports()
    g = torch.Generator().manual_seed(N)
    S = 150
    pos = torch.randn(S, generator=g) * 4
    neg = torch.randn(S, N, generator=g) * 4
    w = torch.rand(S, generator=g)
    case = dict(kind="logsigmoid", margin=6.0, negative_adversarial_sampling=adversarial,
                negative_adversarial_scale=0.7, loss_scale=1.5)
    p_, n_ = pos.clone().requires_grad_(True), neg.clone().requires_grad_(True)
    want = O.loss_value(H.oracle_loss_cfg(case), p_, n_, w)
    want.backward()
    lossv, dpos, dneg = H.make_loss(case).fwd_bwd(pos.cuda(), neg.clone().cuda(), w.cuda())
    assert_close(lossv.cpu(), want.detach(), rtol=1e-5, atol=1e-4)
    assert_close(dpos.cpu(), p_.grad, rtol=1e-4, atol=1e-6)
    assert_close(dneg.cpu(), n_.grad, rtol=1e-4, atol=1e-7)


def test_metrics_vs_reference_golden():
    L, K, H = _imports()
    from besskge_b200.metric import Evaluation
    _, g = load_golden("metric")
    pos, neg = H.T(g["pos"]).cuda(), H.T(g["neg"]).cuda()
    for mode in ("optimistic", "pessimistic", "average"):
        for winf in (False, True):
            ev = Evaluation(["mrr", "hits@1", "hits@5"], mode=mode, worst_rank_infty=winf)
            rk = ev.ranks_from_scores(pos, neg)
            assert_close(rk.cpu(), H.T(g[f"rank_{mode}_{int(winf)}"]))
            for k, v in ev.dict_metrics_from_ranks(rk).items():
                assert_close(v.cpu(), H.T(g[f"m_{mode}_{int(winf)}_{k}"]))
    for winf in (False, True):
        ev = Evaluation(["mrr"], worst_rank_infty=winf)
        assert_close(ev.ranks_from_indices(H.T(g["truth"]).cuda(), H.T(g["ids"]).cuda()).cpu(),
                     H.T(g[f"idrank_{int(winf)}"]))
    # reference's hand-written vectors (tests/test_metric.py:13-49)
    pos = torch.tensor([2.1, 5.0, 5.9, 2.0]).cuda()
    neg = torch.tensor([[2.1, 3.1, 2.1, 5.2, 8.4], [9.8, 5.0, 1.0, 3.2, 5.0],
                        [4.0, 2.3, 5.9, 3.1, 4.5], [4.0, 2.3, 5.9, 3.1, 4.5]]).cuda()
    ev = Evaluation(["mrr", "hits@1", "hits@5"], mode="pessimistic", worst_rank_infty=True)
    res = ev.dict_metrics_from_ranks(ev.ranks_from_scores(pos, neg))
    assert_close(res["hits@5"].cpu(), torch.tensor([0.0, 1.0, 1.0, 0.0]))
    assert_close(res["mrr"].cpu(), torch.tensor([0.0, 1 / 4, 1 / 2, 0.0]))


@pytest.mark.parametrize("kind", ["logsigmoid", "margin_ranking", "softmax_ce"])
@pytest.mark.parametrize("gdt", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("n_neg", [37, 2048, 5000])
def test_loss_operand_gradient_matches_plain(kind, gdt, n_neg):
    """bess_loss_fwd_bwd_operand = bess_loss_fwd_bwd with the gradient written as GEMM
    operand arrays (tf32 hi/lo or rounded halves)."""
    L, K, H = _imports()
    g = torch.Generator().manual_seed(n_neg)
    n = 33
    pos = torch.randn(n, generator=g).cuda() * 3
    neg = (torch.randn(n, n_neg, generator=g) * 3).cuda()
    w = (torch.rand(n, generator=g) + 0.5).cuda()
    kd = {"logsigmoid": L.LOSS_LOGSIGMOID, "margin_ranking": L.LOSS_MARGIN_RANKING,
          "softmax_ce": L.LOSS_SOFTMAX_CE}[kind]
    outs = []
    for operand in (False, True):
        neg_c = neg.clone()
        row_loss = torch.empty(n, device="cuda")
        d_pos = torch.empty(n, device="cuda")
        if operand:
            ld = (n_neg + 7) // 8 * 8
            hi = torch.zeros(n, ld, dtype=gdt, device="cuda")
            lo = torch.zeros(n, ld, dtype=gdt, device="cuda") if gdt == torch.float32 else None
            K.loss_fwd_bwd_operand(kd, 3.0, True, 0.7, 1.5, 1000, pos, neg_c, n, n_neg, n_neg, w,
                                   row_loss, d_pos, L.dtype_code(gdt), hi, lo, ld)
            grad = hi[:, :n_neg].float() + (lo[:, :n_neg] if lo is not None else 0)
            if lo is not None:
                assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0
        else:
            grad = torch.empty(n, n_neg, device="cuda")
            K.loss_fwd_bwd(kd, 3.0, True, 0.7, 1.5, 1000, pos, neg_c, n, n_neg, n_neg, w, row_loss,
                           d_pos, grad)
        torch.cuda.synchronize()
        outs.append((row_loss, d_pos, grad, neg_c))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][3], outs[1][3])
    tol = {torch.float32: 2.0 ** -21, torch.bfloat16: 2.0 ** -8, torch.float16: 2.0 ** -10}[gdt]
    ref, got = outs[0][2], outs[1][2]
    big = ref.abs() > 1e-30 if gdt != torch.float16 else ref.abs() > 1e-4
    assert float(((got - ref).abs()[big] / ref.abs()[big]).max()) <= tol


@pytest.mark.parametrize("kind", ["logsigmoid", "margin_ranking", "softmax_ce"])
@pytest.mark.parametrize("n_neg", [1, 37, 2048, 5000])
def test_loss_kernel_vs_oracle_long_rows(kind, n_neg):
    """Register-cached loss kernel (rows longer than the 4096-score cache re-read the tail)."""
    L, K, H = _imports()
    g = torch.Generator().manual_seed(7 + n_neg)
    n = 19
    pos = torch.randn(n, generator=g) * 2
    neg = torch.randn(n, n_neg, generator=g) * 2
    w = torch.rand(n, generator=g) + 0.5
    cfg = dict(kind=kind, margin=2.0, adversarial=True, adv_scale=1.3, loss_scale=1.0, n_entity=777)
    negr = neg.clone().requires_grad_(True)
    posr = pos.clone().requires_grad_(True)
    want = O.loss_value(cfg, posr, negr, w)
    want.backward()
    kd = {"logsigmoid": L.LOSS_LOGSIGMOID, "margin_ranking": L.LOSS_MARGIN_RANKING,
          "softmax_ce": L.LOSS_SOFTMAX_CE}[kind]
    row_loss = torch.empty(n, device="cuda")
    d_pos = torch.empty(n, device="cuda")
    grad = torch.empty(n, n_neg, device="cuda")
    K.loss_fwd_bwd(kd, 2.0, True, 1.3, 1.0, 777, pos.cuda(), neg.cuda(), n, n_neg, n_neg, w.cuda(),
                   row_loss, d_pos, grad)
    torch.cuda.synchronize()
    assert_close(row_loss.sum().cpu(), want.detach(), rtol=1e-5, atol=1e-5)
    assert_close(grad.cpu(), negr.grad, rtol=2e-5, atol=1e-7)
    assert_close(d_pos.cpu(), posr.grad, rtol=2e-5, atol=1e-7)


def _merge_reference(win, ids, best_s, best_i):
    """stable descending merge: current list first, then window columns in order"""
    k = best_s.shape[1]
    cat_s = torch.cat([best_s, win], dim=1)
    cat_i = torch.cat([best_i, ids], dim=1)
    order = torch.sort(cat_s, dim=1, descending=True, stable=True).indices[:, :k]
    return torch.gather(cat_s, 1, order), torch.gather(cat_i, 1, order)


@pytest.mark.parametrize("k", [1, 11, 32, 33, 64])
@pytest.mark.parametrize("n_win,ld,off", [(4096, 4096, 0), (1024, 1032, 0), (1500, 1504, 0),
                                          (777, 777, 0), (2048, 2052, 1), (5, 8, 0)])
@pytest.mark.parametrize("id_mode", ["base", "shared", "per_query"])
def test_topk_merge_stable_with_ties(k, n_win, ld, off, id_mode):
    """bess_topk_merge (register-list kernel for k <= 32, shared-memory list above) ==
    stable sort of cat([current list, window]); quantised scores force ties"""
    L, K, H = _imports()
    g = torch.Generator().manual_seed(k * 7919 + n_win)
    nq = 37
    buf = torch.full((nq * ld + off,), float("nan"))
    win = (torch.randn(nq, n_win, generator=g) * 4).round() / 4  # heavy ties
    view = buf[off:].view(nq, ld)
    view[:, :n_win] = win
    buf = buf.cuda()
    best_s = torch.full((nq, k), -50000.0)
    best_i = torch.full((nq, k), 123456, dtype=torch.int32)
    want_s, want_i = best_s.clone(), best_i.clone()
    best_s, best_i = best_s.cuda(), best_i.cuda()
    for rnd in range(3):  # three windows folded into the same running list
        if id_mode == "base":
            ids_dev, ld_ids, id0 = None, 0, 1000 * rnd
            ids = (torch.arange(n_win, dtype=torch.int32) + id0).expand(nq, n_win)
        elif id_mode == "shared":
            row = torch.randint(1 << 20, (n_win,), generator=g, dtype=torch.int32)
            ids_dev, ld_ids, id0 = row.cuda(), 0, 0
            ids = row.expand(nq, n_win)
        else:
            ids = torch.randint(1 << 20, (nq, n_win), generator=g, dtype=torch.int32)
            ids_dev, ld_ids, id0 = ids.cuda(), n_win, 0
        K.topk_merge(buf[off:], ld, nq, n_win, ids_dev, ld_ids, id0, best_s, best_i, k)
        torch.cuda.synchronize()
        want_s, want_i = _merge_reference(win, ids, want_s, want_i)
        assert torch.equal(best_s.cpu(), want_s), f"round {rnd}"
        assert torch.equal(best_i.cpu(), want_i), f"round {rnd}"
        win = (torch.randn(nq, n_win, generator=g) * 4).round() / 4 + 0.5 * rnd
        view = torch.full((nq * ld + off,), float("nan"))
        view[off:].view(nq, ld)[:, :n_win] = win
        buf = view.cuda()


@pytest.mark.parametrize("rows,W", [(1, 4), (1000, 512), (4099, 260)])
def test_table_operand_cache_rebuilds_only_on_change(rows, W):
    """bess_table_operand_refresh: hi / lo == bess_split_operand; the device-side checksum gates
    the rebuild (an untouched table leaves poisoned outputs alone, a one-word change rebuilds)"""
    L, K, H = _imports()
    g = torch.Generator().manual_seed(rows)
    table = torch.randn(rows, W, generator=g).cuda()
    ld = (W + 7) // 8 * 8
    hi = torch.zeros(rows, ld, device="cuda")
    lo = torch.zeros(rows, ld, device="cuda")
    state = torch.zeros(4, dtype=torch.int64, device="cuda")
    want_hi, want_lo = torch.zeros_like(hi), torch.zeros_like(lo)

    def expect():
        K.split_operand(L.F32, L.rows(table), rows, W, None, L.F32, want_hi, want_lo, ld, None, None,
                        0, table.device)

    K.table_operand_refresh(table, hi, lo, ld, state, True)
    expect()
    torch.cuda.synchronize()
    assert int(state[3]) == 1
    assert torch.equal(hi[:, :W], want_hi[:, :W]) and torch.equal(lo[:, :W], want_lo[:, :W])
    # products are fp32-grade: hi + lo reproduces x to ~2^-22 relative
    assert_close(hi[:, :W] + lo[:, :W], table, rtol=2.0 ** -21, atol=0)
    digest = int(state[1])
    hi.fill_(7.0)
    K.table_operand_refresh(table, hi, lo, ld, state, False)
    torch.cuda.synchronize()
    assert int(state[3]) == 0 and int(state[1]) == digest and int(state[0]) == 0 and int(state[2]) == 0
    assert bool((hi == 7.0).all())  # not rebuilt
    flat = table.view(-1)
    flat[flat.numel() // 2] += 1.0  # one word changes
    K.table_operand_refresh(table, hi, lo, ld, state, False)
    expect()
    torch.cuda.synchronize()
    assert int(state[3]) == 1 and int(state[1]) != digest
    assert torch.equal(hi[:, :W], want_hi[:, :W]) and torch.equal(lo[:, :W], want_lo[:, :W])
    # swapping two different words is a change too (position-dependent checksum)
    if flat.numel() > 2 and float(flat[0]) != float(flat[1]):
        a, b = float(flat[0]), float(flat[1])
        flat[0], flat[1] = b, a
        K.table_operand_refresh(table, hi, lo, ld, state, False)
        torch.cuda.synchronize()
        assert int(state[3]) == 1


def test_select_and_pairs_kernels():
    """bess_select_scores / bess_pairs_get / bess_pairs_set == torch fancy indexing (bit-exact)"""
    L, K, H = _imports()
    g = torch.Generator().manual_seed(3)
    src = torch.randn(50, 1203, generator=g)
    rows = torch.randint(50, (37,), generator=g, dtype=torch.int32)
    cols = torch.randint(-1, 1203, (2000,), generator=g, dtype=torch.int32)
    out = torch.zeros(37, 2000 + 8, device="cuda")[:, :2000]  # row pitch != width
    K.select_scores(src.cuda(), rows.cuda(), 37, cols.cuda(), float("-inf"), out)
    want = src[rows.long()][:, cols.clamp(min=0).long()]
    want[:, cols < 0] = float("-inf")
    assert torch.equal(out.cpu(), want)
    out2 = torch.zeros(50, 2000, device="cuda")
    K.select_scores(src.cuda(), None, 50, cols.cuda(), 7.0, out2)
    want2 = src[:, cols.clamp(min=0).long()]
    want2[:, cols < 0] = 7.0
    assert torch.equal(out2.cpu(), want2)
    mat = src.cuda().clone()
    pc = torch.randint(1203, (50,), generator=g, dtype=torch.int32)
    got = torch.empty(50, device="cuda")
    K.pairs_get(mat, None, pc.cuda(), got)
    assert torch.equal(got.cpu(), src[torch.arange(50), pc.long()])
    pr = torch.randint(50, (300,), generator=g, dtype=torch.int32)
    pc2 = torch.randint(1203, (300,), generator=g, dtype=torch.int32)
    K.pairs_set(mat, pr.cuda(), pc2.cuda(), None, float("-inf"))
    ref = src.clone()
    ref[pr.long(), pc2.long()] = float("-inf")
    assert torch.equal(mat.cpu(), ref)
    vals = torch.randn(50, generator=g)
    K.pairs_set(mat, None, pc.cuda(), vals.cuda())
    ref[torch.arange(50), pc.long()] = vals
    assert torch.equal(mat.cpu(), ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_l2_expansion_kernels(dtype):
    """bess_row_sqnorm / bess_l2_from_dot / bess_l2_coef / bess_rows_axpy vs torch (fp64 reference)"""
    L, K, H = _imports()
    g = torch.Generator().manual_seed(9)
    nq, nc, W, ld, col0 = 70, 45, 40, 64, 8
    q = torch.randn(nq, W, generator=g).to(dtype)
    c = torch.randn(nc, W, generator=g).to(dtype)
    dt = L.dtype_code(dtype)
    qn = torch.empty(nq, device="cuda")
    K.row_sqnorm(dt, L.rows(q.cuda()), nq, W, qn)
    assert_close(qn.cpu().double(), (q.double() ** 2).sum(1), rtol=1e-6, atol=1e-6)
    idx = torch.randperm(nc, generator=g).to(torch.int32)
    cn = torch.empty(nc, device="cuda")
    K.row_sqnorm(dt, L.rows(c.cuda(), idx=idx.cuda()), nc, W, cn)  # rows through an index list
    assert_close(cn.cpu().double(), (c.double()[idx.long()] ** 2).sum(1), rtol=1e-6, atol=1e-6)
    cs = c[idx.long()]
    dots = (q.double() @ cs.double().T).float()
    score = torch.full((nq, ld), 7.0)
    score[:, col0:col0 + nc] = dots
    score = score.cuda()
    K.l2_from_dot(score, L.IDENT, ld, col0, nq, nc, qn, cn)
    want = -torch.cdist(q.double(), cs.double())
    assert_close(score.cpu()[:, col0:col0 + nc].double(), want, rtol=1e-5, atol=2e-4)
    assert bool((score.cpu()[:, :col0] == 7.0).all())
    # coefficient transform + sums
    gs = torch.randn(nq, ld, generator=g)
    sc = score.cpu().clone()
    sc[3, col0 + 2] = 0.0  # zero distance -> coefficient 0
    coef = torch.zeros(nq, 48, device="cuda")
    rb, cb = torch.empty(nq, device="cuda"), torch.empty(nc, device="cuda")
    ws = torch.empty(max(K.l2_coef_workspace(nq, nc) // 4, 1), device="cuda")
    K.l2_coef(gs.cuda(), sc.cuda(), L.IDENT, ld, col0, nq, nc, coef, 48, rb, cb, ws)
    s_blk, g_blk = sc[:, col0:col0 + nc], gs[:, col0:col0 + nc]
    b = torch.where(s_blk != 0, -g_blk / s_blk, torch.zeros_like(g_blk))
    assert_close(coef.cpu()[:, :nc], b, rtol=1e-6, atol=1e-7)
    assert_close(rb.cpu(), b.sum(1), rtol=1e-5, atol=1e-5)
    assert_close(cb.cpu(), b.sum(0), rtol=1e-5, atol=1e-5)
    # out_i += scale * alpha_i * src_i
    out = torch.randn(nc, W, generator=g)
    alpha = torch.randn(nc, generator=g)
    o = out.clone().cuda()
    K.rows_axpy(dt, alpha.cuda(), -1.0, L.rows(c.cuda(), idx=idx.cuda()), L.rows(o), nc, W)
    assert_close(o.cpu(), out - alpha[:, None] * cs.float(), rtol=1e-6, atol=1e-6)
