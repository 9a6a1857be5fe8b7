"""-m gpu: the BESS modules (gather -> exchange -> score -> loss -> backward ->
scatter/optimizer) on CUDA vs fixtures produced by the unmodified reference and
vs the oracle."""
import numpy as np
import pytest
import torch
from torch.testing import assert_close

from oracle import besskge_oracle as O

from .conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu


def _imports():
    import besskge_b200 as B
    from . import gpu_helpers as H
    return B, H


@pytest.mark.parametrize("name", golden_names("bess_"))
def test_bess_forward_vs_reference_golden(name):
    B, H = _imports()
    from besskge_b200.metric import Evaluation
    from besskge_b200.sharding import Sharding
    cfg, g = load_golden(name)
    sh = Sharding.create(cfg["n_entity"], cfg["n_shard"], seed=cfg["seed"])
    sf = H.make_score_fn(cfg["family"], cfg["flat"], cfg["p"], sh, cfg["n_rel"], cfg["d"],
                         H.T(g["ent"]), H.T(g["rel"]))
    ns = H.fake_sampler(cfg["scheme"], cfg["flat"])
    cls = getattr(B.bess, cfg["model"] + "BessKGE")
    ev = Evaluation(["mrr", "hits@3"], mode="average", reduction="sum", return_ranks=True)
    model = cls(negative_sampler=ns, score_fn=sf, evaluation=ev, return_scores=True)
    batch = {k[3:]: H.T(v).flatten(end_dim=1) for k, v in g.items() if k.startswith("in_")}
    res = model(**batch)
    torch.cuda.synchronize()
    assert_close(res["positive_score"].cpu(), H.T(g["positive_score"]), rtol=1e-5, atol=5e-5)
    assert_close(res["negative_score"].cpu(), H.T(g["negative_score"]), rtol=1e-5, atol=5e-5)
    # ranks: a rank can differ from the reference's only where some candidate lies within
    # twice the OBSERVED score error of the positive (in these fixtures: candidates that ARE
    # the true entity, scored through another reduction tree); everywhere else — and that is
    # nearly everywhere — ranks are asserted equal, and MRR over those rows to fp32 precision
    ranks, want = res["ranks"].cpu(), H.T(g["ranks"])
    pos_g, neg_g = H.T(g["positive_score"]).reshape(-1, 1), H.T(g["negative_score"])
    err = max(float((res["positive_score"].cpu().reshape(-1, 1) - pos_g).abs().max()),
              float((res["negative_score"].cpu() - neg_g).abs().max()))
    near = ((neg_g - pos_g).abs() <= 2 * err).any(-1)
    assert bool((ranks[~near] == want[~near]).all())
    lower = 1.0 + (neg_g > pos_g + 2 * err).sum(-1).float()
    upper = 1.0 + (neg_g >= pos_g - 2 * err).sum(-1).float()
    assert bool(((ranks >= lower) & (ranks <= upper)).all())
    # (the fixtures draw candidates from few entities, so 10-20 % of the rows have the true entity
    # among their candidates: an exact tie in exact arithmetic, an ulp apart in any fp32 one)
    assert int(near.sum()) <= 0.25 * near.numel(), f"{int(near.sum())} near-tie rows (err {err:g})"
    assert res["metrics"].shape == tuple(g["metrics"].shape)
    assert_close((1.0 / ranks[~near]).sum(), (1.0 / want[~near]).sum(), rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", golden_names("shbess_"))
def test_bess_forward_shared_nonflat_vs_reference_golden(name):
    """non-flat negatives with negative_sample_sharing (h / t / ht; both BESS variants) against
    the reference's own outputs (bess.py:400-466, 523-566)."""
    B, H = _imports()
    from besskge_b200.sharding import Sharding
    cfg, g = load_golden(name)
    sh = Sharding.create(cfg["n_entity"], cfg["n_shard"], seed=cfg["seed"])
    sf = H.make_score_fn(cfg["family"], True, cfg["p"], sh, cfg["n_rel"], cfg["d"],
                         H.T(g["ent"]), H.T(g["rel"]))
    ns = H.fake_sampler(cfg["scheme"], False, triple_based=False)
    cls = getattr(B.bess, cfg["model"] + "BessKGE")
    model = cls(negative_sampler=ns, score_fn=sf, return_scores=True)
    batch = {k[3:]: H.T(v).flatten(end_dim=1) for k, v in g.items() if k.startswith("in_")}
    res = model(**batch)
    torch.cuda.synchronize()
    want_neg = H.T(g["negative_score"])
    assert res["negative_score"].shape == want_neg.shape
    assert_close(res["positive_score"].cpu(), H.T(g["positive_score"]), rtol=1e-5, atol=5e-5)
    assert_close(res["negative_score"].cpu(), want_neg, rtol=1e-5, atol=5e-5)


@pytest.mark.parametrize("name", golden_names("train_"))
def test_training_vs_reference_golden(name):
    B, H = _imports()
    from besskge_b200.bess import EmbeddingMovingBessKGE, training_model
    from besskge_b200.optim import SGD, AdamW
    from besskge_b200.sharding import Sharding
    cfg, g = load_golden(name)
    sh = Sharding.create(cfg["n_entity"], cfg["n_shard"], seed=cfg["seed"])
    sf = H.make_score_fn(cfg["fam"], cfg["flat"], cfg["p"], sh, cfg["n_rel"], cfg["d"],
                         H.T(g["ent0"]), H.T(g["rel0"]), **cfg.get("kw", {}))
    ns = H.fake_sampler(cfg["scheme"], cfg["flat"], triple_based=False)
    model = EmbeddingMovingBessKGE(ns, sf, loss_fn=H.make_loss(cfg["loss"]))
    o = cfg["opt"]
    opt = (SGD(lr=o["lr"], momentum=o.get("momentum", 0.0)) if o["kind"] == "sgd"
           else AdamW(lr=o["lr"]))
    step = training_model(model, opt, relation_grad_reduction="mean")
    for s in range(cfg["n_step"]):
        batch = {k[len(f"s{s}_in_"):]: H.T(v).flatten(end_dim=1) for k, v in g.items()
                 if k.startswith(f"s{s}_in_")}
        res = step(**batch)
        torch.cuda.synchronize()
        assert_close(res["loss"].cpu(), H.T(g[f"s{s}_loss"]), rtol=1e-5, atol=1e-4,
                     msg=lambda m: f"step {s} loss: {m}")
        assert_close(sf.entity_embedding.detach().cpu(), H.T(g[f"s{s}_ent"]), rtol=1e-5, atol=2e-6,
                     msg=lambda m: f"step {s} entity table: {m}")
        assert_close(sf.relation_embedding.detach().cpu(), H.T(g[f"s{s}_rel"]), rtol=1e-5,
                     atol=2e-6, msg=lambda m: f"step {s} relation table: {m}")


@pytest.mark.parametrize("fam,p,scheme,flat,shared,loss_kind", [
    ("TransE", 1, "t", True, True, "softmax_ce"),
    ("TransE", 2, "h", False, True, "logsigmoid"),
    ("DistMult", 2, "t", False, False, "margin_ranking"),
    ("ComplEx", 2, "ht", True, True, "logsigmoid"),
    ("TransE", 1, "ht", False, True, "logsigmoid"),       # 'ht' with non-flat SHARED negatives
    ("DistMult", 2, "ht", False, True, "margin_ranking"),
    ("RotatE", 1, "ht", False, True, "logsigmoid"),
    ("PairRE", 1, "t", True, True, "logsigmoid"),         # augment + normalised candidates
    ("InterHT", 2, "ht", True, True, "logsigmoid"),
])
@pytest.mark.parametrize("augment", [False, True])
def test_training_vs_oracle_extra(fam, p, scheme, flat, shared, loss_kind, augment):
    """Paths no reference fixture can cover (augment_negative, non-flat shared
    negatives, sampled-softmax): product vs oracle on seeded inputs."""
    B, H = _imports()
    from besskge_b200.bess import EmbeddingMovingBessKGE, training_model
    from besskge_b200.optim import SGD
    from besskge_b200.sharding import Sharding
    if augment and not (flat and shared):
        pytest.skip("augment_negative needs flat shared negatives")
    n, p_part, Nn, d, n_rel, n_ent = 2, 6, 5, 16, 4, 60
    sh = Sharding.create(n_ent, n, seed=3)
    gen = torch.Generator().manual_seed(11)
    ew = 2 if fam in ("RotatE", "ComplEx", "BoxE", "InterHT", "TranS") else 1
    rw = {"ComplEx": 2 * d, "PairRE": 2 * d, "BoxE": 4 * d + 2, "TranS": 3 * d}.get(fam, d)
    ent = torch.randn(n, sh.max_entity_per_shard, ew * d, generator=gen) * 0.5
    rel = torch.randn(n_rel, rw, generator=gen) * 0.5
    S = n * p_part
    Bn = (2 if scheme == "ht" else 1) if flat else S
    lcfg = dict(kind=loss_kind, margin=2.0, negative_adversarial_sampling=True, n_entity=n_ent)
    batches = []
    for _ in range(2):
        batches.append(dict(
            head=torch.randint(sh.shard_counts.min(), (n, n, p_part), generator=gen, dtype=torch.int32),
            tail=torch.randint(sh.shard_counts.min(), (n, n, p_part), generator=gen, dtype=torch.int32),
            relation=torch.randint(n_rel, (n, n, p_part), generator=gen, dtype=torch.int32),
            negative=torch.randint(sh.shard_counts.min(), (n, n, Bn, Nn), generator=gen,
                                   dtype=torch.int32)))
    want = O.training_steps(H.score_cfg(fam, d, p), H.oracle_loss_cfg(lcfg), dict(kind="sgd", lr=0.1),
                            ent, rel, batches, scheme, flat, shared, "mean", augment=augment)
    sf = H.make_score_fn(fam, shared, p, sh, n_rel, d, ent, rel)
    ns = H.fake_sampler(scheme, flat, triple_based=False)
    model = EmbeddingMovingBessKGE(ns, sf, loss_fn=H.make_loss(lcfg), augment_negative=augment)
    step = training_model(model, SGD(lr=0.1))
    for s, b in enumerate(batches):
        res = step(**b)
        assert_close(res["loss"].cpu(), want["loss"][s], rtol=1e-5, atol=1e-4)
    torch.cuda.synchronize()
    assert_close(sf.entity_embedding.detach().cpu(), want["ent"], rtol=1e-5, atol=2e-6)
    assert_close(sf.relation_embedding.detach().cpu(), want["rel"], rtol=1e-5, atol=2e-6)


def _topk_compare(ids, scores, want_ids, want_scores, mask):
    """ids must equal the reference's wherever the reference's neighbouring scores are
    separated by more than the score tolerance; inside a near-tie group the SET of ids
    must agree."""
    assert_close(scores[mask], want_scores[mask], rtol=1e-5, atol=1e-5)
    ids, want_ids, want_scores = ids[mask], want_ids[mask], want_scores[mask]
    tol = 1e-5 + 1e-5 * want_scores.abs()
    gap_ok = (want_scores[:, :-1] - want_scores[:, 1:]) > tol[:, 1:]
    clean = torch.ones_like(want_ids, dtype=torch.bool)
    clean[:, :-1] &= gap_ok
    clean[:, 1:] &= gap_ok
    clean[:, -1] = False  # the k-th entry may trade places with the unseen (k+1)-th
    assert int(clean.sum()) > 0.8 * clean[:, :-1].numel()
    assert bool((ids[clean] == want_ids[clean]).all())


@pytest.mark.parametrize("name", golden_names("topk_"))
def test_topk_query_vs_reference_golden(name):
    B, H = _imports()
    from besskge_b200.bess import TopKQueryBessKGE
    from besskge_b200.metric import Evaluation
    from besskge_b200.negative_sampler import PlaceholderNegativeSampler
    from besskge_b200.sharding import Sharding
    cfg, g = load_golden(name)
    sh = Sharding.create(cfg["n_entity"], cfg["n_shard"], seed=cfg["seed"])
    sf = H.make_score_fn(cfg["family"], True, cfg["p"], sh, cfg["n_rel"], cfg["d"],
                         H.T(g["ent"]), H.T(g["rel"]))
    ns = PlaceholderNegativeSampler(corruption_scheme=cfg["scheme"], seed=cfg["seed"])
    ev = Evaluation(["mrr", "hits@3"], worst_rank_infty=True, reduction="sum", return_ranks=True)
    model = TopKQueryBessKGE(k=cfg["k"], candidate_sampler=ns, score_fn=sf, evaluation=ev,
                             return_scores=True, window_size=cfg["window"])
    model.device_window = 24  # several windows per shard (51 rows)
    batch = {k[3:]: H.T(v).flatten(end_dim=1) for k, v in g.items() if k.startswith("in_")}
    res = model(**batch, triple_mask=H.T(g["triple_mask"]).flatten(end_dim=1))
    torch.cuda.synchronize()
    mask = H.T(g["triple_mask"]).flatten()
    ids, scores = res["topk_global_id"].cpu(), res["topk_scores"].cpu()
    _topk_compare(ids, scores, H.T(g["topk_global_id"]), H.T(g["topk_scores"]), mask)
    # ranks follow from the ids: exact where the ids are exact
    same = (ids == H.T(g["topk_global_id"])).all(-1) & mask
    assert_close(res["ranks"].cpu()[same], H.T(g["ranks"])[same])
    assert res["metrics"].shape == tuple(g["metrics"].shape)
    if bool(same[mask].all()):
        assert_close(res["metrics"].cpu(), H.T(g["metrics"]), rtol=1e-6, atol=1e-6)
    # exact ranking: ids and scores BIT-IDENTICAL to the oracle's fixed-order arithmetic (which
    # tests/test_oracle_golden.py pins to these same reference fixtures) — no tolerance at all
    sh_d = dict(shard_counts=sh.shard_counts, shard_and_idx_to_entity=sh.shard_and_idx_to_entity)
    fixed_key = "in_head" if cfg["scheme"] == "t" else "in_tail"
    ex_ids, ex_sc = [], []
    for s in range(cfg["bps"]):
        i_, s_ = O.topk_exact(H.score_cfg(cfg["family"], cfg["d"], cfg["p"]), H.T(g["ent"]),
                              H.T(g["rel"]), sh_d, H.T(g["in_relation"])[s], H.T(g[fixed_key])[s],
                              cfg["scheme"], cfg["k"])
        ex_ids.append(i_.flatten(end_dim=1))
        ex_sc.append(s_.flatten(end_dim=1))
    assert torch.equal(ids[mask].long(), torch.cat(ex_ids)[mask])
    assert torch.equal(scores[mask], torch.cat(ex_sc)[mask])


@pytest.mark.parametrize("fam,p", [("DistMult", 0), ("TransE", 1), ("RotatE", 2), ("PairRE", 1)])
@pytest.mark.parametrize("flat", [True, False])
def test_topk_query_given_candidates_vs_dense(fam, p, flat):
    """Candidate-list variants (bess.py:712-724): top-k over masked candidate sets equals a
    dense fp64-free restatement with the product's own score functions (the un-sharded
    `score_tails` / `score_heads` evaluated on the same candidates)."""
    B, H = _imports()
    from besskge_b200.bess import TopKQueryBessKGE
    from besskge_b200.sharding import Sharding
    n, S, Nn, d, n_rel, n_ent, k = 2, 6, 37, 16, 4, 90, 4
    sh = Sharding.create(n_ent, n, seed=5)
    gen = torch.Generator().manual_seed(3)
    ew = 2 if fam in ("RotatE", "ComplEx", "BoxE") else 1
    rw = 2 * d if fam in ("ComplEx", "PairRE") else d
    ent = torch.randn(n, sh.max_entity_per_shard, ew * d, generator=gen)
    rel = torch.randn(n_rel, rw, generator=gen)
    sf = H.make_score_fn(fam, flat, p, sh, n_rel, d, ent, rel)
    ns = H.fake_sampler("t", flat, triple_based=True)
    ns.mask_on_gather = True
    model = TopKQueryBessKGE(k=k, candidate_sampler=ns, score_fn=sf, return_scores=True)
    model.device_window = 16
    cmin = int(sh.shard_counts.min())
    Bn = 1 if flat else S
    head = torch.randint(cmin, (n, S), generator=gen, dtype=torch.int32)
    relation = torch.randint(n_rel, (n, S), generator=gen, dtype=torch.int32)
    # row j of `negative`: candidates stored on shard j, for the queries of every shard
    negative = torch.stack([torch.stack([torch.randperm(cmin, generator=gen)[:Nn]
                                         for _ in range(n * Bn)]).view(n, Bn, Nn)
                            for _ in range(n)]).to(torch.int32)
    nmask = torch.rand(n, n, Bn, Nn, generator=gen) > 0.2
    res = model(relation=relation, head=head, negative=negative, negative_mask=nmask)
    torch.cuda.synchronize()
    ids, scores = res["topk_global_id"].cpu(), res["topk_scores"].cpu()
    s2e = torch.from_numpy(sh.shard_and_idx_to_entity.astype("int64"))
    for r in range(n):
        q = ent[r][head[r].long()].cuda()
        all_sc, all_id = [], []
        for j in range(n):
            if flat:
                cidx = negative[j, 0, 0].long()
                sc = sf.score_tails(q, relation[r].cuda(), ent[j][cidx].cuda().unsqueeze(0)).cpu()
                m = nmask[j, 0, 0].expand(S, Nn)
                gid = s2e[j][cidx].expand(S, Nn)
            else:
                cidx = negative[j, r].long()  # [S, Nn]
                sc = sf.score_tails(q, relation[r].cuda(), ent[j][cidx].cuda()).cpu()
                m = nmask[j, r]
                gid = s2e[j][cidx]
            all_sc.append(torch.where(m, sc, sc + K_BAD))
            all_id.append(gid)
        sc, gid = torch.cat(all_sc, 1), torch.cat(all_id, 1)
        top = torch.topk(sc, k, dim=1)
        assert_close(scores[r * S:(r + 1) * S], top.values, rtol=1e-5, atol=1e-5)
        want_ids = torch.gather(gid, 1, top.indices).to(torch.int32)
        _topk_compare(ids[r * S:(r + 1) * S], scores[r * S:(r + 1) * S], want_ids, top.values,
                      torch.ones(S, dtype=torch.bool))


K_BAD = -50000.0


@pytest.mark.parametrize("fam,scheme", [("TransE", "t"), ("RotatE", "ht"), ("TransE", "h")])
def test_l2_tensor_core_path_matches_tile_path_and_oracle(fam, scheme):
    """norm-expanded L2 on the tcgen05 GEMM (csrc/l2.cu) == exact CUDA-core tile kernels == oracle
    for one training step at a multi-tile shape (flat shared negatives, fp32)."""
    B, H = _imports()
    import besskge_b200.bess as bess_mod
    from besskge_b200.bess import EmbeddingMovingBessKGE, training_model
    from besskge_b200.optim import SGD
    from besskge_b200.sharding import Sharding
    n, p_part, Nn, d, n_rel, n_ent = 2, 80, 150, 24, 5, 900
    sh = Sharding.create(n_ent, n, seed=3)
    gen = torch.Generator().manual_seed(7)
    ew = 2 if fam == "RotatE" else 1
    ent = torch.randn(n, sh.max_entity_per_shard, ew * d, generator=gen) * 0.5
    rel = torch.randn(n_rel, d, generator=gen) * 0.5
    Bn = 2 if scheme == "ht" else 1
    lo = int(sh.shard_counts.min())
    batch = dict(
        head=torch.randint(lo, (n, n, p_part), generator=gen, dtype=torch.int32),
        tail=torch.randint(lo, (n, n, p_part), generator=gen, dtype=torch.int32),
        relation=torch.randint(n_rel, (n, n, p_part), generator=gen, dtype=torch.int32),
        negative=torch.randint(lo, (n, n, Bn, Nn), generator=gen, dtype=torch.int32))
    lcfg = dict(kind="logsigmoid", margin=6.0, negative_adversarial_sampling=True)
    want = O.training_steps(H.score_cfg(fam, d, 2), H.oracle_loss_cfg(lcfg), dict(kind="sgd", lr=0.1),
                            ent, rel, [batch], scheme, True, True, "mean")
    got = {}
    min_work = bess_mod.L2_TC_MIN_WORK
    for flag in (True, False):
        bess_mod.USE_L2_TENSOR_CORES = flag
        bess_mod.L2_TC_MIN_WORK = 0  # the test shape is below the production threshold
        try:
            sf = H.make_score_fn(fam, True, 2, sh, n_rel, d, ent, rel)
            model = EmbeddingMovingBessKGE(H.fake_sampler(scheme, True, triple_based=False), sf,
                                           loss_fn=H.make_loss(lcfg), return_scores=True)
            step = training_model(model, SGD(lr=0.1), cuda_graph=False)
            res = step(**batch)
            torch.cuda.synchronize()
        finally:
            bess_mod.USE_L2_TENSOR_CORES = True
            bess_mod.L2_TC_MIN_WORK = min_work
        got[flag] = (res["negative_score"].cpu(), res["loss"].cpu(),
                     sf.entity_embedding.detach().cpu().clone(), sf.relation_embedding.detach().cpu().clone())
        assert_close(got[flag][1], want["loss"][0], rtol=1e-5, atol=1e-3)
        assert_close(got[flag][2], want["ent"], rtol=1e-5, atol=5e-6)
        assert_close(got[flag][3], want["rel"], rtol=1e-5, atol=5e-6)
    assert_close(got[True][0], got[False][0], rtol=1e-5, atol=2e-4)


@pytest.mark.parametrize("fam,p,scheme,flat,shared,opt_kind,loss_kind", [
    ("TransE", 1, "t", True, True, "sgd", "logsigmoid"),
    ("DistMult", 2, "ht", True, True, "sgd", "margin_ranking"),
    ("RotatE", 1, "h", False, False, "sgd", "logsigmoid"),
    ("PairRE", 1, "t", False, False, "adamw", "logsigmoid"),
    ("TransE", 2, "ht", False, False, "sgdm", "logsigmoid"),
    ("TransE", 2, "ht", False, True, "sgd", "logsigmoid"),    # non-flat SHARED negatives
    ("ComplEx", 2, "t", False, True, "sgd", "softmax_ce"),
    ("BoxE", 1, "t", True, True, "sgd", "logsigmoid"),
    ("InterHT", 2, "t", True, True, "sgd", "logsigmoid"),
    ("TranS", 1, "h", False, False, "sgd", "logsigmoid"),
])
def test_score_moving_training_vs_oracle(fam, p, scheme, flat, shared, opt_kind, loss_kind):
    """ScoreMovingBessKGE is a full BessKGE in the reference (bess.py:471-603 + the loss of
    bess.py:254-261): training through it — score gradients back to the scoring shards,
    candidate rows updated where they live, query gradients summed over the scoring shards —
    against the oracle (reference forward restated + autograd + torch.optim)."""
    B, H = _imports()
    from besskge_b200.bess import ScoreMovingBessKGE, training_model
    from besskge_b200.optim import SGD, AdamW
    from besskge_b200.sharding import Sharding
    n, p_part, Nn, d, n_rel, n_ent = 2, 6, 5, 16, 4, 80
    sh = Sharding.create(n_ent, n, seed=3)
    gen = torch.Generator().manual_seed(17)
    ew = 2 if fam in ("RotatE", "ComplEx", "BoxE", "InterHT", "TranS") else 1
    rw = {"ComplEx": 2 * d, "PairRE": 2 * d, "BoxE": 4 * d + 2, "TranS": 3 * d}.get(fam, d)
    ent = torch.randn(n, sh.max_entity_per_shard, ew * d, generator=gen) * 0.5
    rel = torch.randn(n_rel, rw, generator=gen) * 0.5
    S = n * p_part
    Bn = (2 if scheme == "ht" else 1) if flat else S
    lo = int(sh.shard_counts.min())
    lcfg = dict(kind=loss_kind, margin=2.0, negative_adversarial_sampling=True, n_entity=n_ent)
    batches = [dict(
        head=torch.randint(lo, (n, n, p_part), generator=gen, dtype=torch.int32),
        tail=torch.randint(lo, (n, n, p_part), generator=gen, dtype=torch.int32),
        relation=torch.randint(n_rel, (n, n, p_part), generator=gen, dtype=torch.int32),
        negative=torch.randint(lo, (n, n, Bn, Nn), generator=gen, dtype=torch.int32))
        for _ in range(3)]
    ocfg = {"sgd": dict(kind="sgd", lr=0.1), "sgdm": dict(kind="sgd", lr=0.1, momentum=0.9),
            "adamw": dict(kind="adamw", lr=0.01, eps=1e-4)}[opt_kind]
    want = O.training_steps(H.score_cfg(fam, d, p), H.oracle_loss_cfg(lcfg), ocfg, ent, rel, batches,
                            scheme, flat, shared, "mean", model="score_moving", triple_based=False)
    sf = H.make_score_fn(fam, shared, p, sh, n_rel, d, ent, rel)
    ns = H.fake_sampler(scheme, flat, triple_based=False)
    model = ScoreMovingBessKGE(ns, sf, loss_fn=H.make_loss(lcfg))
    opt = {"sgd": SGD(lr=0.1), "sgdm": SGD(lr=0.1, momentum=0.9),
           "adamw": AdamW(lr=0.01, eps=1e-4)}[opt_kind]
    step = training_model(model, opt)
    adam = opt_kind == "adamw"
    for s, b in enumerate(batches):
        res = step(**b)
        assert_close(res["loss"].cpu(), want["loss"][s], rtol=2e-5 if adam else 1e-5, atol=1e-4,
                     msg=lambda m: f"step {s}: {m}")
    torch.cuda.synchronize()
    tol = dict(rtol=1e-4, atol=2e-5) if adam else dict(rtol=1e-5, atol=2e-6)
    assert_close(sf.entity_embedding.detach().cpu(), want["ent"], **tol)
    if fam == "BoxE":  # width gradients pass through the log / exp of the geometric-mean normalisation
        tol = dict(rtol=1e-4, atol=3e-5)
    assert_close(sf.relation_embedding.detach().cpu(), want["rel"], **tol)
