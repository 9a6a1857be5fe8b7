"""Pins the CPU oracle (oracle/besskge_oracle.py) against fixtures produced by
the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch
from numpy.testing import assert_array_equal
from torch.testing import assert_close

from oracle import besskge_oracle as O

from .conftest import golden_names, load_golden

FAMS = ["TransE", "RotatE", "DistMult", "ComplEx", "PairRE", "BoxE", "TripleRE", "InterHT", "TranS"]


def T(a):
    return torch.from_numpy(np.asarray(a))


def score_cfg(fam, d, v):
    return dict(family=fam, d=d, norm_p=v.get("p", 2) or 2,
                normalize=v.get("normalize_entities", True),
                apply_tanh=v.get("apply_tanh", True), per_dim=v.get("dist_func_per_dim", True),
                eps=1e-6,
                rel_u=v.get("u", v.get("offset", 1.0 if fam in ("InterHT", "TranS") else 0.0)))


def test_sharding_oracle():
    cfg, g = load_golden("host_sharding")
    sh = O.sharding_create(cfg["n_entity"], cfg["n_shard"], cfg["seed"],
                           np.array(cfg["type_offsets"]))
    for k in ("entity_to_shard", "entity_to_idx", "shard_and_idx_to_entity", "shard_counts",
              "entity_type_counts", "entity_type_offsets"):
        assert_array_equal(sh[k], g[k])
    for mode in ("h_shard", "t_shard", "ht_shardpair"):
        st, counts, offsets, order = O.partition_triples(g["triples"], sh, mode)
        tag = f"part_{mode}_0"
        assert_array_equal(st, g[f"{tag}_triples"])
        assert_array_equal(counts, g[f"{tag}_counts"])
        assert_array_equal(offsets, g[f"{tag}_offsets"])
        assert_array_equal(order, g[f"{tag}_sort"])


@pytest.mark.parametrize("fam", FAMS)
def test_scores_oracle(fam):
    cfg, g = load_golden(f"scores_{fam}")
    d = cfg["d"]
    rel, h, t, r = T(g["rel"]), T(g["h"]), T(g["t"]), T(g["r"])
    for vi, v in enumerate(cfg["variants"]):
        c = score_cfg(fam, d, v)
        assert_close(O.score_triple(c, h, rel, r, t), T(g[f"v{vi}_triple"]), rtol=1e-5, atol=1e-5)
        for sharing in (True, False):
            key = f"v{vi}_s{int(sharing)}_heads"
            if key not in g:
                continue
            cand = T(g["c_shared"] if sharing else g["c_per"])
            assert_close(O.score_candidates(c, "h", t, rel, r, cand, sharing), T(g[key]),
                         rtol=1e-5, atol=1e-5)
            assert_close(O.score_candidates(c, "t", h, rel, r, cand, sharing),
                         T(g[f"v{vi}_s{int(sharing)}_tails"]), rtol=1e-5, atol=1e-5)


def loss_cfg(case):
    return dict(kind=case["kind"], margin=case.get("margin", 0.0),
                adversarial=case.get("negative_adversarial_sampling", False),
                adv_scale=case.get("negative_adversarial_scale", 1.0),
                loss_scale=case.get("loss_scale", 1.0), n_entity=case.get("n_entity", 2))


def test_loss_oracle():
    cfg, g = load_golden("loss")
    for i, case in enumerate(cfg["cases"]):
        for wi, w in enumerate((T(g["w"]), torch.tensor([1.0]))):
            pos = T(g["pos"]).clone().requires_grad_(True)
            neg = T(g["neg"]).clone().requires_grad_(True)
            loss = O.loss_value(loss_cfg(case), pos, neg, w)
            loss.backward()
            assert_close(loss.detach(), T(g[f"c{i}_w{wi}_loss"]), rtol=1e-5, atol=1e-6)
            assert_close(pos.grad, T(g[f"c{i}_w{wi}_dpos"]), rtol=1e-5, atol=1e-6)
            assert_close(neg.grad, T(g[f"c{i}_w{wi}_dneg"]), rtol=1e-5, atol=1e-6)


def test_metric_oracle():
    _, g = load_golden("metric")
    for mode in ("optimistic", "pessimistic", "average"):
        for winf in (False, True):
            rk = O.ranks_from_scores(T(g["pos"]), T(g["neg"]), mode, winf)
            assert_close(rk, T(g[f"rank_{mode}_{int(winf)}"]))
    for winf in (False, True):
        assert_close(O.ranks_from_indices(T(g["truth"]), T(g["ids"]), winf),
                     T(g[f"idrank_{int(winf)}"]))
    # hand-written golden vectors of the reference (tests/test_metric.py:13-49)
    pos = torch.tensor([2.1, 5.0, 5.9, 2.0])
    neg = torch.tensor([[2.1, 3.1, 2.1, 5.2, 8.4], [9.8, 5.0, 1.0, 3.2, 5.0],
                        [4.0, 2.3, 5.9, 3.1, 4.5], [4.0, 2.3, 5.9, 3.1, 4.5]])
    assert_close(1 / O.ranks_from_scores(pos, neg, "pessimistic", True),
                 torch.tensor([0.0, 1 / 4, 1 / 2, 0.0]))
    assert_close(1 / O.ranks_from_scores(pos, neg, "optimistic", False),
                 torch.tensor([1 / 4, 1 / 2, 1.0, 1 / 6]))


def bess_inputs(cfg, g):
    n, bps = cfg["n_shard"], cfg["bps"]
    ins = {k[3:]: T(v) for k, v in g.items() if k.startswith("in_")}
    return n, bps, ins


@pytest.mark.parametrize("name", golden_names("bess_"))
def test_bess_forward_oracle(name):
    cfg, g = load_golden(name)
    n, bps, ins = bess_inputs(cfg, g)
    c = score_cfg(cfg["family"], cfg["d"], dict(p=cfg["p"]))
    ent, rel = T(g["ent"]), T(g["rel"])
    S = ins["head"].shape[2] * ins["head"].shape[3]
    pos_all, neg_all = [], []
    for s in range(bps):
        kw = dict(scheme=cfg["scheme"], flat=cfg["flat"], shared=cfg["flat"],
                  negative_mask=ins["negative_mask"][s])
        if cfg["model"] == "EmbeddingMoving":
            pos, neg = O.embedding_moving_forward(c, ent, rel, ins["head"][s], ins["relation"][s],
                                                  ins["tail"][s], ins["negative"][s], **kw)
        else:
            pos, neg = O.score_moving_forward(c, ent, rel, ins["head"][s], ins["relation"][s],
                                              ins["tail"][s], ins["negative"][s],
                                              triple_based=True, **kw)
        pos_all.append(pos.flatten())
        neg_all.append(neg.flatten(end_dim=1))
    assert_close(torch.cat(pos_all), T(g["positive_score"]), rtol=1e-5, atol=1e-4)
    assert_close(torch.cat(neg_all), T(g["negative_score"]), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("name", golden_names("shbess_"))
def test_bess_forward_shared_nonflat_oracle(name):
    """non-flat negatives with negative_sample_sharing (h / t / ht, both BESS variants): the
    reference's outputs from tests/golden/make_golden.py::golden_bess_shared."""
    cfg, g = load_golden(name)
    n, bps, ins = bess_inputs(cfg, g)
    c = score_cfg(cfg["family"], cfg["d"], dict(p=cfg["p"]))
    ent, rel = T(g["ent"]), T(g["rel"])
    kw = dict(scheme=cfg["scheme"], flat=False, shared=True, negative_mask=None)
    if cfg["model"] == "EmbeddingMoving":
        pos, neg = O.embedding_moving_forward(c, ent, rel, ins["head"][0], ins["relation"][0],
                                              ins["tail"][0], ins["negative"][0], **kw)
    else:
        pos, neg = O.score_moving_forward(c, ent, rel, ins["head"][0], ins["relation"][0],
                                          ins["tail"][0], ins["negative"][0], triple_based=False,
                                          **kw)
    assert_close(pos.flatten(), T(g["positive_score"]), rtol=1e-5, atol=1e-4)
    assert_close(neg.flatten(end_dim=1), T(g["negative_score"]), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("name", golden_names("train_"))
def test_training_oracle(name):
    cfg, g = load_golden(name)
    c = score_cfg(cfg["fam"], cfg["d"], dict(p=cfg["p"], **cfg.get("kw", {})))
    batches = []
    for s in range(cfg["n_step"]):
        b = {k[len(f"s{s}_in_"):]: T(v)[0] for k, v in g.items() if k.startswith(f"s{s}_in_")}
        batches.append(b)
    res = O.training_steps(c, loss_cfg(cfg["loss"]), cfg["opt"], T(g["ent0"]), T(g["rel0"]), batches,
                           cfg["scheme"], cfg["flat"], cfg["flat"], "mean")
    for s in range(cfg["n_step"]):
        assert_close(res["loss"][s], T(g[f"s{s}_loss"]), rtol=1e-5, atol=1e-5)
        assert_close(res["grad_ent"][s], T(g[f"s{s}_grad_ent"]), rtol=1e-4, atol=1e-6)
        assert_close(res["grad_rel"][s], T(g[f"s{s}_grad_rel"]), rtol=1e-4, atol=1e-6)
    last = cfg["n_step"] - 1
    assert_close(res["ent"], T(g[f"s{last}_ent"]), rtol=1e-5, atol=1e-6)
    assert_close(res["rel"], T(g[f"s{last}_rel"]), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("name", golden_names("topk_"))
def test_topk_oracle(name):
    cfg, g = load_golden(name)
    n, bps = cfg["n_shard"], cfg["bps"]
    c = score_cfg(cfg["family"], cfg["d"], dict(p=cfg["p"]))
    sh = O.sharding_create(cfg["n_entity"], n, cfg["seed"])
    ent, rel = T(g["ent"]), T(g["rel"])
    fixed_key = "in_head" if cfg["scheme"] == "t" else "in_tail"
    ids, scs = [], []
    for s in range(bps):
        i, sc = O.topk_forward(c, ent, rel, sh, T(g["in_relation"])[s], T(g[fixed_key])[s],
                               cfg["scheme"], cfg["k"])
        ids.append(i.flatten(end_dim=1))
        scs.append(sc.flatten(end_dim=1))
    mask = T(g["triple_mask"]).flatten()
    assert_close(torch.cat(scs)[mask], T(g["topk_scores"])[mask], rtol=1e-5, atol=1e-5)
    assert_array_equal(torch.cat(ids)[mask].numpy(), g["topk_global_id"][mask.numpy()])


@pytest.mark.parametrize("name", golden_names("topk_"))
def test_topk_exact_arithmetic_vs_reference(name):
    """The fixed-order fp32 arithmetic that defines exact ranking (O.exact_scores /
    O.topk_exact, csrc/exact.cu) against the reference's own TopKQueryBessKGE output: scores
    within 1e-5, ids IDENTICAL on every fixture (none of them holds a pair of candidates
    closer than the rounding noise, which is the only place two fp32 orders can disagree)."""
    cfg, g = load_golden(name)
    n, bps = cfg["n_shard"], cfg["bps"]
    c = score_cfg(cfg["family"], cfg["d"], dict(p=cfg["p"]))
    sh = O.sharding_create(cfg["n_entity"], n, cfg["seed"])
    ent, rel = T(g["ent"]), T(g["rel"])
    fixed_key = "in_head" if cfg["scheme"] == "t" else "in_tail"
    ids, scs = [], []
    for s in range(bps):
        i, sc = O.topk_exact(c, ent, rel, sh, T(g["in_relation"])[s], T(g[fixed_key])[s],
                             cfg["scheme"], cfg["k"])
        ids.append(i.flatten(end_dim=1))
        scs.append(sc.flatten(end_dim=1))
    mask = T(g["triple_mask"]).flatten()
    assert_close(torch.cat(scs)[mask], T(g["topk_scores"])[mask], rtol=1e-5, atol=1e-5)
    assert_array_equal(torch.cat(ids)[mask].numpy(), g["topk_global_id"][mask.numpy()])


@pytest.mark.parametrize("name", golden_names("pipeline_"))
def test_all_scores_pipeline_oracle(name):
    """oracle dense restatement of AllScoresPipeline == the reference pipeline's outputs"""
    from .pipeline_cases import filter_list, load_case
    cfg, g, ent, rel = load_case(name)
    c = score_cfg(cfg["family"], cfg["d"], dict(p=cfg["p"]))
    sh = O.sharding_create(cfg["n_entity"], cfg["n_shard"], cfg["seed"])
    mode = "h_shard" if cfg["scheme"] == "t" else "t_shard"
    _, _, _, order = O.partition_triples(g["triples"], sh, mode)
    tr = T(g["triples"][order[g["triple_idx"]]])
    fl = filter_list(cfg, g["triples"])
    res = O.all_scores_pipeline(c, ent, rel, tr, cfg["scheme"],
                                None if fl is None else T(np.concatenate(fl, axis=0)),
                                g["cand_ents"] if cfg["use_candidates"] else None, cfg["k"])
    want = T(g["scores_sub"])
    got = res["scores"][:, ::cfg["score_col_stride"]]
    assert torch.equal(torch.isinf(got), torch.isinf(want))
    fin = torch.isfinite(want)
    assert_close(got[fin], want[fin], rtol=1e-5, atol=1e-5)
    assert_array_equal((~torch.isfinite(res["scores"])).sum(dim=1).numpy(), g["neg_inf_count"])
    assert_close(res["ranks"], T(g["ranks"]))
    assert_array_equal(res["topk_global_id"].numpy(), g["topk_global_id"])
    assert_close((1.0 / res["ranks"]).sum(), T(g["mrr"]), rtol=1e-5, atol=1e-6)
