"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (needs /root/reference):
    python tests/golden/make_golden.py
The reference is imported through `oracle/ref_loader.py` (stubs only for the
absent third-party modules); every array written here is an output of the
reference's own code on seeded inputs.  Fixtures are small (.npz, < 1 MB each)
and committed; the GPU box never needs /root/reference.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref_loader  # noqa: E402

OUT = Path(__file__).resolve().parent
SEED = 1234


def save(name: str, cfg: dict, **arrays) -> None:
    clean = {}
    for k, v in arrays.items():
        if v is None:
            continue
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        clean[k] = np.asarray(v)
    np.savez_compressed(OUT / f"{name}.npz", __cfg__=np.array(json.dumps(cfg)), **clean)
    print(f"wrote {name}.npz  ({sum(a.nbytes for a in clean.values()) / 1e3:.1f} kB raw)")


def make_dataset(ref, n_entity, n_rel, n_triple, n_neg, rng, type_offsets=None, flat_neg=False):
    h = rng.integers(n_entity, size=n_triple)
    t = rng.integers(n_entity, size=n_triple)
    r = rng.integers(n_rel, size=n_triple)
    triples = {"test": np.stack([h, r, t], axis=1)}
    outer = 1 if flat_neg else n_triple
    nh = rng.integers(n_entity, size=(outer, n_neg), dtype=np.int32)
    nt = rng.integers(n_entity, size=(outer, n_neg), dtype=np.int32)
    ds = ref.dataset.KGDataset(
        n_entity=n_entity, n_relation_type=n_rel, entity_dict=None, relation_dict=None,
        type_offsets=None if type_offsets is None else {str(i): int(o) for i, o in enumerate(type_offsets)},
        triples=triples, original_triple_ids={"test": np.arange(n_triple)},
        neg_heads={"test": nh}, neg_tails={"test": nt})
    return ds


# --------------------------------------------------------------------- host --
def golden_host(ref) -> None:
    rng = np.random.default_rng(SEED)
    type_offsets = np.array([0, 20, 60, 200, 341, 342])
    n_entity, n_shard, n_rel, n_triple, n_neg = 501, 5, 11, 2001, 17
    sh = ref.sharding.Sharding.create(n_entity, n_shard, seed=SEED, type_offsets=type_offsets)
    arrays = dict(
        entity_to_shard=sh.entity_to_shard, entity_to_idx=sh.entity_to_idx,
        shard_and_idx_to_entity=sh.shard_and_idx_to_entity, shard_counts=sh.shard_counts,
        entity_type_counts=sh.entity_type_counts, entity_type_offsets=sh.entity_type_offsets)
    sh2 = ref.sharding.Sharding.create(500, 4, seed=SEED)
    arrays.update(s2_entity_to_shard=sh2.entity_to_shard, s2_entity_to_idx=sh2.entity_to_idx,
                  s2_table=sh2.shard_and_idx_to_entity, s2_counts=sh2.shard_counts)
    ds = make_dataset(ref, n_entity, n_rel, n_triple, n_neg, rng, type_offsets)
    arrays.update(triples=ds.triples["test"], neg_heads=ds.neg_heads["test"],
                  neg_tails=ds.neg_tails["test"])
    for mode in ("h_shard", "t_shard", "ht_shardpair"):
        for inv in (False, True):
            pts = ref.sharding.PartitionedTripleSet.create_from_dataset(
                ds, "test", sh, mode, add_inverse_triples=inv)
            tag = f"part_{mode}_{int(inv)}"
            arrays.update({f"{tag}_triples": pts.triples, f"{tag}_counts": pts.triple_counts,
                           f"{tag}_offsets": pts.triple_offsets, f"{tag}_sort": pts.triple_sort_idx,
                           f"{tag}_types": pts.types, f"{tag}_nh": pts.neg_heads,
                           f"{tag}_nt": pts.neg_tails})
    # queries
    queries = ds.triples["test"][:301, :2]
    ptq = ref.sharding.PartitionedTripleSet.create_from_queries(
        ds, sh, queries, "hr", ground_truth=ds.triples["test"][:301, 2])
    arrays.update(q_triples=ptq.triples, q_counts=ptq.triple_counts, q_offsets=ptq.triple_offsets,
                  q_sort=ptq.triple_sort_idx, q_nt=ptq.neg_tails)
    save("host_sharding", dict(n_entity=n_entity, n_shard=n_shard, n_rel=n_rel, seed=SEED,
                               type_offsets=type_offsets.tolist(), n_query=301), **arrays)

    # ---- samplers on the ht_shardpair partition
    pts = ref.sharding.PartitionedTripleSet.create_from_dataset(ds, "test", sh, "ht_shardpair")
    arrays = {}
    cases = []
    for scheme in ("h", "t", "ht"):
        for flat in (True, False):
            for local in (False, True):
                ns = ref.negative_sampler.RandomShardedNegativeSampler(
                    n_negative=7, sharding=sh, seed=SEED, corruption_scheme=scheme,
                    local_sampling=local, flat_negative_format=flat)
                bs = ref.batch_sampler.RandomShardedBatchSampler(
                    pts, ns, shard_bs=20, batches_per_step=3, seed=SEED,
                    hrt_freq_weighting=(scheme == "t"), weight_smoothing=0.5)
                tag = f"rand_{scheme}_{int(flat)}_{int(local)}"
                for it in range(2):
                    batch = bs[[it]]
                    for k, v in batch.items():
                        arrays[f"{tag}_{it}_{k}"] = v.numpy()
                cases.append(dict(tag=tag, scheme=scheme, flat=flat, local=local, sampler="random"))
    # type-based
    for scheme in ("h", "ht"):
        ns = ref.negative_sampler.TypeBasedShardedNegativeSampler(
            triple_types=pts.types, n_negative=5, sharding=sh, corruption_scheme=scheme,
            local_sampling=False, seed=SEED)
        bs = ref.batch_sampler.RandomShardedBatchSampler(pts, ns, shard_bs=20, batches_per_step=2,
                                                         seed=SEED)
        tag = f"type_{scheme}"
        batch = bs[[0]]
        for k, v in batch.items():
            arrays[f"{tag}_0_{k}"] = v.numpy()
        cases.append(dict(tag=tag, scheme=scheme, sampler="type"))
    # triple-based + rigid
    for scheme in ("h", "t", "ht"):
        for flat in (True, False):
            ds2 = make_dataset(ref, n_entity, n_rel, n_triple, n_neg, np.random.default_rng(SEED + 1),
                               type_offsets, flat_neg=flat)
            pts2 = ref.sharding.PartitionedTripleSet.create_from_dataset(ds2, "test", sh,
                                                                         "ht_shardpair")
            for mog in (False, True):
                ns = ref.negative_sampler.TripleBasedShardedNegativeSampler(
                    pts2.neg_heads, pts2.neg_tails, sh, corruption_scheme=scheme, seed=SEED,
                    mask_on_gather=mog, return_sort_idx=True)
                bs = ref.batch_sampler.RigidShardedBatchSampler(
                    pts2, ns, shard_bs=20, batches_per_step=2, seed=SEED,
                    duplicate_batch=(scheme == "ht"), return_triple_idx=True)
                tag = f"trip_{scheme}_{int(flat)}_{int(mog)}"
                sampler = list(bs.get_dataloader_sampler(shuffle=False))
                arrays[f"{tag}_len"] = np.array([len(bs), len(sampler)])
                for it in (0, len(sampler) - 1):
                    batch = bs[sampler[it]]
                    for k, v in batch.items():
                        arrays[f"{tag}_{it}_{k}"] = v.numpy()
                cases.append(dict(tag=tag, scheme=scheme, flat=flat, mog=mog, sampler="triple",
                                  last=len(sampler) - 1))
    save("host_samplers", dict(n_entity=n_entity, n_shard=n_shard, n_rel=n_rel, n_triple=n_triple,
                               n_neg=n_neg, seed=SEED, type_offsets=type_offsets.tolist(),
                               cases=cases), **arrays)


# ------------------------------------------------------------------ scoring --
FAMILIES = {
    "TransE": dict(cls="TransE", norm=True, ew=1, rw=lambda d: d),
    "RotatE": dict(cls="RotatE", norm=True, ew=2, rw=lambda d: d),
    "DistMult": dict(cls="DistMult", norm=False, ew=1, rw=lambda d: d),
    "ComplEx": dict(cls="ComplEx", norm=False, ew=2, rw=lambda d: 2 * d),
    "PairRE": dict(cls="PairRE", norm=True, ew=1, rw=lambda d: 2 * d),
    "BoxE": dict(cls="BoxE", norm=True, ew=2, rw=lambda d: 4 * d + 2),
    "TripleRE": dict(cls="TripleRE", norm=True, ew=1, rw=lambda d: 3 * d),
    "InterHT": dict(cls="InterHT", norm=True, ew=2, rw=lambda d: d),
    "TranS": dict(cls="TranS", norm=True, ew=2, rw=lambda d: 3 * d),
}


def build_score_fn(ref, fam, sharing, p, sh, n_rel, d, ent, rel, **kw):
    cls = getattr(ref.scoring, FAMILIES[fam]["cls"])
    if FAMILIES[fam]["norm"]:
        return cls(negative_sample_sharing=sharing, scoring_norm=p, sharding=sh,
                   n_relation_type=n_rel, embedding_size=d, entity_initializer=ent,
                   relation_initializer=rel, **kw)
    return cls(negative_sample_sharing=sharing, sharding=sh, n_relation_type=n_rel,
               embedding_size=d, entity_initializer=ent, relation_initializer=rel, **kw)


def tables(fam, sh, n_rel, d, gen, scale=1.0):
    W = FAMILIES[fam]["ew"] * d
    Wr = FAMILIES[fam]["rw"](d)
    ent = torch.randn(sh.n_shard, sh.max_entity_per_shard, W, generator=gen) * scale
    rel = torch.randn(n_rel, Wr, generator=gen) * scale
    return ent, rel


def golden_scores(ref, only=None) -> None:
    sh = ref.sharding.Sharding.create(60, 1, seed=SEED)
    n_rel, d, b, nn_ = 7, 16, 12, 9
    for fam in FAMILIES:
        if only is not None and fam not in only:
            continue
        gen = torch.Generator().manual_seed(SEED)
        arrays = {}
        variants = [dict(p=1), dict(p=2)] if FAMILIES[fam]["norm"] else [dict(p=0)]
        if fam == "PairRE":
            variants.append(dict(p=1, normalize_entities=False))
        if fam == "TripleRE":
            variants += [dict(p=1, u=0.5), dict(p=2, normalize_entities=False, u=1.25)]
        if fam == "BoxE":
            variants += [dict(p=2, apply_tanh=False), dict(p=1, dist_func_per_dim=False)]
        if fam == "InterHT":
            variants.append(dict(p=1, normalize_entities=False, offset=0.5))
        if fam == "TranS":
            variants.append(dict(p=2, normalize_entities=False, offset=0.25))
        ent, rel = tables(fam, sh, n_rel, d, gen)
        W = ent.shape[-1]
        h = torch.randn(b, W, generator=gen)
        t = torch.randn(b, W, generator=gen)
        r = torch.randint(n_rel, (b,), generator=gen)
        c_shared = torch.randn(1, nn_, W, generator=gen)
        c_per = torch.randn(b, nn_, W, generator=gen)
        arrays.update(ent=ent, rel=rel, h=h, t=t, r=r, c_shared=c_shared, c_per=c_per)
        vcfg = []
        for vi, v in enumerate(variants):
            kw = {k: val for k, val in v.items() if k != "p"}
            for sharing in (True, False):
                sf = build_score_fn(ref, fam, sharing, v["p"], sh, n_rel, d, ent, rel, **kw)
                with torch.no_grad():
                    arrays[f"v{vi}_triple"] = sf.score_triple(h, r, t)
                    cand = c_shared if sharing else c_per
                    if fam == "BoxE" and not v.get("dist_func_per_dim", True):
                        continue  # broadcast form differs only by the all() reduction; triple only
                    arrays[f"v{vi}_s{int(sharing)}_heads"] = sf.score_heads(cand, r, t)
                    arrays[f"v{vi}_s{int(sharing)}_tails"] = sf.score_tails(h, r, cand)
            vcfg.append(v)
        save(f"scores_{fam}", dict(family=fam, d=d, n_rel=n_rel, variants=vcfg), **arrays)


def golden_loss(ref) -> None:
    gen = torch.Generator().manual_seed(SEED)
    S, N = 10, 13
    pos = torch.randn(S, generator=gen) * 3
    neg = torch.randn(S, N, generator=gen) * 3
    w = torch.rand(S, generator=gen)
    arrays = dict(pos=pos, neg=neg, w=w)
    cases = []
    specs = [
        ("logsigmoid", dict(margin=2.0, negative_adversarial_sampling=True,
                            negative_adversarial_scale=0.7, loss_scale=1.5)),
        ("logsigmoid", dict(margin=12.0, negative_adversarial_sampling=False)),
        ("margin_ranking", dict(margin=1.0, negative_adversarial_sampling=True)),
        ("margin_ranking", dict(margin=0.5, negative_adversarial_sampling=False, loss_scale=2.0)),
        ("softmax_ce", dict(n_entity=1000, loss_scale=1.25)),
    ]
    for i, (kind, kw) in enumerate(specs):
        cls = {"logsigmoid": ref.loss.LogSigmoidLoss, "margin_ranking": ref.loss.MarginRankingLoss,
               "softmax_ce": ref.loss.SampledSoftmaxCrossEntropyLoss}[kind]
        fn = cls(**kw)
        for wi, wt in enumerate((w, torch.tensor([1.0]))):
            p_ = pos.clone().requires_grad_(True)
            n_ = neg.clone().requires_grad_(True)
            loss = fn(p_, n_ * 1.0, wt)  # `* 1.0`: the CE loss mutates its argument in place
            loss.backward()
            arrays[f"c{i}_w{wi}_loss"] = loss.detach()
            arrays[f"c{i}_w{wi}_dpos"] = p_.grad
            arrays[f"c{i}_w{wi}_dneg"] = n_.grad
        cases.append(dict(kind=kind, **kw))
    save("loss", dict(cases=cases), **arrays)


def golden_metric(ref) -> None:
    gen = torch.Generator().manual_seed(SEED)
    pos = torch.randn(9, generator=gen)
    neg = torch.randn(9, 14, generator=gen)
    neg[0, 3] = pos[0]
    neg[1] = pos[1] + 1.0
    arrays = dict(pos=pos, neg=neg)
    for mode in ("optimistic", "pessimistic", "average"):
        for winf in (False, True):
            ev = ref.metric.Evaluation(["mrr", "hits@1", "hits@5"], mode=mode, worst_rank_infty=winf)
            rk = ev.ranks_from_scores(pos.clone(), neg)
            arrays[f"rank_{mode}_{int(winf)}"] = rk
            m = ev.dict_metrics_from_ranks(rk)
            for k, v in m.items():
                arrays[f"m_{mode}_{int(winf)}_{k}"] = v
    truth = torch.tensor([6, 0, 2])
    ids = torch.tensor([[6, 1, 45, 33, 28], [5, 2, 12, 0, 44], [27, 9, 1, 6, 17]])
    arrays.update(truth=truth, ids=ids)
    for winf in (False, True):
        ev = ref.metric.Evaluation(["mrr"], worst_rank_infty=winf)
        arrays[f"idrank_{int(winf)}"] = ev.ranks_from_indices(truth, ids)
    save("metric", {}, **arrays)


# --------------------------------------------------------------------- bess --
def golden_bess(ref, only=None) -> None:
    n_entity, n_rel, n_shard, n_triple, bps, shard_bs, n_neg, d = 200, 6, 4, 400, 2, 16, 24, 16
    sh = ref.sharding.Sharding.create(n_entity, n_shard, seed=SEED)
    combos = [
        ("TransE", 1), ("TransE", 2), ("RotatE", 1), ("DistMult", 0), ("ComplEx", 0),
        ("PairRE", 1), ("BoxE", 1), ("BoxE", 2), ("TripleRE", 1), ("InterHT", 1), ("TranS", 2),
    ]
    if only is not None:
        combos = [c for c in combos if c[0] in only]
    for model_name in ("EmbeddingMoving", "ScoreMoving"):
        for fam, p in combos:
            if model_name == "ScoreMoving" and fam in ("RotatE", "ComplEx", "BoxE") and p != 2:
                pass
            for scheme, dup in (("h", False), ("t", False), ("ht", True)):
                for flat in (True, False):
                    if fam not in ("TransE",) and (scheme == "h" and not flat):
                        continue  # keep the fixture set small: full grid only for TransE
                    rng = np.random.default_rng(SEED + 7)
                    ds = make_dataset(ref, n_entity, n_rel, n_triple, n_neg, rng, flat_neg=flat)
                    pts = ref.sharding.PartitionedTripleSet.create_from_dataset(
                        ds, "test", sh, partition_mode="ht_shardpair")
                    gen = torch.Generator().manual_seed(SEED)
                    ent, rel = tables(fam, sh, n_rel, d, gen)
                    sf = build_score_fn(ref, fam, flat, p, sh, n_rel, d, ent, rel)
                    ns = ref.negative_sampler.TripleBasedShardedNegativeSampler(
                        pts.neg_heads, pts.neg_tails, sh, corruption_scheme=scheme, seed=SEED,
                        return_sort_idx=False, mask_on_gather=False)
                    bs = ref.batch_sampler.RigidShardedBatchSampler(
                        partitioned_triple_set=pts, negative_sampler=ns, shard_bs=shard_bs,
                        batches_per_step=bps, seed=SEED, duplicate_batch=dup)
                    cls = getattr(ref.bess, model_name + "BessKGE")
                    ev = ref.metric.Evaluation(["mrr", "hits@3"], mode="average", reduction="sum",
                                               return_ranks=True)
                    model = cls(negative_sampler=ns, score_fn=sf, evaluation=ev, return_scores=True)
                    batch = bs[list(bs.get_dataloader_sampler(shuffle=False))[0]]
                    with torch.no_grad():
                        res = ref_loader.run_replicated(model, batch, n_shard, bps)
                    name = f"bess_{model_name}_{fam}{p}_{scheme}_{int(flat)}"
                    save(name, dict(model=model_name, family=fam, p=p, scheme=scheme, flat=flat,
                                    dup=dup, d=d, n_rel=n_rel, n_entity=n_entity, n_shard=n_shard,
                                    bps=bps, shard_bs=shard_bs, seed=SEED),
                         ent=ent, rel=rel,
                         **{f"in_{k}": v for k, v in batch.items()},
                         positive_score=res["positive_score"], negative_score=res["negative_score"],
                         ranks=res["ranks"], metrics=res["metrics"])


def golden_bess_shared(ref) -> None:
    """Non-flat negatives WITH negative_sample_sharing (every query of a group is scored against
    the negatives of all its queries: bess.py:400-466 / 523-566 with scoring.py's shared
    broadcasting) — the combination the `bess_` grid above leaves out, because the triple-based
    sampler there carries a padding mask that the reference cannot combine with sharing."""
    n_entity, n_rel, n_shard, n_triple, shard_bs, n_neg, d = 200, 6, 4, 400, 16, 3, 16
    sh = ref.sharding.Sharding.create(n_entity, n_shard, seed=SEED)
    for model_name in ("EmbeddingMoving", "ScoreMoving"):
        for fam, p in (("TransE", 1), ("DistMult", 0), ("RotatE", 2)):
            for scheme in ("h", "t", "ht"):
                rng = np.random.default_rng(SEED + 13)
                ds = make_dataset(ref, n_entity, n_rel, n_triple, 4, rng)
                pts = ref.sharding.PartitionedTripleSet.create_from_dataset(
                    ds, "test", sh, "ht_shardpair")
                gen = torch.Generator().manual_seed(SEED)
                ent, rel = tables(fam, sh, n_rel, d, gen)
                sf = build_score_fn(ref, fam, True, p, sh, n_rel, d, ent, rel)
                ns = ref.negative_sampler.RandomShardedNegativeSampler(
                    n_negative=n_neg, sharding=sh, seed=SEED, corruption_scheme=scheme,
                    local_sampling=False, flat_negative_format=False)
                bs = ref.batch_sampler.RandomShardedBatchSampler(pts, ns, shard_bs=shard_bs,
                                                                 batches_per_step=1, seed=SEED)
                cls = getattr(ref.bess, model_name + "BessKGE")
                model = cls(negative_sampler=ns, score_fn=sf, return_scores=True)
                batch = bs[[0]]
                with torch.no_grad():
                    res = ref_loader.run_replicated(model, batch, n_shard, 1)
                save(f"shbess_{model_name}_{fam}{p}_{scheme}",
                     dict(model=model_name, family=fam, p=p, scheme=scheme, flat=False, shared=True,
                          d=d, n_rel=n_rel, n_entity=n_entity, n_shard=n_shard, bps=1,
                          shard_bs=shard_bs, seed=SEED),
                     ent=ent, rel=rel, **{f"in_{k}": v for k, v in batch.items()},
                     positive_score=res["positive_score"], negative_score=res["negative_score"])


def golden_train(ref, only=None) -> None:
    """reference forward -> torch.autograd -> dense torch.optim, n_shard 4 and 1."""
    n_entity, n_rel, n_triple, d, n_step = 200, 6, 600, 16, 3
    specs = [
        dict(fam="TransE", p=2, n_shard=1, scheme="t", flat=True, n_neg=8, shard_bs=32,
             loss=("logsigmoid", dict(margin=12.0, negative_adversarial_sampling=True)),
             opt=dict(kind="sgd", lr=0.05)),
        dict(fam="TransE", p=1, n_shard=4, scheme="ht", flat=True, n_neg=5, shard_bs=16,
             loss=("logsigmoid", dict(margin=3.0, negative_adversarial_sampling=True,
                                      negative_adversarial_scale=0.5)),
             opt=dict(kind="sgd", lr=0.05)),
        dict(fam="DistMult", p=0, n_shard=4, scheme="t", flat=True, n_neg=6, shard_bs=16,
             loss=("margin_ranking", dict(margin=1.0, negative_adversarial_sampling=False)),
             opt=dict(kind="sgd", lr=0.05, momentum=0.9)),
        dict(fam="ComplEx", p=0, n_shard=4, scheme="h", flat=True, n_neg=6, shard_bs=16,
             loss=("logsigmoid", dict(margin=1.0, negative_adversarial_sampling=False)),
             opt=dict(kind="adamw", lr=0.01)),
        dict(fam="RotatE", p=1, n_shard=4, scheme="t", flat=False, n_neg=3, shard_bs=16,
             loss=("logsigmoid", dict(margin=2.0, negative_adversarial_sampling=True)),
             opt=dict(kind="sgd", lr=0.05)),
        dict(fam="PairRE", p=2, n_shard=2, scheme="ht", flat=False, n_neg=3, shard_bs=16,
             loss=("margin_ranking", dict(margin=2.0, negative_adversarial_sampling=True)),
             opt=dict(kind="sgd", lr=0.05)),
        dict(fam="BoxE", p=2, n_shard=2, scheme="t", flat=True, n_neg=6, shard_bs=16,
             loss=("logsigmoid", dict(margin=3.0, negative_adversarial_sampling=False)),
             opt=dict(kind="sgd", lr=0.05)),
        # NOTE: augment_negative cannot be generated from the unmodified reference on CPU
        # torch: bess.py:388 calls .view() on a non-contiguous split (works only when
        # traced by PopTorch).  That path is checked against the oracle only.
        dict(fam="TripleRE", p=1, n_shard=4, scheme="t", flat=True, n_neg=6, shard_bs=16,
             loss=("logsigmoid", dict(margin=3.0, negative_adversarial_sampling=True)),
             opt=dict(kind="sgd", lr=0.05), kw=dict(u=0.5)),
        dict(fam="InterHT", p=1, n_shard=4, scheme="t", flat=True, n_neg=6, shard_bs=16,
             loss=("logsigmoid", dict(margin=3.0, negative_adversarial_sampling=True)),
             opt=dict(kind="sgd", lr=0.05)),
        dict(fam="TranS", p=2, n_shard=2, scheme="ht", flat=False, n_neg=3, shard_bs=16,
             loss=("margin_ranking", dict(margin=2.0, negative_adversarial_sampling=True)),
             opt=dict(kind="sgd", lr=0.05), kw=dict(offset=0.5)),
    ]
    for si, sp in enumerate(specs):
        if only is not None and sp["fam"] not in only:
            continue
        n_shard = sp["n_shard"]
        sh = ref.sharding.Sharding.create(n_entity, n_shard, seed=SEED)
        rng = np.random.default_rng(SEED + 11)
        ds = make_dataset(ref, n_entity, n_rel, n_triple, 4, rng)
        pts = ref.sharding.PartitionedTripleSet.create_from_dataset(ds, "test", sh, "ht_shardpair")
        gen = torch.Generator().manual_seed(SEED + si)
        ent, rel = tables(sp["fam"], sh, n_rel, d, gen, scale=0.5)
        sf = build_score_fn(ref, sp["fam"], sp["flat"], sp["p"], sh, n_rel, d, ent.clone(),
                            rel.clone(), **sp.get("kw", {}))
        ns = ref.negative_sampler.RandomShardedNegativeSampler(
            n_negative=sp["n_neg"], sharding=sh, seed=SEED, corruption_scheme=sp["scheme"],
            local_sampling=False, flat_negative_format=sp["flat"])
        bs = ref.batch_sampler.RandomShardedBatchSampler(pts, ns, shard_bs=sp["shard_bs"],
                                                         batches_per_step=1, seed=SEED)
        kind, kw = sp["loss"]
        loss_fn = {"logsigmoid": ref.loss.LogSigmoidLoss, "margin_ranking": ref.loss.MarginRankingLoss,
                   "softmax_ce": ref.loss.SampledSoftmaxCrossEntropyLoss}[kind](**kw)
        model = ref.bess.EmbeddingMovingBessKGE(negative_sampler=ns, score_fn=sf, loss_fn=loss_fn,
                                                augment_negative=sp.get("augment", False))
        o = sp["opt"]
        params = [model.entity_embedding, sf.relation_embedding]
        if o["kind"] == "sgd":
            opt = torch.optim.SGD(params, lr=o["lr"], momentum=o.get("momentum", 0.0))
        else:
            opt = torch.optim.AdamW(params, lr=o["lr"])
        arrays = dict(ent0=ent, rel0=rel)
        for step in range(n_step):
            batch = bs[[step]]
            opt.zero_grad(set_to_none=True)
            res = ref_loader.run_replicated(model, batch, n_shard, 1, grad=True)
            res["loss"].sum().backward()
            sf.relation_embedding.grad.div_(n_shard)  # PopTorch default: mean over replicas
            for k, v in batch.items():
                arrays[f"s{step}_in_{k}"] = v
            arrays[f"s{step}_loss"] = res["loss"].detach()
            arrays[f"s{step}_grad_ent"] = model.entity_embedding.grad.detach().clone()
            arrays[f"s{step}_grad_rel"] = sf.relation_embedding.grad.detach().clone()
            opt.step()
            arrays[f"s{step}_ent"] = model.entity_embedding.detach().clone()
            arrays[f"s{step}_rel"] = sf.relation_embedding.detach().clone()
        cfg = dict(sp)
        cfg["loss"] = dict(kind=kind, **kw)
        cfg.update(d=d, n_rel=n_rel, n_entity=n_entity, n_step=n_step, seed=SEED)
        save(f"train_{si}_{sp['fam']}", cfg, **arrays)


def golden_topk(ref) -> None:
    n_entity, n_rel, n_shard, n_triple, d, k = 203, 6, 4, 300, 16, 5
    sh = ref.sharding.Sharding.create(n_entity, n_shard, seed=SEED)
    for fam, p in (("DistMult", 0), ("ComplEx", 0), ("TransE", 1)):
        for scheme in ("t", "h"):
            rng = np.random.default_rng(SEED + 3)
            ds = make_dataset(ref, n_entity, n_rel, n_triple, 4, rng)
            mode = "h_shard" if scheme == "t" else "t_shard"
            pts = ref.sharding.PartitionedTripleSet.create_from_dataset(ds, "test", sh, mode)
            gen = torch.Generator().manual_seed(SEED)
            ent, rel = tables(fam, sh, n_rel, d, gen)
            sf = build_score_fn(ref, fam, True, p, sh, n_rel, d, ent, rel)
            ns = ref.negative_sampler.PlaceholderNegativeSampler(corruption_scheme=scheme, seed=SEED)
            bs = ref.batch_sampler.RigidShardedBatchSampler(pts, ns, shard_bs=12, batches_per_step=2,
                                                            seed=SEED, return_triple_idx=True)
            ev = ref.metric.Evaluation(["mrr", "hits@3"], worst_rank_infty=True, reduction="sum",
                                       return_ranks=True)
            model = ref.bess.TopKQueryBessKGE(k=k, candidate_sampler=ns, score_fn=sf, evaluation=ev,
                                              return_scores=True, window_size=17)
            batch = bs[list(bs.get_dataloader_sampler(shuffle=False))[0]]
            tmask = batch.pop("triple_mask")
            tidx = batch.pop("triple_idx")
            with torch.no_grad():
                res = ref_loader.run_replicated(model, dict(batch, triple_mask=tmask), n_shard, 2)
            save(f"topk_{fam}_{scheme}", dict(family=fam, p=p, scheme=scheme, d=d, n_rel=n_rel,
                                              n_entity=n_entity, n_shard=n_shard, k=k, window=17,
                                              bps=2, shard_bs=12, seed=SEED),
                 ent=ent, rel=rel, triple_mask=tmask, triple_idx=tidx,
                 **{f"in_{kk}": v for kk, v in batch.items()},
                 topk_global_id=res["topk_global_id"], topk_scores=res["topk_scores"],
                 ranks=res["ranks"], metrics=res["metrics"])


def golden_pipeline(ref) -> None:
    """AllScoresPipeline / AllScoresBESS of the unmodified reference (pipeline.py, bess.py:924-1062)
    under the replica emulation: the parameter grid of the reference's own tests/test_pipeline.py
    (corruption scheme x filters x candidate subset) at a fixture-sized shape, for ComplEx (GEMM
    path) and TransE-L1 (tile path).  n_entity is not a multiple of n_shard (padding entities)
    and window_size does not divide the shard (clamped last block)."""
    n_entity, n_rel, n_shard, n_triple, bps, shard_bs, d, window, k = 2003, 11, 4, 450, 2, 64, 16, 190, 7
    sh = ref.sharding.Sharding.create(n_entity, n_shard, seed=SEED)
    cand_ents = np.arange(0, n_entity - 1, step=3)
    for fam, p in (("ComplEx", 0), ("TransE", 1)):
        for scheme in ("t", "h"):
            for filt, extra_only in ((True, True), (True, False), (False, False)):
                for use_cand in (True, False):
                    rng = np.random.default_rng(SEED + 7)
                    h = rng.choice(n_entity - 1, size=n_triple, replace=False)
                    t = rng.choice(n_entity - 1, size=n_triple, replace=False)
                    r = rng.integers(n_rel, size=n_triple)
                    triples = np.stack([h, r, t], axis=1)
                    ds = ref.dataset.KGDataset(
                        n_entity=n_entity, n_relation_type=n_rel, entity_dict=None, relation_dict=None,
                        type_offsets=None, triples={"test": triples},
                        original_triple_ids={"test": np.arange(n_triple)})
                    mode = "h_shard" if scheme == "t" else "t_shard"
                    pts = ref.sharding.PartitionedTripleSet.create_from_dataset(ds, "test", sh, mode)
                    gen = torch.Generator().manual_seed(SEED)
                    W = FAMILIES[fam]["ew"] * d
                    ent = torch.randn(n_entity, W, generator=gen)
                    rel = torch.randn(n_rel, FAMILIES[fam]["rw"](d), generator=gen)
                    sf = build_score_fn(ref, fam, True, p, sh, n_rel, d, ent, rel)
                    ns = ref.negative_sampler.PlaceholderNegativeSampler(corruption_scheme=scheme, seed=SEED)
                    bs = ref.batch_sampler.RigidShardedBatchSampler(
                        pts, ns, shard_bs=shard_bs, batches_per_step=bps, seed=SEED,
                        return_triple_idx=True)
                    ev = ref.metric.Evaluation(["mrr", "hits@10"], mode="average", reduction="sum",
                                               return_ranks=True)
                    gt_col = 0 if scheme == "h" else 2
                    to_filter = None
                    if filt:
                        extra = np.copy(triples)
                        extra[:, gt_col] += 1
                        to_filter = [extra] if extra_only else [triples, extra]
                    pipe = ref.pipeline.AllScoresPipeline(
                        bs, scheme, sf, ev, filter_triples=to_filter,
                        candidate_ents=cand_ents if use_cand else None, return_scores=True,
                        return_topk=True, k=k, window_size=window, use_ipu_model=True)
                    with torch.no_grad():
                        out = pipe()
                        # one raw block of AllScoresBESS (the clamped last one) for the first batch
                        batch = bs[list(bs.get_dataloader_sampler(shuffle=False))[0]]
                        for key in ("triple_mask", "triple_idx"):
                            batch.pop(key)
                        batch.pop("head" if scheme == "h" else "tail")
                        last = pipe.bess_module.n_step - 1
                        step = torch.full((bps, n_shard, 1), last, dtype=torch.int32)
                        block = ref_loader.run_replicated(
                            pipe.bess_module, dict(batch, step=step), n_shard, bps)["__tensor__"]
                    sc = out["scores"]
                    finite = torch.isfinite(sc)
                    tag = f"pipeline_{fam}_{scheme}_f{int(filt)}{int(extra_only)}_c{int(use_cand)}"
                    save(tag, dict(family=fam, p=p, scheme=scheme, d=d, n_rel=n_rel, n_entity=n_entity,
                                   n_shard=n_shard, n_triple=n_triple, bps=bps, shard_bs=shard_bs,
                                   window=window, k=k, filter=filt, extra_only=extra_only,
                                   use_candidates=use_cand, seed=SEED, score_col_stride=61,
                                   block_row_stride=4, block_col_stride=7),
                         triples=triples, cand_ents=cand_ents,
                         # tables are torch.randn(generator seeded SEED): regenerated by the test, pinned here
                         ent_probe=ent[::97, ::5], rel_probe=rel[:, ::5],
                         scores_sub=sc[:, ::61], neg_inf_count=(~finite).sum(dim=1),
                         topk_global_id=out["topk_global_id"], triple_idx=out["triple_idx"],
                         ranks=out["ranks"], mrr=out["metrics"]["mrr"], hits10=out["metrics"]["hits@10"],
                         mrr_avg=out["metrics_avg"]["mrr"],
                         **{f"in_{kk}": v for kk, v in batch.items()}, block_last_sub=block[::4, ::7])


def dataset_frames():
    """Labelled triples used by golden_dataset and tests/test_host_golden.py (seeded)."""
    import pandas as pd
    rng = np.random.default_rng(SEED)
    ents = [f"e{i:03d}" for i in range(60)]
    types = {e: f"type{rng.integers(4)}" for e in ents}
    rels = [f"r{i}" for i in range(7)]
    n = 400
    df = pd.DataFrame(dict(h=rng.choice(ents, n), r=rng.choice(rels, n), t=rng.choice(ents, n)))
    return df, types


def golden_dataset(ref) -> None:
    """KGDataset.from_triples / from_dataframe of the reference (dataset.py:83-239)."""
    df, types = dataset_frames()
    arrays = {}
    for tag, kw in (("typed", dict(entity_types=types)), ("plain", dict())):
        ds = ref.dataset.KGDataset.from_dataframe(df, "h", "r", "t", seed=SEED, **kw)
        for part in ("train", "valid", "test"):
            arrays[f"{tag}_{part}"] = ds.triples[part]
            arrays[f"{tag}_ids_{part}"] = ds.original_triple_ids[part]
        arrays[f"{tag}_entity_dict"] = np.array(ds.entity_dict)
        arrays[f"{tag}_relation_dict"] = np.array(ds.relation_dict)
        if ds.type_offsets is not None:
            arrays[f"{tag}_type_names"] = np.array(list(ds.type_offsets.keys()))
            arrays[f"{tag}_type_offsets"] = np.array(list(ds.type_offsets.values()))
    parts = {"train": df.iloc[:300], "valid": df.iloc[300:]}
    ds = ref.dataset.KGDataset.from_dataframe(parts, "h", "r", "t", entity_types=types)
    arrays.update(split_train=ds.triples["train"], split_valid=ds.triples["valid"],
                  split_entity_dict=np.array(ds.entity_dict))
    rng = np.random.default_rng(SEED + 1)
    raw = np.stack([rng.integers(50, size=333), rng.integers(5, size=333),
                    rng.integers(50, size=333)], axis=1).astype(np.int32)
    ds = ref.dataset.KGDataset.from_triples(raw, split=(0.6, 0.3, 0.1), seed=7)
    arrays.update(raw=raw, raw_train=ds.triples["train"], raw_valid=ds.triples["valid"],
                  raw_test=ds.triples["test"], raw_ids_test=ds.original_triple_ids["test"])
    save("host_dataset", dict(seed=SEED, n_entity=int(ds.n_entity), n_rel=int(ds.n_relation_type)),
         **arrays)


GENERATORS = dict(host=golden_host, dataset=golden_dataset, scores=golden_scores, loss=golden_loss,
                  metric=golden_metric, bess=golden_bess, train=golden_train, topk=golden_topk,
                  pipeline=golden_pipeline, bess_shared=golden_bess_shared)


def main() -> None:
    """`python tests/golden/make_golden.py [name ...] [--only Family,...]` — all generators, or
    the named ones; --only restricts scores / bess / train to the given score-function families
    (how fixtures of a newly added family are generated without touching the others)."""
    ref = ref_loader.load_reference()
    args = sys.argv[1:]
    only = None
    if "--only" in args:
        i = args.index("--only")
        only = args[i + 1].split(",")
        args = args[:i] + args[i + 2:]
    for name in (args or list(GENERATORS)):
        torch.manual_seed(SEED)
        if only is not None and name in ("scores", "bess", "train"):
            GENERATORS[name](ref, only=only)
        else:
            GENERATORS[name](ref)


if __name__ == "__main__":
    main()
