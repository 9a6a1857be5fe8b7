"""-m gpu: AllScoresBESS / AllScoresPipeline (CUDA path) vs fixtures produced by the unmodified
reference pipeline (tests/golden/pipeline_*.npz) and vs the oracle's dense restatement."""
import numpy as np
import pytest
import torch
from numpy.testing import assert_array_equal
from torch.testing import assert_close

from oracle import besskge_oracle as O

from .conftest import golden_names
from .pipeline_cases import filter_list, load_case

pytestmark = pytest.mark.gpu


def _build(cfg, g, ent, rel):
    from besskge_b200.batch_sampler import RigidShardedBatchSampler
    from besskge_b200.dataset import KGDataset
    from besskge_b200.metric import Evaluation
    from besskge_b200.negative_sampler import PlaceholderNegativeSampler
    from besskge_b200.sharding import PartitionedTripleSet, Sharding
    from . import gpu_helpers as H
    sh = Sharding.create(cfg["n_entity"], cfg["n_shard"], seed=cfg["seed"])
    ds = KGDataset(n_entity=cfg["n_entity"], n_relation_type=cfg["n_rel"],
                   triples={"test": g["triples"]},
                   original_triple_ids={"test": np.arange(cfg["n_triple"])})
    mode = "h_shard" if cfg["scheme"] == "t" else "t_shard"
    pts = PartitionedTripleSet.create_from_dataset(ds, "test", sh, partition_mode=mode)
    sf = H.make_score_fn(cfg["family"], True, cfg["p"], sh, cfg["n_rel"], cfg["d"], ent, rel)
    ns = PlaceholderNegativeSampler(corruption_scheme=cfg["scheme"], seed=cfg["seed"])
    bs = RigidShardedBatchSampler(pts, ns, shard_bs=cfg["shard_bs"], batches_per_step=cfg["bps"],
                                  seed=cfg["seed"], return_triple_idx=True)
    ev = Evaluation(["mrr", "hits@10"], mode="average", reduction="sum", return_ranks=True)
    return sh, pts, sf, bs, ev


@pytest.mark.parametrize("name", golden_names("pipeline_"))
def test_all_scores_pipeline_vs_reference_golden(name):
    from besskge_b200.pipeline import AllScoresPipeline
    from .test_oracle_golden import score_cfg
    cfg, g, ent, rel = load_case(name)
    sh, pts, sf, bs, ev = _build(cfg, g, ent, rel)
    pipe = AllScoresPipeline(bs, cfg["scheme"], sf, ev, filter_triples=filter_list(cfg, g["triples"]),
                             candidate_ents=g["cand_ents"] if cfg["use_candidates"] else None,
                             return_scores=True, return_topk=True, k=cfg["k"], window_size=cfg["window"])
    pipe.bess_module.device_window = 96  # several GEMM / tile windows per shard (501 rows)
    out = pipe()
    torch.cuda.synchronize()
    # bookkeeping: bit-exact
    assert_array_equal(out["triple_idx"].numpy(), g["triple_idx"])
    sc = out["scores"]
    assert sc.shape == (cfg["n_triple"], cfg["n_entity"])
    want = torch.from_numpy(g["scores_sub"])
    got = sc[:, ::cfg["score_col_stride"]]
    assert torch.equal(torch.isinf(got), torch.isinf(want))
    fin = torch.isfinite(want) & (want > -1e30)
    assert_close(got[fin], want[fin], rtol=1e-5, atol=1e-5)
    assert torch.equal(got[~fin], want[~fin])  # -inf masks and the reference's -FLT_MAX quirk
    assert_array_equal((~torch.isfinite(sc)).sum(dim=1).numpy(), g["neg_inf_count"])
    # ranks / top-k: exact except where two scores are closer than the score tolerance
    tr = torch.from_numpy(g["triples"][pts.triple_sort_idx[g["triple_idx"]]])
    fl = filter_list(cfg, g["triples"])
    c = score_cfg(cfg["family"], cfg["d"], dict(p=cfg["p"]))
    ora = O.all_scores_pipeline(c, ent, rel, tr, cfg["scheme"],
                                None if fl is None else torch.from_numpy(np.concatenate(fl, axis=0)),
                                g["cand_ents"] if cfg["use_candidates"] else None, cfg["k"])
    want_ranks = torch.from_numpy(g["ranks"])
    diff = (out["ranks"] - want_ranks).abs()
    assert bool((diff <= 1).all()) and int((diff != 0).sum()) <= cfg["n_triple"] // 100
    top_sc = torch.gather(ora["scores"], 1, torch.from_numpy(g["topk_global_id"]))
    tol = 1e-5 + 1e-5 * top_sc.abs()
    clean = torch.ones_like(top_sc, dtype=torch.bool)
    gap_ok = (top_sc[:, :-1] - top_sc[:, 1:]) > tol[:, 1:]
    clean[:, :-1] &= gap_ok
    clean[:, 1:] &= gap_ok
    clean[:, -1] = False
    assert int(clean.sum()) > 0.8 * clean[:, :-1].numel()
    assert bool((out["topk_global_id"][clean] == torch.from_numpy(g["topk_global_id"])[clean]).all())
    if int((diff != 0).sum()) == 0:
        assert_close(out["metrics"]["mrr"], torch.from_numpy(g["mrr"]), rtol=1e-6, atol=1e-6)
        assert_close(out["metrics"]["hits@10"], torch.from_numpy(g["hits10"]))
        assert_close(out["metrics_avg"]["mrr"], torch.from_numpy(g["mrr_avg"]), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["pipeline_ComplEx_t_f00_c0", "pipeline_ComplEx_h_f00_c0",
                                  "pipeline_TransE_t_f00_c0", "pipeline_TransE_h_f00_c0"])
def test_all_scores_bess_block_vs_reference_golden(name):
    """AllScoresBESS.forward(step) — the reference's per-block call, last (clamped) block"""
    from besskge_b200.bess import AllScoresBESS
    from besskge_b200.negative_sampler import PlaceholderNegativeSampler
    cfg, g, ent, rel = load_case(name)
    sh, pts, sf, bs, ev = _build(cfg, g, ent, rel)
    mod = AllScoresBESS(PlaceholderNegativeSampler(cfg["scheme"]), sf, cfg["window"])
    mod.device_window = 64
    batch = {k[3:]: torch.from_numpy(v).flatten(end_dim=1) for k, v in g.items() if k.startswith("in_")}
    n, bps = cfg["n_shard"], cfg["bps"]
    step = torch.full((bps * n, 1), mod.n_step - 1, dtype=torch.int32)
    got = mod(step=step, **batch).cpu()
    assert got.shape == (bps * n * cfg["shard_bs"], n * cfg["window"])
    want = torch.from_numpy(g["block_last_sub"])
    assert_close(got[::4, ::7], want, rtol=1e-5, atol=1e-5)


def test_all_scores_pipeline_validation():
    from besskge_b200.pipeline import AllScoresPipeline
    cfg, g, ent, rel = load_case("pipeline_TransE_t_f00_c0")
    sh, pts, sf, bs, ev = _build(cfg, g, ent, rel)
    with pytest.raises(ValueError):
        AllScoresPipeline(bs, "t", sf, None, return_scores=False)
    with pytest.raises(ValueError):
        AllScoresPipeline(bs, "h", sf, ev)  # 'h' needs t_shard-partitioned triples
    with pytest.raises(ValueError):
        AllScoresPipeline(bs, "ht", sf, ev)
