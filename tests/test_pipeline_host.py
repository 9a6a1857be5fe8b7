"""CPU: host-side logic of AllScoresPipeline / AllScoresBESS (validation, column map, filter index,
no CPU fallback).  The device path is covered by tests/test_gpu_pipeline.py."""
import numpy as np
import pytest
import torch

import besskge_b200
from besskge_b200.batch_sampler import RigidShardedBatchSampler
from besskge_b200.bess import AllScoresBESS
from besskge_b200.dataset import KGDataset
from besskge_b200.metric import Evaluation
from besskge_b200.negative_sampler import PlaceholderNegativeSampler, RandomShardedNegativeSampler
from besskge_b200.pipeline import AllScoresPipeline
from besskge_b200.scoring import ComplEx, TransE
from besskge_b200.sharding import PartitionedTripleSet, Sharding


def _setup(scheme="t", return_idx=True):
    n_ent, n_rel, n = 211, 5, 4
    rng = np.random.default_rng(0)
    triples = np.stack([rng.integers(n_ent, size=300), rng.integers(n_rel, size=300),
                        rng.integers(n_ent, size=300)], axis=1).astype(np.int32)
    ds = KGDataset(n_ent, n_rel, {"test": triples})
    sh = Sharding.create(n_ent, n, seed=0)
    mode = "h_shard" if scheme == "t" else "t_shard"
    pts = PartitionedTripleSet.create_from_dataset(ds, "test", sh, partition_mode=mode)
    bs = RigidShardedBatchSampler(pts, PlaceholderNegativeSampler(scheme), shard_bs=16,
                                  batches_per_step=2, seed=0, return_triple_idx=return_idx)
    return ds, sh, bs, triples


def test_pipeline_validation_and_column_map():
    ds, sh, bs, triples = _setup("t")
    sf = ComplEx(True, sh, ds.n_relation_type, 8)
    ev = Evaluation(["mrr", "hits@10"], reduction="sum")
    with pytest.raises(ValueError):
        AllScoresPipeline(bs, "t", sf, None, return_scores=False)  # nothing to return
    with pytest.raises(ValueError):
        AllScoresPipeline(bs, "h", sf, ev)  # 'h' needs t_shard-partitioned triples
    with pytest.raises(ValueError):
        AllScoresPipeline(bs, "ht", sf, ev)
    with pytest.raises(ValueError):
        AllScoresPipeline(bs, "t", sf, ev, return_topk=True, k=65)
    _, _, bs_noidx, _ = _setup("t", return_idx=False)
    with pytest.raises(ValueError):
        AllScoresPipeline(bs_noidx, "t", sf, ev, filter_triples=[triples])
    cand = np.arange(0, ds.n_entity, 3)
    pipe = AllScoresPipeline(bs, "t", sf, ev, filter_triples=[triples], candidate_ents=cand)
    # column of entity e in the [*, n_shard * Es] block scores; -1 for non-candidates
    col = pipe._col_of_entity.numpy()
    Es = sh.max_entity_per_shard
    for e in range(ds.n_entity):
        if e % 3 == 0:
            j, l = divmod(int(col[e]), Es)
            assert sh.shard_and_idx_to_entity[j, l] == e
        else:
            assert col[e] == -1
    # global ids of the partitioned triples are restored for the filter join (pipeline.py:150-171)
    order = bs.triples  # local head ids
    assert torch.equal(pipe.triples[:, 1:], torch.from_numpy(order[:, 1:]))
    assert set(map(tuple, pipe.triples.numpy().tolist())) == set(map(tuple, triples.tolist()))
    # no CPU fallback
    with pytest.raises(besskge_b200.BessLibraryError):
        pipe()


def test_all_scores_bess_validation():
    ds, sh, bs, _ = _setup("t")
    with pytest.raises(ValueError):
        AllScoresBESS(PlaceholderNegativeSampler("t"), TransE(False, 1, sh, ds.n_relation_type, 8))
    with pytest.raises(ValueError):
        AllScoresBESS(PlaceholderNegativeSampler("ht"), TransE(True, 1, sh, ds.n_relation_type, 8))
    with pytest.raises(ValueError):
        AllScoresBESS(RandomShardedNegativeSampler(4, sh, 0, "t", False, True),
                      TransE(True, 1, sh, ds.n_relation_type, 8))
    mod = AllScoresBESS(PlaceholderNegativeSampler("t"), TransE(True, 1, sh, ds.n_relation_type, 8),
                        window_size=20)
    assert mod.n_step == int(np.ceil(sh.max_entity_per_shard / 20))
    with pytest.raises(besskge_b200.BessLibraryError):
        mod(step=torch.zeros(4, 1, dtype=torch.int32), relation=torch.zeros(4, 3, dtype=torch.int32),
            head=torch.zeros(4, 3, dtype=torch.int32))
