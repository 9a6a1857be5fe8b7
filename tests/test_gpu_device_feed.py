"""-m gpu: DeviceBatchFeed (device-resident triples / candidate lists, host draws only the indices)
is bit-identical to ShardedBatchSampler.__getitem__, and the BESS modules accept its CUDA batches."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(scheme, sampler_kind, neg_kind, partition="ht_shardpair", dup=False, weights=False,
          mask_on_gather=False, sort_idx=False, seed=11):
    from besskge_b200.batch_sampler import RandomShardedBatchSampler, RigidShardedBatchSampler
    from besskge_b200.dataset import KGDataset
    from besskge_b200.negative_sampler import (PlaceholderNegativeSampler,
                                               RandomShardedNegativeSampler,
                                               TripleBasedShardedNegativeSampler)
    from besskge_b200.sharding import PartitionedTripleSet, Sharding
    n_ent, n_rel, n_shard, n_trip, n_neg = 403, 7, 4, 900, 13
    rng = np.random.default_rng(seed)
    triples = np.stack([rng.integers(n_ent, size=n_trip), rng.integers(n_rel, size=n_trip),
                        rng.integers(n_ent, size=n_trip)], axis=1).astype(np.int32)
    ds = KGDataset(n_ent, n_rel, {"train": triples},
                   neg_heads={"train": rng.integers(n_ent, size=(n_trip, n_neg), dtype=np.int32)},
                   neg_tails={"train": rng.integers(n_ent, size=(n_trip, n_neg), dtype=np.int32)})
    sh = Sharding.create(n_ent, n_shard, seed=seed)
    pts = PartitionedTripleSet.create_from_dataset(ds, "train", sh, partition_mode=partition)
    if neg_kind == "random":
        ns = RandomShardedNegativeSampler(5, sh, seed, scheme, False, True)
    elif neg_kind == "triple":
        ns = TripleBasedShardedNegativeSampler(pts.neg_heads, pts.neg_tails, sh, scheme, seed,
                                               mask_on_gather=mask_on_gather, return_sort_idx=sort_idx)
    else:
        ns = PlaceholderNegativeSampler(scheme, seed)
    cls = RandomShardedBatchSampler if sampler_kind == "random" else RigidShardedBatchSampler
    bs = cls(pts, ns, shard_bs=24, batches_per_step=2, seed=seed, duplicate_batch=dup,
             hrt_freq_weighting=weights, return_triple_idx=True)
    return bs, sh, n_rel


CASES = [
    ("t", "random", "random", "ht_shardpair", False, False, False, False),
    ("ht", "random", "random", "ht_shardpair", True, True, False, False),
    ("t", "rigid", "triple", "ht_shardpair", False, False, False, False),
    ("h", "rigid", "triple", "ht_shardpair", False, False, True, True),
    ("ht", "rigid", "triple", "ht_shardpair", True, False, False, False),  # host candidate split
    ("t", "rigid", "placeholder", "h_shard", False, False, False, False),
    ("h", "random", "triple", "t_shard", False, True, False, False),
]


@pytest.mark.parametrize("scheme,sk,nk,part,dup,weights,mog,srt", CASES)
def test_device_feed_bit_identical_to_host_sampler(scheme, sk, nk, part, dup, weights, mog, srt):
    from besskge_b200.device_feed import DeviceBatchFeed
    host, _, _ = _make(scheme, sk, nk, part, dup, weights, mog, srt)
    twin, _, _ = _make(scheme, sk, nk, part, dup, weights, mog, srt)
    feed = DeviceBatchFeed(twin, torch.device("cuda"))
    size = host.partition_sample_size
    for step in range(3):  # the RNG streams must stay aligned over several steps
        idx = [(step * size + j) % len(host) for j in range(size)] if sk == "rigid" else [step]
        want = host[idx]
        got = feed[idx]
        torch.cuda.synchronize()
        assert set(got) == set(want), (sorted(got), sorted(want))
        for k, v in want.items():
            g = got[k]
            assert g.is_cuda and g.dtype == v.dtype and tuple(g.shape) == tuple(v.shape), k
            assert torch.equal(g.cpu(), v), k


def test_modules_accept_device_batches():
    """the same training step from a host batch and from a device-fed batch"""
    from besskge_b200.bess import EmbeddingMovingBessKGE, training_model
    from besskge_b200.device_feed import DeviceBatchFeed
    from besskge_b200.loss import LogSigmoidLoss
    from besskge_b200.optim import SGD
    from besskge_b200.scoring import TransE
    outs = []
    for use_feed in (False, True):
        bs, sh, n_rel = _make("t", "random", "random")
        torch.manual_seed(0)
        sf = TransE(True, 1, sh, n_rel, 32).cuda()
        model = EmbeddingMovingBessKGE(bs.negative_sampler, sf, loss_fn=LogSigmoidLoss(2.0, True))
        step = training_model(model, SGD(lr=0.1), cuda_graph=False)
        src = DeviceBatchFeed(bs, torch.device("cuda")) if use_feed else bs
        for i in range(2):
            b = {k: v.flatten(end_dim=1) for k, v in src[[i]].items() if k != "triple_idx"}
            res = step(**b)
        torch.cuda.synchronize()
        outs.append((res["loss"].cpu(), sf.entity_embedding.detach().cpu().clone()))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
