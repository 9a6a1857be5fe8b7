"""Host-side product classes (numpy) vs fixtures produced by the unmodified
reference: sharding, partitioning, negative samplers, batch samplers must be
BIT-EXACT (reference tests: tests/test_sharding.py, test_negative_sampler.py,
test_batch_sampler.py)."""
import numpy as np
import pytest
from numpy.testing import assert_array_equal

from besskge_b200.batch_sampler import RandomShardedBatchSampler, RigidShardedBatchSampler
from besskge_b200.dataset import KGDataset
from besskge_b200.negative_sampler import (
    RandomShardedNegativeSampler,
    TripleBasedShardedNegativeSampler,
    TypeBasedShardedNegativeSampler,
)
from besskge_b200.sharding import PartitionedTripleSet, Sharding

from .conftest import load_golden


def _dataset(cfg, g, neg_heads=None, neg_tails=None):
    return KGDataset(
        n_entity=cfg["n_entity"], n_relation_type=cfg["n_rel"],
        triples={"test": g["triples"]},
        original_triple_ids={"test": np.arange(g["triples"].shape[0])},
        type_offsets={str(i): int(o) for i, o in enumerate(cfg["type_offsets"])},
        neg_heads={"test": g["neg_heads"] if neg_heads is None else neg_heads},
        neg_tails={"test": g["neg_tails"] if neg_tails is None else neg_tails},
    )


def test_sharding_bit_exact():
    cfg, g = load_golden("host_sharding")
    sh = Sharding.create(cfg["n_entity"], cfg["n_shard"], seed=cfg["seed"],
                         type_offsets=np.array(cfg["type_offsets"]))
    for k in ("entity_to_shard", "entity_to_idx", "shard_and_idx_to_entity", "shard_counts",
              "entity_type_counts", "entity_type_offsets"):
        assert_array_equal(getattr(sh, k), g[k])
        assert getattr(sh, k).dtype == g[k].dtype
    sh2 = Sharding.create(500, 4, seed=cfg["seed"])
    assert_array_equal(sh2.entity_to_shard, g["s2_entity_to_shard"])
    assert_array_equal(sh2.entity_to_idx, g["s2_entity_to_idx"])
    assert_array_equal(sh2.shard_and_idx_to_entity, g["s2_table"])
    assert_array_equal(sh2.shard_counts, g["s2_counts"])
    # invariants of reference tests/test_sharding.py:43-72
    assert sh.shard_counts.sum() == cfg["n_entity"]
    assert_array_equal(sh.shard_and_idx_to_entity[sh.entity_to_shard, sh.entity_to_idx],
                       np.arange(cfg["n_entity"]))
    assert_array_equal(sh.entity_type_counts.sum(-1), sh.shard_counts)


def test_sharding_save_load(tmp_path):
    sh = Sharding.create(101, 3, seed=7)
    sh.save(tmp_path / "s.npz")
    sh2 = Sharding.load(tmp_path / "s.npz")
    assert sh2.n_shard == 3
    assert_array_equal(sh.shard_and_idx_to_entity, sh2.shard_and_idx_to_entity)
    assert sh2.entity_type_counts is None


def test_sharding_load_never_unpickles(tmp_path):
    """None fields are absent from the file; a reference-style file that stores None as an
    object array loads with the field None instead of demanding allow_pickle."""
    sh = Sharding.create(101, 3, seed=7, type_offsets=np.array([0, 40, 70]))
    sh.save(tmp_path / "typed.npz")
    with np.load(tmp_path / "typed.npz", allow_pickle=False) as f:
        assert all(f[k].dtype != object for k in f.files)
    sh2 = Sharding.load(tmp_path / "typed.npz")
    assert_array_equal(sh.entity_type_counts, sh2.entity_type_counts)
    assert_array_equal(sh.entity_type_offsets, sh2.entity_type_offsets)
    plain = Sharding.create(101, 3, seed=7)
    import dataclasses
    np.savez(tmp_path / "ref_style.npz", **dataclasses.asdict(plain))  # what the reference writes
    sh3 = Sharding.load(tmp_path / "ref_style.npz")
    assert sh3.entity_type_counts is None and sh3.entity_type_offsets is None
    assert_array_equal(plain.entity_to_idx, sh3.entity_to_idx)


def test_dataset_constructors_vs_reference_golden(tmp_path):
    """KGDataset.from_dataframe / from_triples (dataset.py:83-239): ids, type offsets and
    the random split equal the reference's on the same labelled triples."""
    from besskge_b200.dataset import KGDataset
    from .golden.make_golden import dataset_frames
    cfg, g = load_golden("host_dataset")
    df, types = dataset_frames()
    for tag, kw in (("typed", dict(entity_types=types)), ("plain", dict())):
        ds = KGDataset.from_dataframe(df, "h", "r", "t", seed=cfg["seed"], **kw)
        for part in ("train", "valid", "test"):
            assert_array_equal(ds.triples[part], g[f"{tag}_{part}"])
            assert_array_equal(ds.original_triple_ids[part], g[f"{tag}_ids_{part}"])
        assert ds.entity_dict == g[f"{tag}_entity_dict"].tolist()
        assert ds.relation_dict == g[f"{tag}_relation_dict"].tolist()
        if tag == "typed":
            assert list(ds.type_offsets.keys()) == g["typed_type_names"].tolist()
            assert list(ds.type_offsets.values()) == g["typed_type_offsets"].tolist()
            assert ds.ht_types["train"].shape == (ds.triples["train"].shape[0], 2)
        else:
            assert ds.type_offsets is None
    parts = {"train": df.iloc[:300], "valid": df.iloc[300:]}
    ds = KGDataset.from_dataframe(parts, "h", "r", "t", entity_types=types)
    assert_array_equal(ds.triples["train"], g["split_train"])
    assert_array_equal(ds.triples["valid"], g["split_valid"])
    assert ds.entity_dict == g["split_entity_dict"].tolist()
    raw = KGDataset.from_triples(g["raw"], split=(0.6, 0.3, 0.1), seed=7)
    for part in ("train", "valid", "test"):
        assert_array_equal(raw.triples[part], g[f"raw_{part}"])
    assert_array_equal(raw.original_triple_ids["test"], g["raw_ids_test"])
    assert (raw.n_entity, raw.n_relation_type) == (cfg["n_entity"], cfg["n_rel"])
    # save / load round trip (data-only container, no pickle)
    ds.save(tmp_path / "ds.npz")
    back = KGDataset.load(tmp_path / "ds.npz")
    assert back.entity_dict == ds.entity_dict and back.type_offsets == ds.type_offsets
    assert back.neg_heads is None
    for part in ds.triples:
        assert_array_equal(back.triples[part], ds.triples[part])
    with pytest.raises(ValueError):
        np.savez(tmp_path / "other.npz", a=np.zeros(3))
        KGDataset.load(tmp_path / "other.npz")


@pytest.mark.parametrize("mode", ["h_shard", "t_shard", "ht_shardpair"])
@pytest.mark.parametrize("inv", [False, True])
def test_partition_bit_exact(mode, inv):
    cfg, g = load_golden("host_sharding")
    sh = Sharding.create(cfg["n_entity"], cfg["n_shard"], seed=cfg["seed"],
                         type_offsets=np.array(cfg["type_offsets"]))
    pts = PartitionedTripleSet.create_from_dataset(_dataset(cfg, g), "test", sh, mode,
                                                   add_inverse_triples=inv)
    tag = f"part_{mode}_{int(inv)}"
    assert_array_equal(pts.triples, g[f"{tag}_triples"])
    assert_array_equal(pts.triple_counts, g[f"{tag}_counts"])
    assert_array_equal(pts.triple_offsets, g[f"{tag}_offsets"])
    assert_array_equal(pts.triple_sort_idx, g[f"{tag}_sort"])
    assert_array_equal(pts.types, g[f"{tag}_types"])
    assert_array_equal(pts.neg_heads, g[f"{tag}_nh"])
    assert_array_equal(pts.neg_tails, g[f"{tag}_nt"])
    assert pts.dummy == "none"


def test_partition_queries_bit_exact():
    cfg, g = load_golden("host_sharding")
    sh = Sharding.create(cfg["n_entity"], cfg["n_shard"], seed=cfg["seed"],
                         type_offsets=np.array(cfg["type_offsets"]))
    nq = cfg["n_query"]
    pts = PartitionedTripleSet.create_from_queries(
        _dataset(cfg, g), sh, g["triples"][:nq, :2], "hr", ground_truth=g["triples"][:nq, 2])
    assert pts.partition_mode == "h_shard" and pts.dummy is None
    assert_array_equal(pts.triples, g["q_triples"])
    assert_array_equal(pts.triple_counts, g["q_counts"])
    assert_array_equal(pts.triple_offsets, g["q_offsets"])
    assert_array_equal(pts.triple_sort_idx, g["q_sort"])
    assert_array_equal(pts.neg_tails, g["q_nt"])
    with pytest.raises(ValueError):
        PartitionedTripleSet.create_from_queries(_dataset(cfg, g), sh, g["triples"][:4, :2], "xx")
    with pytest.raises(ValueError):
        PartitionedTripleSet.partition_triples(g["triples"], sh, "bad_mode")


def _check_batch(batch, g, tag, it):
    keys = [k[len(f"{tag}_{it}_"):] for k in g if k.startswith(f"{tag}_{it}_")]
    assert set(keys) == set(batch.keys()), (sorted(keys), sorted(batch.keys()))
    for k in keys:
        got = batch[k].numpy()
        want = g[f"{tag}_{it}_{k}"]
        assert got.dtype == want.dtype, (k, got.dtype, want.dtype)
        assert_array_equal(got, want, err_msg=f"{tag} {it} {k}")


def test_samplers_bit_exact():
    cfg, g = load_golden("host_samplers")
    hcfg, hg = load_golden("host_sharding")
    sh = Sharding.create(cfg["n_entity"], cfg["n_shard"], seed=cfg["seed"],
                         type_offsets=np.array(cfg["type_offsets"]))
    ds = _dataset(hcfg, hg)
    pts = PartitionedTripleSet.create_from_dataset(ds, "test", sh, "ht_shardpair")
    rng_trip = {}
    for case in cfg["cases"]:
        tag = case["tag"]
        if case["sampler"] == "random":
            ns = RandomShardedNegativeSampler(7, sh, cfg["seed"], case["scheme"], case["local"],
                                              case["flat"])
            bs = RandomShardedBatchSampler(pts, ns, shard_bs=20, batches_per_step=3,
                                           seed=cfg["seed"],
                                           hrt_freq_weighting=(case["scheme"] == "t"),
                                           weight_smoothing=0.5)
            for it in range(2):
                _check_batch(bs[[it]], g, tag, it)
        elif case["sampler"] == "type":
            ns = TypeBasedShardedNegativeSampler(pts.types, 5, sh, case["scheme"], False,
                                                 cfg["seed"])
            bs = RandomShardedBatchSampler(pts, ns, shard_bs=20, batches_per_step=2,
                                           seed=cfg["seed"])
            _check_batch(bs[[0]], g, tag, 0)
        else:
            flat = case["flat"]
            if flat not in rng_trip:
                rng = np.random.default_rng(cfg["seed"] + 1)
                n_e, n_r, n_t, n_n = cfg["n_entity"], cfg["n_rel"], cfg["n_triple"], cfg["n_neg"]
                h = rng.integers(n_e, size=n_t)
                t = rng.integers(n_e, size=n_t)
                r = rng.integers(n_r, size=n_t)
                outer = 1 if flat else n_t
                nh = rng.integers(n_e, size=(outer, n_n), dtype=np.int32)
                nt = rng.integers(n_e, size=(outer, n_n), dtype=np.int32)
                g2 = dict(triples=np.stack([h, r, t], axis=1), neg_heads=nh, neg_tails=nt)
                ds2 = _dataset(hcfg, g2)
                rng_trip[flat] = PartitionedTripleSet.create_from_dataset(ds2, "test", sh,
                                                                          "ht_shardpair")
            pts2 = rng_trip[flat]
            ns = TripleBasedShardedNegativeSampler(pts2.neg_heads, pts2.neg_tails, sh,
                                                   case["scheme"], cfg["seed"],
                                                   mask_on_gather=case["mog"], return_sort_idx=True)
            bs = RigidShardedBatchSampler(pts2, ns, shard_bs=20, batches_per_step=2,
                                          seed=cfg["seed"], duplicate_batch=(case["scheme"] == "ht"),
                                          return_triple_idx=True)
            sampler = list(bs.get_dataloader_sampler(shuffle=False))
            assert_array_equal(np.array([len(bs), len(sampler)]), g[f"{tag}_len"])
            for it in (0, case["last"]):
                _check_batch(bs[sampler[it]], g, tag, it)


def test_random_negatives_in_range():
    """reference tests/test_negative_sampler.py:30-56."""
    sh = Sharding.create(501, 5, seed=1234)
    ns = RandomShardedNegativeSampler(11, sh, 3, "t", False, False)
    idx = np.zeros((2, 5, 5, 4), dtype=np.int64)
    neg = ns(idx)["negative_entities"]
    assert neg.shape == (2, 5, 5, 20, 11)
    assert np.all(neg < sh.shard_counts[None, :, None, None, None]) and np.all(neg >= 0)


def test_prefetch_loader_order():
    cfg, g = load_golden("host_sharding")
    sh = Sharding.create(cfg["n_entity"], cfg["n_shard"], seed=cfg["seed"])
    pts = PartitionedTripleSet.create_from_dataset(_dataset(cfg, g), "test", sh, "ht_shardpair")
    mk = lambda: RandomShardedBatchSampler(
        pts, RandomShardedNegativeSampler(3, sh, 5, "t", False, True), 20, 2, seed=5)
    a, b = mk(), mk()
    direct = [a[idx] for idx in a.get_dataloader_sampler()]
    loaded = list(b.get_dataloader(shuffle=False, buffer_size=2, pin_memory=False))
    assert len(direct) == len(loaded) > 0
    for x, y in zip(direct, loaded):
        for k in x:
            assert_array_equal(x[k].numpy(), y[k].numpy())
