"""Host sampler vs DeviceBatchFeed: time to produce one step's batch ON THE DEVICE (cfg-5 shape:
wikikg2-shaped graph, 2048 queries per step, 500 predetermined candidate tails per query).
usage: python scripts/feed_bench.py [n_shard]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from besskge_b200.batch_sampler import RigidShardedBatchSampler  # noqa: E402
from besskge_b200.dataset import synthetic_kg  # noqa: E402
from besskge_b200.device_feed import DeviceBatchFeed  # noqa: E402
from besskge_b200.negative_sampler import TripleBasedShardedNegativeSampler  # noqa: E402
from besskge_b200.sharding import PartitionedTripleSet, Sharding  # noqa: E402


def build(n, S, steps):
    ds = synthetic_kg("ogbl-wikikg2", seed=1234, n_triple=n * S * steps)
    ds.neg_tails = {"train": np.random.default_rng(4321).integers(
        ds.n_entity, size=(n * S * steps, 500), dtype=np.int32)}
    sh = Sharding.create(ds.n_entity, n, seed=1234)
    pts = PartitionedTripleSet.create_from_dataset(ds, "train", sh)
    ns = TripleBasedShardedNegativeSampler(pts.neg_heads, pts.neg_tails, sh, "t", 1234)
    return RigidShardedBatchSampler(pts, ns, shard_bs=S, batches_per_step=1, seed=1234)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    S, steps = 2048, 12  # queries per shard per step
    dev = torch.device("cuda", 0)
    host = build(n, S, steps)
    feed = DeviceBatchFeed(build(n, S, steps), dev)
    size, span = host.partition_sample_size, len(host)
    idxs = [[(i * size + j) % span for j in range(size)] for i in range(steps)]
    res = {}
    for name, fn in (("host sampler + pinned H2D", lambda ix: {k: v.pin_memory().to(dev, non_blocking=True)
                                                               for k, v in host[ix].items()}),
                     ("DeviceBatchFeed", lambda ix: feed[ix])):
        for ix in idxs[:2]:
            fn(ix)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for ix in idxs[2:]:
            b = fn(ix)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / (steps - 2)
        res[name] = dt
        print(f"{name:28s} {dt * 1e3:8.3f} ms / step  ({n * S / dt / 1e6:.2f} M queries/s feed rate), "
              f"negative {tuple(b['negative'].shape)}")
    print(f"speed-up {res['host sampler + pinned H2D'] / res['DeviceBatchFeed']:.1f}x")


if __name__ == "__main__":
    main()
