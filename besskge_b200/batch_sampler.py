"""Sharded batch samplers (host side, numpy) — drop-in for reference
`besskge/batch_sampler.py`.

`__getitem__` returns the dict of index tensors one device step consumes:
head / relation / tail int32 [bps, n, n, p] (tail already transposed to
[bps, shard_t, shard_h, p] for the AllToAll, batch_sampler.py:163-167),
negative int32 [bps, n_src, n_dst, B, Nn], plus optional masks / weights.
All arrays are bit-exact with the reference for equal seeds at num_workers=0.

The PopTorch asynchronous DataLoader (batch_sampler.py:236-280) is replaced by
`get_dataloader`, a background-thread prefetcher that preserves the call order
on the single RNG and hands out pinned host tensors.
"""
from __future__ import annotations

import queue
import threading
import warnings
from abc import ABC, abstractmethod
from typing import Dict, Iterator, List, Optional, Union, cast

import einops
import numpy as np
import torch
from numpy.typing import NDArray

from .negative_sampler import ShardedNegativeSampler
from .sharding import PartitionedTripleSet


class ShardedBatchSampler(torch.utils.data.Dataset, ABC):
    def __init__(
        self,
        partitioned_triple_set: PartitionedTripleSet,
        negative_sampler: ShardedNegativeSampler,
        shard_bs: int,
        batches_per_step: int,
        seed: int,
        hrt_freq_weighting: bool = False,
        weight_smoothing: float = 0.0,
        duplicate_batch: bool = False,
        return_triple_idx: bool = False,
    ):
        pts = partitioned_triple_set
        self.n_shard = pts.sharding.n_shard
        self.triples = pts.triples
        self.dummy = pts.dummy
        self.triple_counts = pts.triple_counts
        self.triple_offsets = pts.triple_offsets
        self.triple_partition_mode = pts.partition_mode
        self.negative_sampler = negative_sampler
        self.shard_bs = shard_bs
        self.batches_per_step = batches_per_step
        self.duplicate_batch = duplicate_batch

        # batch_sampler.py:78-90 — a device micro-batch is n blocks of p triples
        p = shard_bs
        if self.triple_partition_mode == "ht_shardpair":
            p = int(np.ceil(shard_bs / self.n_shard))
        if duplicate_batch:
            p //= 2
        if negative_sampler.corruption_scheme == "ht":
            p = (p // 2) * 2
        self.positive_per_partition = p
        self.partition_sample_size = batches_per_step * p

        self.hrt_freq_weighting = hrt_freq_weighting
        self.return_triple_idx = return_triple_idx
        self.seed = seed
        self.rng = np.random.default_rng(seed)

        if hrt_freq_weighting:
            if self.dummy != "none":
                warnings.warn("hrt frequency weights are being computed on dummy entities")
            n_ent = pts.sharding.n_entity
            rel = self.triples[..., 1]
            _, hr_inv, hr_cnt = np.unique(
                self.triples[..., 0] + n_ent * rel, return_counts=True, return_inverse=True
            )
            _, rt_inv, rt_cnt = np.unique(
                self.triples[..., 2] + n_ent * rel, return_counts=True, return_inverse=True
            )
            self.hrt_weights = np.sqrt(
                1.0 / (hr_cnt[hr_inv] + rt_cnt[rt_inv] + weight_smoothing)
            )

    def __len__(self) -> int:
        size = self.partition_sample_size
        return int(np.ceil(self.triple_counts.max() / size)) * size

    def __getitem__(self, idx: List[int]) -> Dict[str, torch.Tensor]:
        """reference: batch_sampler.py:138-196."""
        sampled = self.sample_triples(idx)
        if self.duplicate_batch:
            sampled = {
                k: einops.repeat(v, "step shard ... triple -> step shard ... (2 triple)")
                for k, v in sampled.items()
            }
        sample_idx = cast(NDArray[np.int64], sampled.pop("sample_idx"))
        # np.take(..., axis=0) == triples[sample_idx] (batch_sampler.py:159-162) but ~5x faster than
        # numpy's general fancy-indexing path for a multi-dimensional index into a 2-D array
        head, relation, tail = einops.rearrange(
            np.take(self.triples, sample_idx, axis=0), "... hrt -> hrt ..."
        )
        if self.triple_partition_mode == "ht_shardpair":
            # shard_t-major so that block (shard_t, shard_h) is gathered on shard_t
            tail = einops.rearrange(
                tail, "step shard_h shard_t triple -> step shard_t shard_h triple"
            )
        batch = {
            "head": head.astype(np.int32),
            "relation": relation.astype(np.int32),
            "tail": tail.astype(np.int32),
            **sampled,
        }
        neg = self.negative_sampler(sample_idx)
        if "negative_entities" in neg:
            batch["negative"] = neg.pop("negative_entities").astype(np.int32)
        batch.update(**neg)
        if self.dummy in ("head", "tail"):
            batch.pop(self.dummy)
        if self.hrt_freq_weighting:
            w = einops.rearrange(
                self.hrt_weights[sample_idx],
                "step shard ... triple -> step shard (... triple)",
            )
            w /= np.sum(w, axis=-1, keepdims=True)
            w *= self.shard_bs
            batch["triple_weight"] = w.astype(np.float32)
        if self.return_triple_idx:
            batch["triple_idx"] = sample_idx
        return {k: torch.from_numpy(v) for k, v in batch.items()}

    @abstractmethod
    def sample_triples(
        self, idx: List[int]
    ) -> Dict[str, Union[NDArray[np.int64], NDArray[np.bool_]]]:
        """Per-partition triple indices (+ masks) for one step."""

    def get_dataloader_sampler(self, shuffle: bool) -> torch.utils.data.Sampler:
        base = (
            torch.utils.data.RandomSampler(self)
            if shuffle
            else torch.utils.data.SequentialSampler(self)
        )
        return torch.utils.data.BatchSampler(
            base, batch_size=self.partition_sample_size, drop_last=False
        )

    def get_dataloader(
        self,
        options: Optional[object] = None,
        shuffle: bool = True,
        num_workers: int = 0,
        persistent_workers: bool = False,
        buffer_size: int = 16,
        pin_memory: Optional[bool] = None,
    ) -> "PrefetchLoader":
        """Iterable over step dicts.  `options`, `num_workers` and
        `persistent_workers` are accepted for signature compatibility with the
        reference (batch_sampler.py:236-280) and ignored: one background thread
        produces batches in order so the RNG stream equals num_workers=0."""
        if pin_memory is None:
            pin_memory = torch.cuda.is_available()
        return PrefetchLoader(self, shuffle, buffer_size, pin_memory)

    @staticmethod
    def worker_init_fn(worker_id: int) -> None:
        """Re-seed both RNGs with seed+worker_id (batch_sampler.py:282-296)."""
        info = torch.utils.data.get_worker_info()
        if info:
            ds = cast(ShardedBatchSampler, info.dataset)
            ds.rng = np.random.default_rng(ds.seed + worker_id)
            ds.negative_sampler.rng = np.random.default_rng(ds.seed + worker_id)


class RigidShardedBatchSampler(ShardedBatchSampler):
    """Same positions from every partition; short partitions wrap around and
    the wrapped entries are flagged in `triple_mask`
    (batch_sampler.py:299-363)."""

    def __init__(self, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        length = len(self)
        axes = (0, 1) if self.triple_partition_mode == "ht_shardpair" else (0,)
        pos = np.expand_dims(np.arange(length), axis=axes)
        self.triple_mask = pos < self.triple_counts[..., None]
        wrapped = pos % self.triple_counts[..., None] + self.triple_offsets[..., None]
        # guard for an empty last partition
        self.triple_padded_idx = np.minimum(wrapped, self.triples.shape[0] - 1)

    def sample_triples(self, idx: List[int]):
        pat = "shard ... (step triple) -> step shard ... triple"
        return dict(
            sample_idx=einops.rearrange(
                self.triple_padded_idx[..., idx], pat, step=self.batches_per_step
            ),
            triple_mask=einops.rearrange(
                self.triple_mask[..., idx], pat, step=self.batches_per_step
            ),
        )


class RandomShardedBatchSampler(ShardedBatchSampler):
    """Uniform sampling with replacement inside every partition
    (batch_sampler.py:366-409)."""

    def sample_triples(self, idx: List[int]):
        n, p, bps = self.n_shard, self.positive_per_partition, self.batches_per_step
        size = (bps, n, n, p) if self.triple_partition_mode == "ht_shardpair" else (bps, n, p)
        draw = self.rng.integers(1 << 63, size=size)
        sample_idx = np.expand_dims(self.triple_offsets, axis=(0, -1)) + draw % np.expand_dims(
            self.triple_counts, axis=(0, -1)
        )
        return dict(sample_idx=sample_idx)

    def __len__(self) -> int:
        return int(np.ceil(self.triple_counts.max() / self.partition_sample_size))

    def get_dataloader_sampler(self, shuffle: bool = True) -> torch.utils.data.Sampler:
        return torch.utils.data.BatchSampler(
            torch.utils.data.SequentialSampler(self), batch_size=1, drop_last=False
        )


class PrefetchLoader:
    """Background-thread producer with a bounded queue of (pinned) step dicts."""

    _END = object()

    def __init__(
        self, sampler: ShardedBatchSampler, shuffle: bool, buffer_size: int, pin: bool
    ) -> None:
        self.sampler = sampler
        self.shuffle = shuffle
        self.buffer_size = max(1, buffer_size)
        self.pin = pin

    def __len__(self) -> int:
        return len(self.sampler.get_dataloader_sampler(self.shuffle))

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        q: "queue.Queue" = queue.Queue(maxsize=self.buffer_size)
        stop = threading.Event()

        def produce() -> None:
            try:
                for idx in self.sampler.get_dataloader_sampler(self.shuffle):
                    if stop.is_set():
                        return
                    item = self.sampler[idx]
                    if self.pin:
                        item = {k: v.pin_memory() for k, v in item.items()}
                    q.put(item)
                q.put(self._END)
            except BaseException as e:  # surface producer errors to the consumer
                q.put(e)

        t = threading.Thread(target=produce, daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is self._END:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()
            while t.is_alive():
                try:
                    q.get_nowait()
                except queue.Empty:
                    t.join(timeout=0.01)
