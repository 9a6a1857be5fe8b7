"""Sharded negative samplers (host side, numpy) — drop-in for reference
`besskge/negative_sampler.py`.

Output contract (reference negative_sampler.py:33-54): `negative_entities`
int [bps, n_src_shard, n_dst_shard, B, n_negative] holds LOCAL rows of the
source shard; B = 1 (flat "h"/"t"), 2 (flat "ht") or shard_bs (per triple).
The numpy `Generator` (PCG64) stream is part of the contract: indices must be
bit-exact with the reference, so sampling stays on the host and the draws are
made with the very same calls.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Dict, Optional, Tuple, Union

import einops
import numpy as np
from numpy.typing import NDArray

from .sharding import Sharding

SampleDict = Dict[str, Union[NDArray[np.int32], NDArray[np.bool_]]]


class ShardedNegativeSampler(ABC):
    flat_negative_format: bool
    local_sampling: bool
    corruption_scheme: str  # "h" | "t" | "ht"
    rng: np.random.Generator

    @abstractmethod
    def __call__(self, sample_idx: NDArray[np.int64]) -> SampleDict:
        """sample_idx: [bps, n_shard, (n_shard,) triples_per_partition]."""


class RandomShardedNegativeSampler(ShardedNegativeSampler):
    """Uniform negatives from each source shard (negative_sampler.py:57-132)."""

    def __init__(
        self,
        n_negative: int,
        sharding: Sharding,
        seed: int,
        corruption_scheme: str,
        local_sampling: bool,
        flat_negative_format: bool = False,
    ) -> None:
        self.n_negative = n_negative
        self.sharding = sharding
        self.shard_counts = sharding.shard_counts
        self.corruption_scheme = corruption_scheme
        self.local_sampling = local_sampling
        self.flat_negative_format = flat_negative_format
        self.seed = seed
        self.rng = np.random.default_rng(seed=seed)

    def _draw(self, sample_idx: NDArray[np.int64]) -> NDArray[np.int64]:
        bps, n = sample_idx.shape[:2]
        per_part = sample_idx.shape[-1]
        if self.flat_negative_format:
            b = 2 if self.corruption_scheme == "ht" else 1
        else:
            b = per_part if sample_idx.ndim == 3 else n * per_part
        raw = self.rng.integers(1 << 31, size=(bps, n, n, b, self.n_negative))
        # axis 1 is the SOURCE shard: rows are valid local ids of that shard
        return raw.astype(np.int32) % self.shard_counts[None, :, None, None, None]

    def __call__(self, sample_idx: NDArray[np.int64]) -> SampleDict:
        return dict(negative_entities=self._draw(sample_idx))


class TypeBasedShardedNegativeSampler(RandomShardedNegativeSampler):
    """Negatives of the same entity type as the corrupted entity
    (negative_sampler.py:135-230)."""

    def __init__(
        self,
        triple_types: NDArray[np.int32],
        n_negative: int,
        sharding: Sharding,
        corruption_scheme: str,
        local_sampling: bool,
        seed: int,
    ) -> None:
        super().__init__(
            n_negative, sharding, seed, corruption_scheme, local_sampling, False
        )
        if sharding.entity_type_counts is None or sharding.entity_type_offsets is None:
            raise ValueError("The provided entity sharding does not have entity types")
        self.triple_types = triple_types
        self.type_counts = sharding.entity_type_counts
        self.type_offsets = sharding.entity_type_offsets

    def __call__(self, sample_idx: NDArray[np.int64]) -> SampleDict:
        n = sample_idx.shape[1]
        per_part = sample_idx.shape[-1]
        h_type, t_type = einops.rearrange(
            np.take(self.triple_types, sample_idx, axis=0), "... ht -> ht ..."
        )
        if self.corruption_scheme == "h":
            wanted = h_type
        elif self.corruption_scheme == "t":
            wanted = t_type
        elif self.corruption_scheme == "ht":
            cut = per_part // 2
            wanted = np.concatenate([h_type[..., :cut], t_type[..., cut:]], axis=-1)
        else:
            raise ValueError(
                f"Corruption scheme {self.corruption_scheme} not supported by {type(self)}"
            )
        # broadcast the wanted type over the source-shard axis
        pattern = (
            "step shard ... triple -> step shard r (... triple)"
            if self.local_sampling
            else "step shard ... triple -> step r shard (... triple)"
        )
        wanted = einops.repeat(wanted, pattern, r=n)
        raw = self._draw(sample_idx)
        src = np.arange(n)[None, :, None, None]
        # a shard without entities of the wanted type gives x % 0: numpy defines it as 0 (the
        # reference computes the same value, with a RuntimeWarning that carries no information)
        with np.errstate(divide="ignore"):
            rows = (
                raw % self.type_counts[src, wanted, np.newaxis]
                + self.type_offsets[src, wanted, np.newaxis]
            )
        return dict(negative_entities=rows)


class TripleBasedShardedNegativeSampler(ShardedNegativeSampler):
    """Predetermined (possibly per-triple) candidate lists, split by owning
    shard and padded to a common per-shard length
    (negative_sampler.py:233-540)."""

    _ENT_PER_TRIPLE = (
        "step shard ... triple shard_neg idx_neg -> step shard_neg shard (... triple) idx_neg"
    )
    _ENT_FLAT = "pad shard_neg idx_neg -> step shard_neg shard pad idx_neg"
    _MASK_PER_TRIPLE = (
        "step shard ... triple shard_neg idx_neg -> step shard (... triple) shard_neg idx_neg"
    )
    _MASK_FLAT = "pad shard_neg idx_neg -> step shard pad shard_neg idx_neg"

    def __init__(
        self,
        negative_heads: Optional[NDArray[np.int32]],
        negative_tails: Optional[NDArray[np.int32]],
        sharding: Sharding,
        corruption_scheme: str,
        seed: int,
        mask_on_gather: bool = False,
        return_sort_idx: bool = False,
    ):
        if negative_heads is not None and negative_tails is not None:
            assert (
                negative_heads.shape == negative_tails.shape
            ), "negative_heads and negative_tails need to have the same size"
            negative_heads = negative_heads.reshape(-1, negative_heads.shape[-1])
            negative_tails = negative_tails.reshape(-1, negative_tails.shape[-1])
            self.N, self.n_negative = negative_heads.shape
        elif negative_tails is not None:
            assert corruption_scheme == "t", (
                f"Corruption scheme '{corruption_scheme}' requires providing negative_heads"
            )
            negative_tails = negative_tails.reshape(-1, negative_tails.shape[-1])
            self.N, self.n_negative = negative_tails.shape
        elif negative_heads is not None:
            assert corruption_scheme == "h", (
                f"Corruption scheme '{corruption_scheme}' requires providing negative_tails"
            )
            negative_heads = negative_heads.reshape(-1, negative_heads.shape[-1])
            self.N, self.n_negative = negative_heads.shape
        else:
            raise ValueError(
                "At least one of negative_heads and negative_tails needs to be provided"
            )
        self.sharding = sharding
        self.shard_counts = sharding.shard_counts
        self.corruption_scheme = corruption_scheme
        self.local_sampling = False
        self.flat_negative_format = self.N == 1
        self.return_sort_idx = return_sort_idx
        self.mask_on_gather = mask_on_gather
        self.rng = np.random.default_rng(seed=seed)

        def prepare(neg: NDArray[np.int32]):
            counts, order = self.shard_negatives(neg)
            return neg, counts, order

        if corruption_scheme in ("h", "t"):
            neg, counts, self.sort_neg_idx = prepare(
                negative_heads if corruption_scheme == "h" else negative_tails
            )
            self.padded_shard_length = counts.max()
            self.padded_negatives, self.mask = self.pad_negatives(
                sharding.entity_to_idx[np.take_along_axis(neg, self.sort_neg_idx, axis=-1)],
                counts,
                self.padded_shard_length,
            )
        elif corruption_scheme == "ht":
            nh, ch, self.sort_neg_h_idx = prepare(negative_heads)
            nt, ct, self.sort_neg_t_idx = prepare(negative_tails)
            self.padded_shard_length = np.max([ch.max(), ct.max()])
            self.padded_negatives_h, self.mask_h = self.pad_negatives(
                sharding.entity_to_idx[np.take_along_axis(nh, self.sort_neg_h_idx, axis=-1)],
                ch,
                self.padded_shard_length,
            )
            self.padded_negatives_t, self.mask_t = self.pad_negatives(
                sharding.entity_to_idx[np.take_along_axis(nt, self.sort_neg_t_idx, axis=-1)],
                ct,
                self.padded_shard_length,
            )
        else:
            raise ValueError(
                f"Corruption scheme {corruption_scheme} not supported by {type(self)}"
            )
        # entities are consumed on the shard that stores them (shard_neg); masks
        # on that shard too if mask_on_gather (TopK), else on the scoring shard
        self.ent_rearrange_pattern = self._ENT_PER_TRIPLE
        self.ent_repeat_pattern = self._ENT_FLAT
        if mask_on_gather:
            self.mask_rearrange_pattern = self._ENT_PER_TRIPLE
            self.mask_repeat_pattern = self._ENT_FLAT
        else:
            self.mask_rearrange_pattern = self._MASK_PER_TRIPLE
            self.mask_repeat_pattern = self._MASK_FLAT

    def __call__(self, sample_idx: NDArray[np.int64]) -> SampleDict:
        sort_idx = None
        if self.corruption_scheme in ("h", "t"):
            orig_shape = sample_idx.shape
            if self.flat_negative_format:
                sample_idx = np.full(fill_value=0, shape=(*sample_idx.shape[:2], 1))
            ents = einops.rearrange(
                np.take(self.padded_negatives, sample_idx, axis=0), self.ent_rearrange_pattern
            )
            mask = einops.rearrange(
                np.take(self.mask, sample_idx, axis=0), self.mask_rearrange_pattern
            )
            if self.return_sort_idx:
                pick = (
                    np.full(fill_value=0, shape=orig_shape)
                    if self.flat_negative_format
                    else sample_idx
                )
                sort_idx = np.take(self.sort_neg_idx, pick, axis=0)
        else:  # "ht"
            cut = sample_idx.shape[-1] // 2
            if self.flat_negative_format:
                bps, n = sample_idx.shape[:2]
                ents = einops.repeat(
                    np.concatenate([self.padded_negatives_h, self.padded_negatives_t], axis=0),
                    self.ent_repeat_pattern,
                    step=bps,
                    shard=n,
                )
                mask = einops.repeat(
                    np.concatenate([self.mask_h, self.mask_t], axis=0),
                    self.mask_repeat_pattern,
                    step=bps,
                    shard=n,
                )
                idx_h = np.full(fill_value=0, shape=(*sample_idx.shape[:-1], cut))
                idx_t = np.full(
                    fill_value=0,
                    shape=(*sample_idx.shape[:-1], sample_idx.shape[-1] - cut),
                )
            else:
                idx_h = sample_idx[..., :cut]
                idx_t = sample_idx[..., cut:]
                ents = einops.rearrange(
                    np.concatenate(
                        [np.take(self.padded_negatives_h, idx_h, axis=0),
                         np.take(self.padded_negatives_t, idx_t, axis=0)],
                        axis=-3,
                    ),
                    self.ent_rearrange_pattern,
                )
                mask = einops.rearrange(
                    np.concatenate([np.take(self.mask_h, idx_h, axis=0),
                                    np.take(self.mask_t, idx_t, axis=0)], axis=-3),
                    self.mask_rearrange_pattern,
                )
            if self.return_sort_idx:
                sort_idx = np.concatenate(
                    [np.take(self.sort_neg_h_idx, idx_h, axis=0),
                     np.take(self.sort_neg_t_idx, idx_t, axis=0)], axis=-2
                )
        out: SampleDict = dict(negative_entities=ents, negative_mask=mask)
        if self.return_sort_idx:
            out["negative_sort_idx"] = einops.rearrange(
                sort_idx,
                "step shard ... triple idx_neg -> step shard (... triple) idx_neg",
            )
        return out

    def shard_negatives(
        self, negatives: NDArray[np.int32]
    ) -> Tuple[NDArray[np.int64], NDArray[np.int32]]:
        """Per-row count of negatives owned by each shard and the argsort that
        clusters each row by owning shard (negative_sampler.py:479-501)."""
        n = self.sharding.n_shard
        owner = self.sharding.entity_to_shard[negatives]
        counts = np.bincount(
            (owner + n * np.arange(self.N)[:, None]).flatten(), minlength=n * self.N
        ).reshape(self.N, n)
        order = np.argsort(owner, axis=-1)
        return counts, order.astype(np.int32)

    def pad_negatives(
        self,
        negatives: NDArray[np.int32],
        shard_counts: NDArray[np.int64],
        padded_shard_length: int,
    ) -> Tuple[NDArray[np.int32], NDArray[np.bool_]]:
        """[N, n_neg] rows clustered by shard -> [N, n_shard, padded_len] with
        wrap-around padding, and the validity mask (negative_sampler.py:503-540)."""
        slot = np.arange(padded_shard_length)[None, None, :]
        mask = slot < shard_counts[..., None]
        starts = np.c_[[0] * self.N, np.cumsum(shard_counts, axis=-1)[:, :-1]]
        with np.errstate(divide="ignore"):  # a shard that holds no candidate: x % 0 := 0, masked
            src = np.minimum(
                slot % shard_counts[..., None] + starts[..., None], self.n_negative - 1
            )
        return negatives[np.arange(self.N)[:, None, None], src], mask


class PlaceholderNegativeSampler(ShardedNegativeSampler):
    """No negatives: score against all entities (negative_sampler.py:543-574)."""

    def __init__(self, corruption_scheme: str, seed: int = 0) -> None:
        self.corruption_scheme = corruption_scheme
        self.local_sampling = False
        self.flat_negative_format = True
        self.seed = seed
        self.rng = np.random.default_rng(seed=seed)

    def __call__(self, sample_idx: NDArray[np.int64]) -> SampleDict:
        return dict()
