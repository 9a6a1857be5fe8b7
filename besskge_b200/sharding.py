"""Entity sharding and triple partitioning for BESS (host side, numpy).

Drop-in for reference `besskge/sharding.py`.  Everything here must be
BIT-EXACT with the reference: the same numpy `Generator` calls in the same
order (`sharding.py:90-102`) and the same default (unstable) `np.argsort` for
the partition order (`sharding.py:257`).  There is no device code here: these
run once at set-up time and their outputs (int arrays) feed the samplers.
"""
from __future__ import annotations

import dataclasses
import warnings
from pathlib import Path
from typing import Optional, Tuple

import numpy as np
from numpy.typing import NDArray

from .dataset import KGDataset


@dataclasses.dataclass
class Sharding:
    """Balanced random entity -> (shard, local row) map and its inverse."""

    n_shard: int
    entity_to_shard: NDArray[np.int64]  # [n_entity]
    entity_to_idx: NDArray[np.int64]  # [n_entity]
    shard_and_idx_to_entity: NDArray[np.int64]  # [n_shard, max_entity_per_shard]
    shard_counts: NDArray[np.int64]  # [n_shard] real (non padding) rows
    entity_type_counts: Optional[NDArray[np.int64]]  # [n_shard, n_type]
    entity_type_offsets: Optional[NDArray[np.int64]]  # [n_shard, n_type]

    @property
    def n_entity(self) -> int:
        return len(self.entity_to_shard)

    @property
    def max_entity_per_shard(self) -> int:
        return int(self.shard_and_idx_to_entity.shape[1])

    @classmethod
    def create(
        cls,
        n_entity: int,
        n_shard: int,
        seed: int,
        type_offsets: Optional[NDArray[np.int64]] = None,
    ) -> "Sharding":
        """reference: sharding.py:67-137."""
        rows = int(np.ceil(n_entity / n_shard))
        perm = np.random.default_rng(seed).permutation(n_shard * rows)
        # ids are kept ascending inside a shard so that types stay clustered
        shard_rows = np.sort(perm.reshape(n_shard, rows), axis=1)
        inverse = np.argsort(shard_rows.flatten())[:n_entity]
        ent_shard, ent_idx = np.divmod(inverse, rows)
        # padding ids (>= n_entity) can only sit in the last n_shard columns
        n_pad = np.sum(shard_rows[:, -n_shard:] >= n_entity, axis=-1)
        counts = rows - n_pad

        type_counts = type_off = None
        if type_offsets is not None:
            n_type = len(type_offsets)
            tid = (
                np.digitize(shard_rows, bins=type_offsets)
                + n_type * np.arange(n_shard)[:, None]
                - 1
            )
            type_counts = np.bincount(
                tid.flatten(), minlength=n_type * n_shard
            ).reshape(n_shard, -1)
            type_off = np.c_[[0] * n_shard, np.cumsum(type_counts, axis=1)[:, :-1]]
            type_counts[:, -1] -= n_pad
        return cls(
            n_shard=n_shard,
            entity_to_shard=ent_shard,
            entity_to_idx=ent_idx,
            shard_and_idx_to_entity=shard_rows,
            shard_counts=counts,
            entity_type_counts=type_counts,
            entity_type_offsets=type_off,
        )

    def save(self, out_file: Path) -> None:
        """.npz checkpoint (reference: sharding.py:139-147)."""
        fields = {k: v for k, v in dataclasses.asdict(self).items() if v is not None}
        np.savez(out_file, **fields)  # None fields are simply absent: no object arrays

    @classmethod
    def load(cls, path: Path) -> "Sharding":
        """Reference: sharding.py:149-160.  Never unpickles: the optional per-type fields are
        None when absent (or when a file written by the reference stored None as an object
        array, which numpy refuses to load without pickle)."""
        data = {}
        with np.load(path, allow_pickle=False) as f:
            for k in f.files:
                try:
                    data[k] = f[k]
                except ValueError:
                    if k not in ("entity_type_counts", "entity_type_offsets"):
                        raise
                    data[k] = None
        data.setdefault("entity_type_counts", None)
        data.setdefault("entity_type_offsets", None)
        return cls(n_shard=int(data.pop("n_shard")), **data)


def _partition(
    triples: NDArray[np.int32], sharding: Sharding, mode: str
) -> Tuple[NDArray[np.int32], NDArray[np.int64], NDArray[np.int64], NDArray[np.int64]]:
    """Bucket triples by head shard / tail shard / (head, tail) shard pair and
    rewrite the bucketed entity columns to local rows (sharding.py:226-265)."""
    n = sharding.n_shard
    if mode == "ht_shardpair":
        sh, st = sharding.entity_to_shard[triples[:, [0, 2]].T]
        pid = sh * n + st
        counts = np.bincount(pid, minlength=n * n).reshape(n, n)
        offsets = np.concatenate([np.array([0]), np.cumsum(counts)[:-1]]).reshape(n, n)
    elif mode in ("h_shard", "t_shard"):
        col = 0 if mode == "h_shard" else -1
        pid = sharding.entity_to_shard[triples[:, col]]
        counts = np.bincount(pid, minlength=n)
        offsets = np.concatenate([np.array([0]), np.cumsum(counts)[:-1]])
    else:
        raise ValueError(
            f"Partition mode {mode} not supported for triple partitioning"
        )
    order = np.argsort(pid)  # numpy default sort kind, as the reference
    out = triples[order]
    if mode != "t_shard":
        out[:, 0] = sharding.entity_to_idx[out[:, 0]]
    if mode != "h_shard":
        out[:, -1] = sharding.entity_to_idx[out[:, -1]]
    return out, counts, offsets, order


@dataclasses.dataclass
class PartitionedTripleSet:
    """Triples ordered by partition, with per-partition counts/offsets."""

    sharding: Sharding
    inverse_triples: bool
    partition_mode: str  # "h_shard" | "t_shard" | "ht_shardpair"
    dummy: Optional[str]  # "head" | "tail" | "none" | None
    triples: NDArray[np.int32]
    triple_counts: NDArray[np.int64]
    triple_offsets: NDArray[np.int64]
    triple_sort_idx: NDArray[np.int64]
    types: Optional[NDArray[np.int32]]
    neg_heads: Optional[NDArray[np.int32]]
    neg_tails: Optional[NDArray[np.int32]]

    partition_triples = staticmethod(_partition)

    @classmethod
    def create_from_dataset(
        cls,
        dataset: KGDataset,
        part: str,
        sharding: Sharding,
        partition_mode: str = "ht_shardpair",
        add_inverse_triples: bool = False,
    ) -> "PartitionedTripleSet":
        """reference: sharding.py:267-376."""
        trip = dataset.triples[part]
        n_orig = trip.shape[0]
        if add_inverse_triples:
            inv = np.copy(trip[:, ::-1])
            inv[:, 1] += dataset.n_relation_type
            trip = np.concatenate([trip, inv], axis=0)
        sorted_triples, counts, offsets, order = _partition(trip, sharding, partition_mode)

        types = None
        ht_types = dataset.ht_types
        if ht_types and part in ht_types:
            types = ht_types[part]
            if add_inverse_triples:
                types = np.concatenate([types, types[:, ::-1]], axis=0)
            types = types[order]

        has_nh = bool(dataset.neg_heads) and part in dataset.neg_heads
        has_nt = bool(dataset.neg_tails) and part in dataset.neg_tails
        nh = nt = None
        if add_inverse_triples and has_nh != has_nt:
            raise ValueError(
                "To use inverse triples, either both or neither of negative heads"
                f" and tails need to be defined for the {part} part of the dataset"
            )
        if add_inverse_triples and has_nh:
            width = dataset.neg_heads[part].shape[-1]
            bh = np.broadcast_to(dataset.neg_heads[part], (n_orig, width))
            bt = np.broadcast_to(dataset.neg_tails[part], (n_orig, width))
            nh = np.concatenate([bh, bt], axis=0)
            nt = np.concatenate([bt, bh], axis=0)
        else:
            if has_nh:
                nh = dataset.neg_heads[part]
                nh = nh.reshape(-1, nh.shape[-1])
            if has_nt:
                nt = dataset.neg_tails[part]
                nt = nt.reshape(-1, nt.shape[-1])
        if nh is not None and nh.shape[0] != 1:
            nh = nh[order]
        if nt is not None and nt.shape[0] != 1:
            nt = nt[order]
        return cls(
            sharding=sharding,
            inverse_triples=add_inverse_triples,
            partition_mode=partition_mode,
            dummy="none",
            triples=sorted_triples,
            triple_counts=counts,
            triple_offsets=offsets,
            triple_sort_idx=order,
            types=types,
            neg_heads=nh,
            neg_tails=nt,
        )

    @classmethod
    def create_from_queries(
        cls,
        dataset: KGDataset,
        sharding: Sharding,
        queries: NDArray[np.int32],
        query_mode: str,
        ground_truth: Optional[NDArray[np.int32]] = None,
        negative: Optional[NDArray[np.int32]] = None,
        negative_type: Optional[str] = None,
    ) -> "PartitionedTripleSet":
        """(h,r,?) / (?,r,t) queries completed to triples with a dummy (or the
        ground-truth) entity, partitioned by the known entity's shard
        (reference: sharding.py:378-511)."""
        n_query = queries.shape[0]
        lo = hi = 0
        if negative_type:
            if not dataset.type_offsets or negative_type not in dataset.type_offsets:
                raise ValueError(
                    f"{negative_type} is not the label of a type of entity in the KGDataset"
                )
            starts = list(dataset.type_offsets.values())
            ends = starts[1:] + [dataset.n_entity]
            bounds = {
                k: (a, b - 1) for k, a, b in zip(dataset.type_offsets, starts, ends)
            }
            lo, hi = bounds[negative_type]
            if negative is not None and (np.any(negative < lo) or np.any(negative >= hi)):
                warnings.warn(
                    "The negative entities provided are not all of the specified negative_type"
                )

        if ground_truth is not None:
            fill = ground_truth.reshape(n_query, 1)
        else:
            fill = np.full(fill_value=lo if negative_type else 0, shape=(n_query, 1))

        if negative is not None:
            negative = negative.reshape(-1, negative.shape[-1])
        elif negative_type:
            negative = np.expand_dims(np.arange(lo, hi), axis=0)
        else:
            negative = np.expand_dims(np.arange(sharding.n_entity), axis=0)

        if query_mode == "hr":
            trip = np.concatenate([queries, fill], axis=-1)
            mode, dummy = "h_shard", ("tail" if ground_truth is None else None)
            nh, nt = None, negative
        elif query_mode == "rt":
            trip = np.concatenate([fill, queries], axis=-1)
            mode, dummy = "t_shard", ("head" if ground_truth is None else None)
            nh, nt = negative, None
        else:
            raise ValueError(f"Query mode {query_mode} not supported")

        sorted_triples, counts, offsets, order = _partition(trip, sharding, mode)
        types = None
        if negative_type:
            bins = np.fromiter(dataset.type_offsets.values(), dtype=np.int32)
            types = np.digitize(sorted_triples[:, [0, 2]], bins) - 1
        if nh is not None and nh.shape[0] != 1:
            nh = nh[order]
        if nt is not None and nt.shape[0] != 1:
            nt = nt[order]
        return cls(
            sharding=sharding,
            inverse_triples=False,
            partition_mode=mode,
            dummy=dummy,
            triples=sorted_triples,
            triple_counts=counts,
            triple_offsets=offsets,
            triple_sort_idx=order,
            types=types,
            neg_heads=nh,
            neg_tails=nt,
        )
