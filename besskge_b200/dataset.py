"""Minimal knowledge-graph container for the BESS hot path.

The dataclass (reference: besskge/dataset.py:23-81) with its in-memory constructors
`from_triples` / `from_dataframe` (dataset.py:83-239) and `save` / `load`.  The
`build_*` downloaders of the reference (dataset.py:241-460: OGB, YAGO3-10, OpenBioLink)
need network access and the `ogb` package and stay out of scope: a
downloaded dataset enters through `from_triples` / `from_dataframe`, benchmarks use
synthetic graphs of the named dataset shapes (`synthetic_kg`).
"""
from __future__ import annotations

import dataclasses
import json
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
from numpy.typing import NDArray


@dataclasses.dataclass
class KGDataset:
    n_entity: int
    n_relation_type: int
    #: {part: int[n_triple, 3]} columns are (head, relation, tail) global ids
    triples: Dict[str, NDArray[np.int32]]
    original_triple_ids: Optional[Dict[str, NDArray[np.int32]]] = None
    entity_dict: Optional[List[str]] = None
    relation_dict: Optional[List[str]] = None
    #: {type label: first global id of the type}; ids are clustered by type
    type_offsets: Optional[Dict[str, int]] = None
    #: {part: int32[n_triple or 1, n_neg]}
    neg_heads: Optional[Dict[str, NDArray[np.int32]]] = None
    neg_tails: Optional[Dict[str, NDArray[np.int32]]] = None

    @property
    def ht_types(self) -> Optional[Dict[str, NDArray[np.int32]]]:
        """Type id of head and tail of every triple (dataset.py:63-81)."""
        if not self.type_offsets:
            return None
        bins = np.fromiter(self.type_offsets.values(), dtype=np.int32)
        return {
            part: np.digitize(trip[:, [0, 2]], bins) - 1
            for part, trip in self.triples.items()
        }

    # ------------------------------------------------------------ constructors --
    @classmethod
    def from_triples(
        cls,
        data: NDArray[np.int32],
        split: Tuple[float, float, float] = (0.7, 0.15, 0.15),
        seed: int = 1234,
        entity_dict: Optional[List[str]] = None,
        relation_dict: Optional[List[str]] = None,
        type_offsets: Optional[Dict[str, int]] = None,
    ) -> "KGDataset":
        """Random train / valid / test split of an id-triple array [n, 3] (reference:
        dataset.py:83-145; same permutation stream, so the parts are identical for a seed).
        `original_triple_ids[part]` = rows of `data` that went to `part`."""
        n = data.shape[0]
        cut_train = int(n * split[0])
        cut_valid = cut_train + int(n * split[1])
        order = np.random.default_rng(seed=seed).permutation(np.arange(n))
        ids = dict(zip(("train", "valid", "test"), np.split(order, (cut_train, cut_valid), axis=0)))
        return cls(
            n_entity=data[:, [0, 2]].max() + 1,
            n_relation_type=data[:, 1].max() + 1,
            triples={part: data[rows] for part, rows in ids.items()},
            original_triple_ids=ids,
            entity_dict=entity_dict,
            relation_dict=relation_dict,
            type_offsets=type_offsets,
        )

    @classmethod
    def from_dataframe(
        cls,
        df: Any,
        head_column: Union[int, str],
        relation_column: Union[int, str],
        tail_column: Union[int, str],
        entity_types: Optional[Any] = None,
        split: Tuple[float, float, float] = (0.7, 0.15, 0.15),
        seed: int = 1234,
    ) -> "KGDataset":
        """Labelled (h, r, t) triples in a pandas DataFrame — or {part: DataFrame} for a
        pre-defined split — to ids (reference: dataset.py:147-239).  Entities are numbered in
        order of first appearance (all heads, then all tails, part by part); with
        `entity_types` (label -> type) they are re-ordered so that each type is a contiguous
        id range and `type_offsets` holds the first id of every type."""
        import pandas as pd

        parts = {"all": df} if isinstance(df, pd.DataFrame) else df
        ent_labels = pd.concat(
            [pd.concat([d[head_column], d[tail_column]]) for d in parts.values()]).unique()
        rel_labels = pd.concat([d[relation_column] for d in parts.values()]).unique()
        ent2id = pd.Series(np.arange(len(ent_labels)), index=ent_labels, name="ent_id")
        rel2id = pd.Series(np.arange(len(rel_labels)), index=rel_labels, name="rel_id")
        type_offsets = None
        if entity_types is not None:
            # same pandas calls as the reference: the tie order of sort_values decides the ids
            typed = pd.merge(ent2id, pd.Series(entity_types, name="ent_type"), how="left",
                             left_index=True, right_index=True).sort_values("ent_type")
            ent2id.index = typed.index
            first = typed.groupby("ent_type")["ent_type"].count().cumsum().shift(1)
            first.iloc[0] = 0
            type_offsets = first.astype("int64").to_dict()
        entity_dict, relation_dict = ent2id.index.tolist(), rel2id.index.tolist()
        triples = {
            part: np.stack([d[head_column].map(ent2id).values.astype(np.int32),
                            d[relation_column].map(rel2id).values.astype(np.int32),
                            d[tail_column].map(ent2id).values.astype(np.int32)], axis=1)
            for part, d in parts.items()
        }
        if isinstance(df, pd.DataFrame):
            return cls.from_triples(triples["all"], split, seed, entity_dict, relation_dict,
                                    type_offsets)
        return cls(
            n_entity=len(entity_dict),
            n_relation_type=len(relation_dict),
            triples=triples,
            original_triple_ids={k: np.arange(v.shape[0]) for k, v in triples.items()},
            entity_dict=entity_dict,
            relation_dict=relation_dict,
            type_offsets=type_offsets,
        )

    # ------------------------------------------------------------- save / load --
    _ARRAY_FIELDS = ("triples", "original_triple_ids", "neg_heads", "neg_tails")

    def save(self, out_file: Path) -> None:
        """One `.npz` file: arrays as `<field>/<part>`, everything else as JSON.  (The
        reference pickles the object, dataset.py:462-472; a pickle is code execution on
        load, so this build uses a data-only container.)"""
        arrays: Dict[str, Any] = {}
        for field in self._ARRAY_FIELDS:
            for part, arr in (getattr(self, field) or {}).items():
                arrays[f"{field}/{part}"] = np.asarray(arr)
        meta = dict(n_entity=int(self.n_entity), n_relation_type=int(self.n_relation_type),
                    entity_dict=self.entity_dict, relation_dict=self.relation_dict,
                    type_offsets=None if self.type_offsets is None else
                    {str(k): int(v) for k, v in self.type_offsets.items()},
                    present=[f for f in self._ARRAY_FIELDS if getattr(self, f) is not None])
        with open(out_file, "wb") as f:
            np.savez(f, __meta__=np.array(json.dumps(meta)), **arrays)

    @classmethod
    def load(cls, path: Path) -> "KGDataset":
        with np.load(path, allow_pickle=False) as data:
            if "__meta__" not in data.files:
                raise ValueError(f"File at path {path} is not a KGDataset")
            meta = json.loads(str(data["__meta__"]))
            fields: Dict[str, Any] = {f: ({} if f in meta["present"] else None)
                                      for f in cls._ARRAY_FIELDS}
            for key in data.files:
                if "/" in key:
                    field, part = key.split("/", 1)
                    fields[field][part] = data[key]
        return cls(n_entity=meta["n_entity"], n_relation_type=meta["n_relation_type"],
                   entity_dict=meta["entity_dict"], relation_dict=meta["relation_dict"],
                   type_offsets=meta["type_offsets"], **fields)


#: (n_entity, n_relation_type, n_train_triple) of the datasets named in BASELINE.json
DATASET_SHAPES = {
    "ogbl-biokg": (93_773, 51, 4_762_678),
    "yago3-10": (123_182, 37, 1_079_040),
    "ogbl-wikikg2": (2_500_604, 535, 16_109_182),
}


def synthetic_kg(
    shape: str, seed: int = 1234, n_triple: Optional[int] = None, part: str = "train"
) -> KGDataset:
    """Uniform-random graph with the entity/relation/triple counts of `shape`
    (h,t ~ U[0,E), r ~ U[0,R), int32)."""
    n_entity, n_rel, n_trip = DATASET_SHAPES[shape]
    if n_triple is not None:
        n_trip = n_triple
    rng = np.random.default_rng(seed)
    h = rng.integers(n_entity, size=n_trip, dtype=np.int32)
    r = rng.integers(n_rel, size=n_trip, dtype=np.int32)
    t = rng.integers(n_entity, size=n_trip, dtype=np.int32)
    return KGDataset(
        n_entity=n_entity,
        n_relation_type=n_rel,
        triples={part: np.stack([h, r, t], axis=1)},
    )
