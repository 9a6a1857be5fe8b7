"""Minimal knowledge-graph container for the BESS hot path.

Only the dataclass fields that the sharding / sampling path reads are kept
(reference: besskge/dataset.py:23-81).  Dataset downloaders and pandas
ingestion are out of scope: benchmarks use synthetic graphs of
the named dataset shapes, see `synthetic_kg`.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional

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
