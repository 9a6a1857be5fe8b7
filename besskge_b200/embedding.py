"""Embedding-table construction helpers — drop-in for reference
`besskge/embedding.py`.

Table layout contract (embedding.py:107-190): the entity table is one fp32
tensor `[n_shard, max_entity_per_shard, row]`; slice r is the shard that lives
in GPU r's HBM (row-major, row pitch = row * sizeof(dtype), which must be a
multiple of 16 B for the 128-bit gather path).  Padding rows (global id >=
n_entity) are initialised like real rows and never sampled.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Union

import numpy as np
import torch

from .sharding import Sharding

Initializer = Union[torch.Tensor, List[Callable[..., torch.Tensor]]]


def init_uniform_norm(embedding_table: torch.Tensor) -> torch.Tensor:
    """U(0,1) entries, rows scaled to unit L2 norm (embedding.py:15-26)."""
    return torch.nn.functional.normalize(torch.nn.init.uniform_(embedding_table), dim=-1)


def init_xavier_norm(embedding_table: torch.Tensor, gain: float = 1.0) -> torch.Tensor:
    """N(0, gain*sqrt(2/row)) (embedding.py:29-44)."""
    return torch.nn.init.normal_(
        embedding_table, std=gain * np.sqrt(2.0 / embedding_table.shape[-1])
    )


def init_uniform_rotation(embedding_table: torch.Tensor) -> torch.Tensor:
    """Phases in [0, 2pi) for RotatE relations (embedding.py:47-60)."""
    return torch.rand_like(embedding_table) * 2 * np.pi


def init_KGE_uniform(
    embedding_table: torch.Tensor, b: float = 1.0, divide_by_embedding_size: bool = True
) -> torch.Tensor:
    """U(-b, b), b optionally divided by the row size (embedding.py:63-81)."""
    if divide_by_embedding_size:
        b /= embedding_table.shape[-1]
    return torch.nn.init.uniform_(embedding_table, -b, b)


def init_KGE_normal(
    embedding_table: torch.Tensor, std: float = 1.0, divide_by_embedding_size: bool = True
) -> torch.Tensor:
    """N(0, std), std optionally divided by the row size (embedding.py:84-104)."""
    if divide_by_embedding_size:
        std /= embedding_table.shape[-1]
    return torch.nn.init.normal_(embedding_table, std=std)


def _from_initializers(
    lead_shape: tuple, initializer: List[Callable[..., torch.Tensor]], row_size: Optional[List[int]]
) -> torch.Tensor:
    if not row_size:
        raise ValueError("If not providing an embedding table, row_size needs to be specified")
    if len(initializer) != len(row_size):
        raise ValueError("Different number of embedding splits and initializers provided")
    # one init call per column block, in order, so the torch RNG stream matches
    # the reference (embedding.py:174-188)
    table = torch.empty((*lead_shape, 0), dtype=torch.float32)
    for width, init in zip(row_size, initializer):
        block = init(torch.empty(size=(*lead_shape, width), dtype=torch.float32))
        table = torch.concat([table, block], dim=-1)
    return table


def initialize_entity_embedding(
    sharding: Sharding, initializer: Initializer, row_size: Optional[List[int]] = None
) -> torch.nn.Parameter:
    """reference: embedding.py:107-190."""
    if isinstance(initializer, torch.Tensor):
        if initializer.dim() == 3:
            if initializer.size()[:2] != torch.Size(
                [sharding.n_shard, sharding.max_entity_per_shard]
            ):
                raise ValueError(
                    "Shape of sharded table provided for initialization"
                    " is not compatible with sharding"
                )
            table = initializer.to(torch.float32)
        elif initializer.dim() == 2:
            if initializer.shape[0] != sharding.n_entity:
                raise ValueError(
                    "Number of rows of table provided for initialization"
                    " different from number of entities."
                )
            ids = np.minimum(sharding.shard_and_idx_to_entity, sharding.n_entity - 1)
            table = initializer[torch.from_numpy(ids)].to(torch.float32)
        else:
            raise ValueError("Table for initialization needs to be 2- or 3-dimensional")
        if row_size:
            assert (
                sum(row_size) == table.shape[-1]
            ), "Initialization tensor and row_size provided are incompatible"
    else:
        table = _from_initializers(
            (sharding.n_shard, sharding.max_entity_per_shard), initializer, row_size
        )
    return torch.nn.Parameter(table)


def initialize_relation_embedding(
    n_relation_type: int,
    inverse_relations: bool,
    initializer: Initializer,
    row_size: Optional[List[int]] = None,
) -> torch.nn.Parameter:
    """reference: embedding.py:193-259."""
    if isinstance(initializer, torch.Tensor):
        if initializer.dim() != 2:
            raise ValueError("Table for initialization needs to be 2-dimensional")
        table = initializer.to(torch.float32)
        if row_size:
            assert (
                sum(row_size) == table.shape[-1]
            ), "Initialization tensor and row_size provided are incompatible"
    else:
        n_rows = 2 * n_relation_type if inverse_relations else n_relation_type
        table = _from_initializers((n_rows,), initializer, row_size)
    return torch.nn.Parameter(table)


def refactor_embedding_sharding(
    entity_embedding: torch.nn.Parameter, old_sharding: Sharding, new_sharding: Sharding
) -> torch.nn.Parameter:
    """Re-shard a trained table (embedding.py:262-290)."""
    flat = entity_embedding.detach()[
        torch.from_numpy(old_sharding.entity_to_shard),
        torch.from_numpy(old_sharding.entity_to_idx),
    ]
    return initialize_entity_embedding(initializer=flat.cpu(), sharding=new_sharding)
