"""Score-function classes — drop-in for reference `besskge/scoring.py`
(TransE, RotatE, DistMult, ComplEx, PairRE, BoxE), executing on hand-written
sm_100a kernels (csrc/rows.cu, csrc/pair.cu) through the C-ABI.

Same constructor signatures, same attributes (`entity_embedding`
[n_shard, max_entity_per_shard, row], `relation_embedding`, `sharding`,
`negative_sample_sharing` — mutable at run time) and the same three methods
`score_triple / score_heads / score_tails` (scoring.py:45-113).  The methods
take CUDA tensors and raise on CPU tensors: there is no CPU fallback.  They are
forward-only; training goes through `besskge_b200.bess.training_model`, whose
fused step never materialises the dense table gradient.
"""
from __future__ import annotations

from abc import ABC
from typing import Callable, List, Optional, Union

import torch

from . import _lib as L
from . import kernels as K
from .embedding import (
    init_KGE_normal,
    init_KGE_uniform,
    init_uniform_norm,
    init_uniform_rotation,
    init_xavier_norm,  # noqa: F401  (re-exported like the reference)
    initialize_entity_embedding,
    initialize_relation_embedding,
    refactor_embedding_sharding,
)
from .sharding import Sharding

Initializer = Union[torch.Tensor, List[Callable[..., torch.Tensor]]]


class BaseScoreFunction(torch.nn.Module, ABC):
    """Common machinery: kernel configuration + the three scoring methods."""

    negative_sample_sharing: bool
    sharding: Sharding
    entity_embedding: torch.nn.Parameter
    relation_embedding: torch.nn.Parameter
    _family: int = -1

    # -- kernel configuration ------------------------------------------------
    def kernel_cfg(self) -> L.ScoreCfg:
        return L.ScoreCfg(
            family=self._family,
            norm_p=int(getattr(self, "scoring_norm", 2)),
            d=int(self.embedding_size),
            normalize=int(getattr(self, "normalize", False)),
            apply_tanh=int(getattr(self, "apply_tanh", False)),
            per_dim=int(getattr(self, "dist_func_per_dim", True)),
            eps=float(getattr(self, "eps", 0.0)),
            rel_u=float(getattr(self, "u", 0.0)),
        )

    @property
    def entity_width(self) -> int:
        return int(self.entity_embedding.shape[-1])

    @property
    def relation_width(self) -> int:
        return int(self.relation_embedding.shape[-1])

    def _check_norm(self) -> None:
        p = getattr(self, "scoring_norm", None)
        if p is not None and p not in (1, 2):
            raise NotImplementedError(
                f"scoring_norm={p}: the CUDA kernels implement the 1- and 2-norm"
            )

    # -- public API (scoring.py:45-124) ---------------------------------------
    def score_triple(
        self, head_emb: torch.Tensor, relation_id: torch.Tensor, tail_emb: torch.Tensor
    ) -> torch.Tensor:
        """[b, W], [b], [b, W] -> [b] scores of (h, r, t)."""
        K.require_cuda(head_emb, relation_id, tail_emb, self.relation_embedding)
        self._check_norm()
        cfg, dt = self.kernel_cfg(), L.dtype_code(head_emb.dtype)
        h, t = head_emb.contiguous(), tail_emb.contiguous().to(head_emb.dtype)
        rel = relation_id.to(torch.int32).contiguous()
        rtab = self.relation_embedding.detach().to(head_emb.dtype).contiguous()
        n = h.shape[0]
        out = torch.empty(n, dtype=torch.float32, device=h.device)
        K.triple_fwd(cfg, dt, L.rows(h), L.rows(t), rtab, rel, L.IDENT, n, out, L.IDENT)
        return out.to(head_emb.dtype)

    def _score_candidates(
        self, mode: int, fixed: torch.Tensor, relation_id: torch.Tensor, cand: torch.Tensor
    ) -> torch.Tensor:
        K.require_cuda(fixed, relation_id, cand, self.relation_embedding)
        self._check_norm()
        cfg, dt = self.kernel_cfg(), L.dtype_code(fixed.dtype)
        W = self.entity_width
        x = fixed.contiguous()
        rel = relation_id.to(torch.int32).contiguous()
        rtab = self.relation_embedding.detach().to(fixed.dtype).contiguous()
        nq = x.shape[0]
        nvec = K.call("bess_query_nvec", L.C.byref(cfg))
        qv = torch.empty(nq, nvec, W, dtype=torch.float32, device=x.device)
        K.prologue_fwd(cfg, dt, mode, L.rows(x), rtab, rel, L.IDENT, nq, qv)
        c = cand.contiguous().to(fixed.dtype)
        if c.dim() == 2:
            c = c.unsqueeze(0)
        shared = self.negative_sample_sharing or c.shape[0] == 1
        need_aux = self._family == L.BOXE and cfg.norm_p == 2
        if shared:
            flat = c.reshape(-1, W)
            nc = flat.shape[0]
            out = torch.empty(nq, nc, dtype=torch.float32, device=x.device)
            aux = torch.empty_like(out) if need_aux else None
            scale = None
            if K.needs_cand_scale(cfg):
                scale = K.cand_scales(cfg, dt, L.rows(flat), nc, W,
                                      torch.empty(2 * nc, dtype=torch.float32, device=x.device))
            K.shared_fwd(cfg, dt, mode, qv, nq, L.rows(flat), scale, nc, out, L.IDENT, nc, 0, aux)
        else:
            if c.shape[0] != nq:
                raise ValueError(
                    "per-triple candidates need one candidate set per query "
                    f"({c.shape[0]} sets for {nq} queries)"
                )
            n_per = c.shape[1]
            flat = c.reshape(-1, W)
            out = torch.empty(nq, n_per, dtype=torch.float32, device=x.device)
            aux = torch.empty_like(out) if need_aux else None
            K.pertriple_fwd(cfg, dt, mode, qv, nq, L.rows(flat), n_per, n_per, out, L.IDENT,
                            n_per, 0, aux)
        return out.to(fixed.dtype)

    def score_heads(
        self, head_emb: torch.Tensor, relation_id: torch.Tensor, tail_emb: torch.Tensor
    ) -> torch.Tensor:
        """candidate heads [B, n_heads, W] against (r, t) queries -> [b, B*n_heads]
        with negative sample sharing, else [b, n_heads]."""
        return self._score_candidates(L.MODE_HEADS, tail_emb, relation_id, head_emb)

    def score_tails(
        self, head_emb: torch.Tensor, relation_id: torch.Tensor, tail_emb: torch.Tensor
    ) -> torch.Tensor:
        """candidate tails [B, n_tails, W] against (h, r) queries."""
        return self._score_candidates(L.MODE_TAILS, head_emb, relation_id, tail_emb)

    def forward(
        self, head_emb: torch.Tensor, relation_id: torch.Tensor, tail_emb: torch.Tensor
    ) -> torch.Tensor:
        return self.score_triple(head_emb, relation_id, tail_emb)

    def update_sharding(self, new_sharding: Sharding) -> None:
        """Re-shard the entity table (scoring.py:126-142)."""
        dev, dt = self.entity_embedding.device, self.entity_embedding.dtype
        new = refactor_embedding_sharding(
            torch.nn.Parameter(self.entity_embedding.detach().float().cpu()),
            self.sharding,
            new_sharding,
        )
        self.entity_embedding = torch.nn.Parameter(new.detach().to(device=dev, dtype=dt))
        self.sharding = new_sharding

    # -- shared constructor body ----------------------------------------------
    def _build_tables(
        self,
        sharding: Sharding,
        n_relation_type: int,
        inverse_relations: bool,
        entity_initializer: Initializer,
        relation_initializer: Initializer,
        entity_rows: List[int],
        relation_rows: List[int],
    ) -> None:
        self.sharding = sharding
        self.entity_embedding = initialize_entity_embedding(
            sharding, entity_initializer, entity_rows
        )
        self.relation_embedding = initialize_relation_embedding(
            n_relation_type, inverse_relations, relation_initializer, relation_rows
        )


class DistanceBasedScoreFunction(BaseScoreFunction, ABC):
    def __init__(self, negative_sample_sharing: bool, scoring_norm: int) -> None:
        super().__init__()
        self.negative_sample_sharing = negative_sample_sharing
        self.scoring_norm = scoring_norm


class MatrixDecompositionScoreFunction(BaseScoreFunction, ABC):
    def __init__(self, negative_sample_sharing: bool) -> None:
        super().__init__()
        self.negative_sample_sharing = negative_sample_sharing


class TransE(DistanceBasedScoreFunction):
    """-||h + r - t||_p (scoring.py:258-354)."""

    _family = L.TRANSE

    def __init__(
        self,
        negative_sample_sharing: bool,
        scoring_norm: int,
        sharding: Sharding,
        n_relation_type: int,
        embedding_size: int,
        entity_initializer: Initializer = [init_KGE_uniform],
        relation_initializer: Initializer = [init_KGE_uniform],
        inverse_relations: bool = False,
    ) -> None:
        super().__init__(negative_sample_sharing, scoring_norm)
        self._build_tables(sharding, n_relation_type, inverse_relations, entity_initializer,
                           relation_initializer, [embedding_size], [embedding_size])
        assert (
            self.entity_embedding.shape[-1] == self.relation_embedding.shape[-1] == embedding_size
        ), "TransE requires `embedding_size` embedding parameters for each entity and relation"
        self.embedding_size = embedding_size


class RotatE(DistanceBasedScoreFunction):
    """-||h o exp(i r) - t||_p over the 2d real vector (scoring.py:357-462);
    entity rows are [re | im], relation rows are phases."""

    _family = L.ROTATE

    def __init__(
        self,
        negative_sample_sharing: bool,
        scoring_norm: int,
        sharding: Sharding,
        n_relation_type: int,
        embedding_size: int,
        entity_initializer: Initializer = [init_KGE_uniform],
        relation_initializer: Initializer = [init_uniform_rotation],
        inverse_relations: bool = False,
    ) -> None:
        super().__init__(negative_sample_sharing, scoring_norm)
        self._build_tables(sharding, n_relation_type, inverse_relations, entity_initializer,
                           relation_initializer, [2 * embedding_size], [embedding_size])
        assert (
            self.entity_embedding.shape[-1]
            == 2 * self.relation_embedding.shape[-1]
            == 2 * embedding_size
        ), (
            "RotatE requires `2*embedding_size` embedding parameters for each entity"
            "and `embedding_size` embedding parameters for each relation"
        )
        self.embedding_size = embedding_size


class PairRE(DistanceBasedScoreFunction):
    """-||h^ o r_h - t^ o r_t||_p (scoring.py:465-593)."""

    _family = L.PAIRRE

    def __init__(
        self,
        negative_sample_sharing: bool,
        scoring_norm: int,
        sharding: Sharding,
        n_relation_type: int,
        embedding_size: int,
        entity_initializer: Initializer = [init_KGE_uniform],
        relation_initializer: Initializer = [init_KGE_uniform],
        normalize_entities: bool = True,
        inverse_relations: bool = False,
    ) -> None:
        super().__init__(negative_sample_sharing, scoring_norm)
        self.normalize = normalize_entities
        if isinstance(relation_initializer, list):
            relation_initializer = 2 * relation_initializer
        self._build_tables(sharding, n_relation_type, inverse_relations, entity_initializer,
                           relation_initializer, [embedding_size],
                           [embedding_size, embedding_size])
        assert (
            2 * self.entity_embedding.shape[-1]
            == self.relation_embedding.shape[-1]
            == 2 * embedding_size
        ), (
            "PairRE requires `embedding_size` embedding parameters for each entity"
            "and `2*embedding_size` embedding parameters for each relation"
        )
        self.embedding_size = embedding_size


class TripleRE(DistanceBasedScoreFunction):
    """-||h^ o (r_h + u) - t^ o (r_t + u) + r_m||_p (scoring.py:596-743); relation rows are
    [r_h | r_m | r_t].  Runs on the PairRE kernels: the query prologue folds r_m into the
    query vector, the pair function and the candidate normalisation are PairRE's."""

    _family = L.TRIPLERE

    def __init__(
        self,
        negative_sample_sharing: bool,
        scoring_norm: int,
        sharding: Sharding,
        n_relation_type: int,
        embedding_size: int,
        entity_initializer: Initializer = [init_KGE_uniform],
        relation_initializer: Initializer = [init_KGE_uniform],
        normalize_entities: bool = True,
        u: float = 0.0,
        inverse_relations: bool = False,
    ) -> None:
        super().__init__(negative_sample_sharing, scoring_norm)
        self.normalize = normalize_entities
        if isinstance(relation_initializer, list):
            relation_initializer = 3 * relation_initializer
        self._build_tables(sharding, n_relation_type, inverse_relations, entity_initializer,
                           relation_initializer, [embedding_size],
                           [embedding_size, embedding_size, embedding_size])
        assert (
            3 * self.entity_embedding.shape[-1]
            == self.relation_embedding.shape[-1]
            == 3 * embedding_size
        ), (
            "TripleRE requires `embedding_size` embedding parameters for each entity"
            "and `3*embedding_size` embedding parameters for each relation"
        )
        self.embedding_size = embedding_size
        self.use_v2 = u > 0.0
        # the reference adds rel_u only when u > 0 (scoring.py:692-694); u <= 0 means "no offset"
        self.u = float(u) if self.use_v2 else 0.0
        self.register_buffer("rel_u", torch.tensor([u], dtype=self.entity_embedding.dtype))


class _AuxEntityScoreFunction(DistanceBasedScoreFunction):
    """Shared constructor of InterHT and TranS: entity rows [main | auxiliary] of
    2 * embedding_size elements, each half L2-normalised before use (`normalize_entities`), the
    auxiliary halves shifted by `offset`.  Negative scoring runs on csrc/pair2.cu."""

    _relation_parts = 1

    def __init__(
        self,
        negative_sample_sharing: bool,
        scoring_norm: int,
        sharding: Sharding,
        n_relation_type: int,
        embedding_size: int,
        entity_initializer: Initializer = [init_KGE_uniform],
        relation_initializer: Initializer = [init_KGE_uniform],
        normalize_entities: bool = True,
        offset: float = 1.0,
        inverse_relations: bool = False,
    ) -> None:
        super().__init__(negative_sample_sharing, scoring_norm)
        self.normalize = normalize_entities
        if isinstance(entity_initializer, list):
            entity_initializer = 2 * entity_initializer
        if isinstance(relation_initializer, list):
            relation_initializer = self._relation_parts * relation_initializer
        self._build_tables(sharding, n_relation_type, inverse_relations, entity_initializer,
                           relation_initializer, [embedding_size, embedding_size],
                           self._relation_parts * [embedding_size])
        name = type(self).__name__
        assert (
            self.entity_embedding.shape[-1] == 2 * embedding_size
            and self.relation_embedding.shape[-1] == self._relation_parts * embedding_size
        ), (
            f"{name} requires `2*embedding_size` embedding parameters for each entity and "
            f"`{self._relation_parts}*embedding_size` embedding parameters for each relation"
        )
        self.embedding_size = embedding_size
        self.u = float(offset)  # travels to the kernels in the cfg's `rel_u` slot
        self.register_buffer("offset", torch.tensor([offset], dtype=self.entity_embedding.dtype))


class InterHT(_AuxEntityScoreFunction):
    """-||h^ o (t~^ + offset) + r - t^ o (h~^ + offset)||_p (scoring.py:1418-1572): entity rows
    [e | e~], relation rows [r]."""

    _family = L.INTERHT
    _relation_parts = 1


class TranS(_AuxEntityScoreFunction):
    """-||h^ o (t~^ + offset + r_bar) - t^ o (h~^ + offset - r_hat) + r||_p
    (scoring.py:1575-1750): entity rows [e | e~], relation rows [r | r_bar | r_hat]."""

    _family = L.TRANS
    _relation_parts = 3


class DistMult(MatrixDecompositionScoreFunction):
    """sum h * r * t (scoring.py:746-837)."""

    _family = L.DISTMULT

    def __init__(
        self,
        negative_sample_sharing: bool,
        sharding: Sharding,
        n_relation_type: int,
        embedding_size: int,
        entity_initializer: Initializer = [init_KGE_uniform],
        relation_initializer: Initializer = [init_KGE_uniform],
        inverse_relations: bool = False,
    ) -> None:
        super().__init__(negative_sample_sharing)
        self._build_tables(sharding, n_relation_type, inverse_relations, entity_initializer,
                           relation_initializer, [embedding_size], [embedding_size])
        assert (
            self.entity_embedding.shape[-1] == self.relation_embedding.shape[-1] == embedding_size
        ), "DistMult requires `embedding_size` embedding parameters for each entity and relation"
        self.embedding_size = embedding_size


class ComplEx(MatrixDecompositionScoreFunction):
    """Re-part of <h, r, conj(t)> in split [re | im] layout (scoring.py:840-946)."""

    _family = L.COMPLEX

    def __init__(
        self,
        negative_sample_sharing: bool,
        sharding: Sharding,
        n_relation_type: int,
        embedding_size: int,
        entity_initializer: Initializer = [init_KGE_normal],
        relation_initializer: Initializer = [init_KGE_normal],
        inverse_relations: bool = False,
    ) -> None:
        super().__init__(negative_sample_sharing)
        self._build_tables(sharding, n_relation_type, inverse_relations, entity_initializer,
                           relation_initializer, [2 * embedding_size], [2 * embedding_size])
        assert (
            self.entity_embedding.shape[-1]
            == self.relation_embedding.shape[-1]
            == 2 * embedding_size
        ), "ComplEx requires `2*embedding_size` embedding parameters for each entity and relation"
        self.embedding_size = embedding_size


class BoxE(DistanceBasedScoreFunction):
    """BoxE (scoring.py:1149-1415).  Entity rows [base | bump]; relation rows
    [head centre | tail centre | head width | tail width | head size | tail size]."""

    _family = L.BOXE

    def __init__(
        self,
        negative_sample_sharing: bool,
        scoring_norm: int,
        sharding: Sharding,
        n_relation_type: int,
        embedding_size: int,
        entity_initializer: Initializer = [torch.nn.init.uniform_],
        relation_initializer: Initializer = [torch.nn.init.uniform_, init_uniform_norm],
        apply_tanh: bool = True,
        dist_func_per_dim: bool = True,
        eps: float = 1e-6,
        inverse_relations: bool = False,
    ) -> None:
        super().__init__(negative_sample_sharing, scoring_norm)
        self.apply_tanh = apply_tanh
        self.dist_func_per_dim = dist_func_per_dim
        self.eps = eps
        if isinstance(entity_initializer, list):
            entity_initializer = 2 * entity_initializer
        if isinstance(relation_initializer, list):
            relation_initializer = 4 * [relation_initializer[0]] + 2 * [relation_initializer[1]]
        self._build_tables(
            sharding, n_relation_type, inverse_relations, entity_initializer, relation_initializer,
            [embedding_size, embedding_size],
            [embedding_size, embedding_size, embedding_size, embedding_size, 1, 1],
        )
        assert (
            2 * self.entity_embedding.shape[-1]
            == self.relation_embedding.shape[-1] - 2
            == 4 * embedding_size
        ), (
            "BoxE requires `2*embedding_size` embedding parameters for each entity"
            " and `4*embedding_size + 2` embedding parameters for each relation"
        )
        self.embedding_size = embedding_size
