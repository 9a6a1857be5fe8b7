"""High-level inference pipeline (reference: besskge/pipeline.py:23-320).

`AllScoresPipeline` scores (h, r, ?) / (?, r, t) queries against all entities (or a
candidate subset), filters known completions, and computes ranks / metrics / top-k.

B200 shape of the loop.  The reference brings every `window_size` block of scores back
to the host and does the rest there with numpy / torch fancy indexing: `np.unique` to
put block columns into global-entity order, boolean row masks, `-inf` writes for
non-candidates and filters, `torch.topk` (pipeline.py:266-313) — O(batch x entities)
host work per batch plus an O(batch x filter-set) dense comparison for the filters
(utils.py:63-66).  Here the whole batch stays in HBM: one pass of the windowed scorer
over every shard (`AllScoresBESS.score_all`: tcgen05 GEMM for DistMult / ComplEx),
`bess_select_scores` (column map = padding removal + global order + candidate mask in one
gather), `bess_pairs_get / set` (ground-truth scores, sparse filters),
`bess_rank_from_scores`, `bess_topk_merge`.  The host only builds index lists; the
filter join is a sort-merge (`utils.EntityFilterIndex`).  Results are identical to the
reference's (same masks, same order of rows) and are returned as host tensors like the
reference's.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Union

import numpy as np
import torch
from numpy.typing import NDArray

from . import kernels as K
from .batch_sampler import ShardedBatchSampler
from .bess import AllScoresBESS, _Placement
from .metric import Evaluation
from .negative_sampler import PlaceholderNegativeSampler
from .scoring import BaseScoreFunction
from .utils import EntityFilterIndex


class AllScoresPipeline(torch.nn.Module):
    """To be used with a batch sampler over an "h_shard" / "t_shard"-partitioned triple
    set (pipeline.py:23-190).  `use_ipu_model` is accepted for signature compatibility."""

    def __init__(
        self,
        batch_sampler: ShardedBatchSampler,
        corruption_scheme: str,
        score_fn: BaseScoreFunction,
        evaluation: Optional[Evaluation] = None,
        filter_triples: Optional[List[Union[torch.Tensor, NDArray[np.int32]]]] = None,
        candidate_ents: Optional[Union[torch.Tensor, NDArray[np.int32]]] = None,
        return_scores: bool = False,
        return_topk: bool = False,
        k: int = 10,
        window_size: int = 1000,
        use_ipu_model: bool = False,
    ) -> None:
        super().__init__()
        self.batch_sampler = batch_sampler
        if not (evaluation or return_scores):
            raise ValueError("Nothing to return. Provide `evaluation` or set `return_scores=True`")
        if corruption_scheme not in ["h", "t"]:
            raise ValueError("corruption_scheme needs to be either 'h' or 't'")
        if corruption_scheme == "h" and self.batch_sampler.triple_partition_mode != "t_shard":
            raise ValueError("Corruption scheme 'h' requires 't-shard'-partitioned triples")
        elif corruption_scheme == "t" and self.batch_sampler.triple_partition_mode != "h_shard":
            raise ValueError("Corruption scheme 't' requires 'h-shard'-partitioned triples")
        if return_topk and not 1 <= k <= 64:
            raise ValueError("return_topk supports 1 <= k <= 64")
        self.candidate_sampler = PlaceholderNegativeSampler(corruption_scheme=corruption_scheme)
        self.score_fn = score_fn
        self.evaluation = evaluation
        self.return_scores = return_scores
        self.return_topk = return_topk
        self.k = k
        self.window_size = window_size
        self.corruption_scheme = corruption_scheme
        self.bess_module = AllScoresBESS(self.candidate_sampler, self.score_fn, self.window_size)
        self.dl = self.batch_sampler.get_dataloader(shuffle=False)
        sharding = self.bess_module.sharding

        self.filter_triples: Optional[torch.Tensor] = None
        self._filter_index: Optional[EntityFilterIndex] = None
        if filter_triples:
            if not self.batch_sampler.return_triple_idx:
                raise ValueError("filter_triples needs a batch sampler with return_triple_idx=True")
            # global ids of the (locally indexed) known entity (pipeline.py:150-171)
            local_id_col = 0 if self.batch_sampler.triple_partition_mode == "h_shard" else 2
            offs = np.concatenate([np.array([0]), np.cumsum(batch_sampler.triple_counts)])
            parts = []
            for i in range(len(offs) - 1):
                shard_triples = np.copy(batch_sampler.triples[offs[i]:offs[i + 1]])
                shard_triples[:, local_id_col] = sharding.shard_and_idx_to_entity[i][
                    shard_triples[:, local_id_col]]
                parts.append(shard_triples)
            self.triples = torch.from_numpy(np.concatenate(parts, axis=0))
            self.filter_triples = torch.concat(
                [tr if isinstance(tr, torch.Tensor) else torch.from_numpy(tr)
                 for tr in filter_triples], dim=0)
            self._filter_index = EntityFilterIndex(self.filter_triples, corruption_scheme)
        self.candidate_mask: Optional[torch.Tensor] = None
        if candidate_ents is not None:
            cand = (candidate_ents.cpu().numpy() if isinstance(candidate_ents, torch.Tensor)
                    else np.asarray(candidate_ents))
            self.candidate_mask = torch.from_numpy(np.setdiff1d(np.arange(sharding.n_entity), cand))
        # column of entity e in the [*, n_shard * Es] block scores; -1 = not a candidate
        Es = sharding.max_entity_per_shard
        col = (sharding.entity_to_shard.astype(np.int64) * Es
               + sharding.entity_to_idx.astype(np.int64)).astype(np.int32)
        if self.candidate_mask is not None:
            col[self.candidate_mask.numpy()] = -1
        self._col_of_entity = torch.from_numpy(col)
        self._col_dev: Optional[torch.Tensor] = None

    def forward(self) -> Dict[str, Any]:
        """Scores of all completions and (possibly) metrics; `triple_idx` (wrt
        partitioned_triple_set.triples) orders the rows of every returned tensor."""
        bess = self.bess_module
        sharding = bess.sharding
        n = sharding.n_shard
        E = sharding.n_entity
        dev = self.score_fn.entity_embedding.device
        if dev.type != "cuda":
            from ._lib import BessLibraryError
            raise BessLibraryError("besskge_b200 has no CPU path: move the score function to a CUDA device")
        if self._col_dev is None or self._col_dev.device != dev:
            self._col_dev = self._col_of_entity.to(dev)
        pl = _Placement(n)
        scheme = self.corruption_scheme
        neg_inf = float("-inf")

        scores: List[torch.Tensor] = []
        ids: List[torch.Tensor] = []
        metrics: List[Dict[str, torch.Tensor]] = []
        ranks: List[torch.Tensor] = []
        topk_ids: List[torch.Tensor] = []
        n_triple = 0
        for batch in iter(self.dl):
            batch = dict(batch)
            triple_mask = batch.pop("triple_mask")  # [bps, n, S]
            ground_truth = None
            if scheme == "h" and "head" in batch:
                ground_truth = batch.pop("head")
            elif scheme == "t" and "tail" in batch:
                ground_truth = batch.pop("tail")
            triple_id = batch.pop("triple_idx") if self.batch_sampler.return_triple_idx else None
            # rows of this process: every shard in local mode, its own shard when distributed
            sel = triple_mask.clone()
            if pl.distributed:
                own = torch.zeros_like(sel)
                own[:, pl.rank] = True
                sel &= own
            if triple_id is not None:
                ids.append(triple_id[sel])
            n_triple += int(sel.sum())
            inp = {k_: v.flatten(end_dim=1) for k_, v in batch.items()}
            block = bess.score_all(**inp)  # [bps * R * S, n * Es]
            # valid rows, in (step, shard, triple) order, as indices into `block`
            if pl.distributed:
                row_src = torch.nonzero(triple_mask[:, pl.rank].reshape(-1)).reshape(-1)
            else:
                row_src = torch.nonzero(triple_mask.reshape(-1)).reshape(-1)
            nv = int(row_src.numel())
            out = torch.empty(nv, E, dtype=torch.float32, device=dev)
            if nv == 0:
                continue
            K.select_scores(block, row_src.to(device=dev, dtype=torch.int32), nv, self._col_dev,
                            neg_inf, out)
            truth = true_scores = None
            if ground_truth is not None:
                truth = ground_truth[sel].to(device=dev, dtype=torch.int32)
                true_scores = torch.empty(nv, dtype=torch.float32, device=dev)
                K.pairs_get(out, None, truth, true_scores)
            if self._filter_index is not None:
                bf = self._filter_index.query(self.triples[triple_id[sel]])
                if bf.shape[0]:
                    K.pairs_set(out, bf[:, 0].to(device=dev, dtype=torch.int32),
                                bf[:, 1].to(device=dev, dtype=torch.int32), None, neg_inf)
            if self.evaluation:
                assert ground_truth is not None, "Evaluation requires providing ground truth entities"
                K.pairs_set(out, None, truth, None, neg_inf)
                batch_ranks = self.evaluation.ranks_from_scores(true_scores, out)
                metrics.append({m: v.cpu() for m, v in
                                self.evaluation.dict_metrics_from_ranks(batch_ranks).items()})
                if self.evaluation.return_ranks:
                    ranks.append(batch_ranks.cpu())
            if ground_truth is not None:
                K.pairs_set(out, None, truth, true_scores)
            if self.return_topk:
                best_s = torch.full((nv, self.k), neg_inf, dtype=torch.float32, device=dev)
                best_i = torch.full((nv, self.k), -1, dtype=torch.int32, device=dev)
                K.topk_merge(out, out.stride(0), nv, E, None, 0, 0, best_s, best_i, self.k)
                topk_ids.append(best_i.cpu().to(torch.int64))
            if self.return_scores:
                scores.append(out.cpu())

        res: Dict[str, Any] = dict()
        if scores:
            res["scores"] = torch.concat(scores, dim=0)
        if topk_ids:
            res["topk_global_id"] = torch.concat(topk_ids, dim=0)
        if ids:
            res["triple_idx"] = torch.concat(ids, dim=0)
        if self.evaluation and metrics:
            final_metrics = dict()
            for m in metrics[0].keys():
                final_metrics[m] = self.evaluation.reduction(
                    torch.concat([met[m].reshape(-1) for met in metrics]))
            res["metrics"] = final_metrics
            res["metrics_avg"] = {m: v.sum() / n_triple for m, v in final_metrics.items()}
            if ranks:
                res["ranks"] = torch.concat(ranks, dim=0)
        return res
