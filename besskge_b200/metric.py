"""Link-prediction metrics — drop-in for reference `besskge/metric.py`.

`ranks_from_scores` runs the row-reduction kernel `bess_rank_from_scores`
(csrc/loss.cu) and requires CUDA tensors.  The remaining methods are tiny
elementwise maps over the [batch] rank vector (reciprocal, <= K, masking,
sum) and are expressed with torch ops on whatever device the ranks live on —
they are also used host-side on rank vectors copied back from the device.
"""
from __future__ import annotations

import re
from abc import ABC, abstractmethod
from typing import Callable, Dict, List, Optional

import torch

from . import kernels as K


class BaseMetric(ABC):
    @abstractmethod
    def __call__(self, prediction_rank: torch.Tensor) -> torch.Tensor:
        """[batch] ranks -> [batch] metric values."""


class ReciprocalRank(BaseMetric):
    def __call__(self, prediction_rank: torch.Tensor) -> torch.Tensor:
        return torch.reciprocal(prediction_rank)


class HitsAtK(BaseMetric):
    def __init__(self, k: int) -> None:
        self.K = k

    def __call__(self, prediction_rank: torch.Tensor) -> torch.Tensor:
        return (prediction_rank <= self.K).to(torch.float)


METRICS_DICT = {"mrr": ReciprocalRank, "hits@k": HitsAtK}

_MODES = {"optimistic": 0, "pessimistic": 1, "average": 2}


class Evaluation:
    """reference: metric.py:74-273."""

    def __init__(
        self,
        metric_list: List[str],
        mode: str = "average",
        worst_rank_infty: bool = False,
        reduction: str = "none",
        return_ranks: bool = False,
    ) -> None:
        if mode not in _MODES:
            raise ValueError(f"Mode {mode} not supported for evaluation")
        if reduction not in ("none", "sum"):
            raise ValueError(f"Reduction {reduction} not supported for evaluation")
        self.mode = mode
        self.return_ranks = return_ranks
        self.worst_rank_infty = worst_rank_infty
        self._reduction_name = reduction
        self.reduction: Callable[[torch.Tensor], torch.Tensor] = (
            (lambda x: x) if reduction == "none" else (lambda x: torch.sum(x, dim=0))
        )
        # hits@K entries first (in list order), then the others — same ordering
        # rule as the reference, which fixes the row order of stacked metrics
        hits = [re.search(r"hits@(\d+)", m) for m in metric_list]
        self.metrics: Dict[str, Callable[[torch.Tensor], torch.Tensor]] = {
            h[0]: HitsAtK(k=int(h[1])) for h in hits if h
        }
        self.metrics.update(
            {m: METRICS_DICT[m]() for m in list(set(metric_list) - set(self.metrics.keys()))}
        )

    def ranks_from_scores(
        self, pos_score: torch.Tensor, candidate_score: torch.Tensor
    ) -> torch.Tensor:
        """Rank of the positive among the candidates (metric.py:129-183)."""
        n, n_neg = candidate_score.shape
        pos = pos_score.reshape(-1)
        if pos.shape[0] != n:
            raise ValueError(
                "`pos_score` and `candidate_score` need to have same size at dimension 0"
            )
        K.require_cuda(pos, candidate_score)
        # metric.py:151: in place on the caller's tensor, torch defaults for the infinities
        # (nan -> -inf, +-inf -> +-FLT_MAX) — a positive masked to -inf still outranks -inf
        # candidates, and AllScoresPipeline returns -FLT_MAX for it
        pos.nan_to_num_(-torch.inf)
        pos = pos.float().contiguous()
        cand = candidate_score
        if cand.dtype != torch.float32 or cand.stride(-1) != 1:
            cand = cand.float().contiguous()
        rank = torch.empty(n, dtype=torch.float32, device=pos.device)
        K.rank_from_scores(pos, cand, n, n_neg, cand.stride(0), _MODES[self.mode],
                           self.worst_rank_infty, rank)
        return rank

    def ranks_from_indices(
        self, ground_truth: torch.Tensor, candidate_indices: torch.Tensor
    ) -> torch.Tensor:
        """Position (1-based) of the ground truth in the ORDERED candidate ids,
        worst rank if absent (metric.py:185-220)."""
        n, n_cand = candidate_indices.shape
        truth = ground_truth.reshape(-1, 1)
        if truth.shape[0] != n:
            raise ValueError(
                "`pos_score` and `candidate_score` need to have the same size for dimension 0"
            )
        worst = torch.inf if self.worst_rank_infty else float(n_cand + 1)
        pos = torch.arange(1, n_cand + 1, dtype=torch.float32, device=truth.device)
        return torch.where(truth == candidate_indices, pos, worst).min(dim=-1)[0]

    def dict_metrics_from_ranks(
        self, batch_rank: torch.Tensor, triple_mask: Optional[torch.Tensor] = None
    ) -> Dict[str, torch.Tensor]:
        out = {}
        for name, fn in self.metrics.items():
            val = fn(batch_rank)
            if triple_mask is not None:
                val = torch.where(triple_mask, val, torch.zeros((), dtype=val.dtype, device=val.device))
            out[name] = self.reduction(val)
        return out

    def stacked_metrics_from_ranks(
        self, batch_rank: torch.Tensor, triple_mask: Optional[torch.Tensor] = None
    ) -> torch.Tensor:
        """[1, n_metrics(, batch)] in the order of `self.metrics`."""
        return torch.stack(
            list(self.dict_metrics_from_ranks(batch_rank, triple_mask).values())
        ).unsqueeze(0)
