"""Loss functions — drop-in for reference `besskge/loss.py`.

`forward(positive_score [S], negative_score [S, N], triple_weight [S] | [1])`
returns the fp32 replica loss (a SUM over the micro-batch, loss.py:131-134,
191-195, 249-251).  Executed by `csrc/loss.cu` (one CTA per score row, fused
loss + dL/dscore); inputs must be CUDA tensors.  Inside the fused training
step the same kernel also yields the score gradients.
"""
from __future__ import annotations

from abc import ABC
from typing import Optional, Tuple

import torch

from . import _lib as L
from . import kernels as K


class BaseLossFunction(torch.nn.Module, ABC):
    negative_adversarial_sampling: bool
    negative_adversarial_scale: torch.Tensor
    loss_scale: torch.Tensor
    _kind: int = -1

    def kernel_params(self) -> dict:
        return dict(
            kind=self._kind,
            margin=float(getattr(self, "margin", torch.tensor(0.0))),
            adversarial=bool(self.negative_adversarial_sampling),
            adv_scale=float(self.negative_adversarial_scale),
            loss_scale=float(self.loss_scale),
            n_entity=int(getattr(self, "n_entity", 2)),
        )

    def fwd_bwd(
        self,
        positive_score: torch.Tensor,
        negative_score: torch.Tensor,
        triple_weight: torch.Tensor,
        d_pos: Optional[torch.Tensor] = None,
        d_neg: Optional[torch.Tensor] = None,
    ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(loss [], dL/dpos [S], dL/dneg [S, N]); fp32 CUDA tensors."""
        K.require_cuda(positive_score, negative_score, triple_weight)
        pos = positive_score.float().contiguous()
        neg = negative_score
        if neg.dtype != torch.float32 or neg.stride(-1) != 1:
            neg = neg.float().contiguous()
        n, n_neg = neg.shape
        w = triple_weight.float().contiguous().reshape(-1)
        dev = pos.device
        row_loss = torch.empty(n, dtype=torch.float32, device=dev)
        if d_pos is None:
            d_pos = torch.empty(n, dtype=torch.float32, device=dev)
        if d_neg is None:
            d_neg = torch.empty(n, n_neg, dtype=torch.float32, device=dev)
        p = self.kernel_params()
        K.loss_fwd_bwd(p["kind"], p["margin"], p["adversarial"], p["adv_scale"], p["loss_scale"],
                       p["n_entity"], pos, neg, n, n_neg, neg.stride(0), w, row_loss, d_pos, d_neg)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        K.sum_f32(row_loss, n, loss)
        if self._kind == L.LOSS_SOFTMAX_CE and neg.data_ptr() != negative_score.data_ptr():
            # the reference adjusts negative_score in place (loss.py:233-237)
            negative_score.copy_(neg)
        return loss, d_pos, d_neg

    def forward(
        self,
        positive_score: torch.Tensor,
        negative_score: torch.Tensor,
        triple_weight: torch.Tensor,
    ) -> torch.Tensor:
        return self.fwd_bwd(positive_score, negative_score, triple_weight)[0]


class MarginBasedLossFunction(BaseLossFunction, ABC):
    def __init__(
        self,
        margin: float,
        negative_adversarial_sampling: bool,
        negative_adversarial_scale: float = 1.0,
        loss_scale: float = 1.0,
    ) -> None:
        super().__init__()
        self.negative_adversarial_sampling = negative_adversarial_sampling
        self.negative_adversarial_scale = torch.tensor(
            negative_adversarial_scale, dtype=torch.float32
        )
        self.loss_scale = torch.tensor(loss_scale, dtype=torch.float32)
        self.margin = torch.tensor(margin, dtype=torch.float32)


class LogSigmoidLoss(MarginBasedLossFunction):
    """-1/2 sum_i w_i [logsig(pos_i + m) + sum_j w_ij logsig(-neg_ij - m)]
    (loss.py:109-134)."""

    _kind = L.LOSS_LOGSIGMOID


class MarginRankingLoss(MarginBasedLossFunction):
    """sum_i w_i sum_j w_ij relu(neg_ij - pos_i + m) (loss.py:137-195)."""

    _kind = L.LOSS_MARGIN_RANKING

    def __init__(
        self,
        margin: float,
        negative_adversarial_sampling: bool,
        negative_adversarial_scale: float = 1.0,
        loss_scale: float = 1.0,
        activation_function: str = "relu",
    ) -> None:
        super().__init__(margin, negative_adversarial_sampling, negative_adversarial_scale,
                         loss_scale)
        if activation_function != "relu":
            raise ValueError(
                f"Activation function {activation_function} not supported for MarginRankingLoss"
            )
        self.activation = torch.nn.functional.relu


class SampledSoftmaxCrossEntropyLoss(BaseLossFunction):
    """Sampled-softmax cross entropy with the log((E-1)/N) correction
    (loss.py:198-251); adjusts `negative_score` in place like the reference."""

    _kind = L.LOSS_SOFTMAX_CE

    def __init__(self, n_entity: int, loss_scale: float = 1.0) -> None:
        super().__init__()
        self.negative_adversarial_sampling = False
        self.negative_adversarial_scale = torch.tensor(0.0, dtype=torch.float32)
        self.loss_scale = torch.tensor(loss_scale, dtype=torch.float32)
        self.n_entity = n_entity
