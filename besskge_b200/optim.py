"""Optimizer descriptions for the fused BESS training step.

The reference trains through PopTorch (`poptorch.optim.SGD / AdamW`, notebook
1 cell 28, notebook 3 cell 18); neither the backward pass nor the optimizer is
part of `/root/reference`.  Parity is therefore defined against dense
`torch.optim.SGD` / `torch.optim.AdamW` applied to the autograd gradient of the
reference forward.  These classes only carry hyper-parameters;
the update itself is `bess_scatter_sgd` / `bess_opt_dense` (csrc/scatter.cu).
"""
from __future__ import annotations

import dataclasses

from . import _lib as L


@dataclasses.dataclass
class SGD:
    """torch.optim.SGD semantics.  With momentum == 0 and weight_decay == 0 the
    entity update is sparse and exactly equal to the dense update."""

    lr: float
    momentum: float = 0.0
    dampening: float = 0.0
    weight_decay: float = 0.0

    @property
    def kind(self) -> int:
        return L.OPT_SGDM if self.momentum != 0.0 else L.OPT_SGD

    @property
    def sparse_exact(self) -> bool:
        return self.momentum == 0.0 and self.weight_decay == 0.0


@dataclasses.dataclass
class AdamW:
    """torch.optim.AdamW semantics (decoupled weight decay); dense: every row
    of the shard moves every step once the moments are non-zero."""

    lr: float = 1e-3
    betas: tuple = (0.9, 0.999)
    eps: float = 1e-8
    weight_decay: float = 1e-2

    kind = L.OPT_ADAMW
    sparse_exact = False
    momentum = 0.0
    dampening = 0.0
