"""General-purpose utilities (reference: besskge/utils.py): the device helpers
`gather_indices`, `complex_multiplication`, `complex_rotation` (CUDA kernels through the
C-ABI; CUDA tensors only, like every device entry point of this package) and the host-side
entity filter of the evaluation pipeline."""
from __future__ import annotations

from typing import Union

import numpy as np
import torch
from numpy.typing import NDArray

from . import _lib as L


def gather_indices(x: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """2-D take-along-dim (reference: utils.py:10-33): `x` (a, e), `index` (b, k) ->
    (max(a, b), k) with out[i, j] = x[i, index[i, j]]; b == 1 takes the same columns from
    every row of `x`, a == 1 makes every index row read `x[0]`, otherwise a == b."""
    L.require_cuda(x, index)
    if x.dim() != 2 or index.dim() != 2:
        raise ValueError("gather_indices expects 2-dimensional x and index")
    x = x.contiguous()
    idx = index.to(torch.int32).contiguous()
    a, e = x.shape
    b, k = idx.shape
    out = torch.empty(max(a, b), k, dtype=x.dtype, device=x.device)
    L.call("bess_take_along_rows", x.data_ptr(), a, e, x.element_size(), idx.data_ptr(), b, k,
           out.data_ptr(), L.stream_ptr(x.device))
    return out


def _complex(v1: torch.Tensor, v2: torch.Tensor, rotate: bool) -> torch.Tensor:
    L.require_cuda(v1, v2)
    if v1.dtype != v2.dtype:
        raise TypeError("operands must share one dtype")
    lead = v1.shape[:-1]
    a = v1.reshape(-1, v1.shape[-1]).contiguous()
    e = a.shape[-1] // 2
    b = v2.expand(*lead, v2.shape[-1]).reshape(-1, v2.shape[-1]).contiguous()
    if a.shape[-1] != 2 * e or b.shape[-1] != (e if rotate else 2 * e) or b.shape[0] != a.shape[0]:
        raise ValueError(f"incompatible shapes {tuple(v1.shape)} and {tuple(v2.shape)}")
    out = torch.empty_like(a)
    L.call("bess_complex_mul", L.dtype_code(a.dtype), a.data_ptr(), b.data_ptr(), a.shape[0], e,
           int(rotate), out.data_ptr(), L.stream_ptr(a.device))
    return out.view(*lead, 2 * e)


def complex_multiplication(v1: torch.Tensor, v2: torch.Tensor) -> torch.Tensor:
    """Row-wise complex product of (a, 2e) tensors, `[:, :e]` real and `[:, e:]` imaginary
    parts (reference: utils.py:72-89)."""
    return _complex(v1, v2, False)


def complex_rotation(v: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
    """Rotate `v` (a, 2e) by the unit complex numbers `cos r + i sin r`, `r` (a, e)
    (reference: utils.py:92-112; full-precision sin / cos as the reference does off-IPU)."""
    return _complex(v, r, True)


class EntityFilterIndex:
    """`filter_triples` ordered once by (entity, relation); `query(triples)` returns the
    sparse filter of get_entity_filter for a batch without re-sorting the filter set."""

    def __init__(self, filter_triples: Union[torch.Tensor, NDArray[np.int64]], filter_mode: str,
                 n_relation_hint: int = 0) -> None:
        if filter_mode == "t":
            self.ent_col = 0
        elif filter_mode == "h":
            self.ent_col = 2
        else:
            raise ValueError("`filter_mode` needs to be either 'h' or 't'")
        ft = (filter_triples.cpu().numpy() if isinstance(filter_triples, torch.Tensor)
              else np.asarray(filter_triples))
        self.ft = ft.reshape(-1, 3).astype(np.int64)
        self.n_rel = max(int(self.ft[:, 1].max()) + 1 if self.ft.shape[0] else 1, n_relation_hint, 1)
        fkey = self.ft[:, self.ent_col] * self.n_rel + self.ft[:, 1]
        self.order = np.argsort(fkey, kind="stable")
        self.skey = fkey[self.order]

    def query(self, triples: Union[torch.Tensor, NDArray[np.int64]]) -> torch.Tensor:
        tr = triples.cpu().numpy() if isinstance(triples, torch.Tensor) else np.asarray(triples)
        tr = tr.reshape(-1, 3).astype(np.int64)
        if tr.shape[0] == 0 or self.ft.shape[0] == 0:
            return torch.zeros((0, 2), dtype=torch.int64)
        qkey = tr[:, self.ent_col] * self.n_rel + tr[:, 1]
        # a relation id beyond the filter set's range can never match (and must not alias a key)
        valid = tr[:, 1] < self.n_rel
        lo = np.searchsorted(self.skey, qkey, side="left")
        hi = np.searchsorted(self.skey, qkey, side="right")
        cnt = np.where(valid, hi - lo, 0)
        total = int(cnt.sum())
        rows = np.repeat(np.arange(tr.shape[0], dtype=np.int64), cnt)
        starts = np.cumsum(cnt) - cnt
        within = np.arange(total, dtype=np.int64) - np.repeat(starts, cnt)
        src = self.order[np.repeat(lo, cnt) + within]
        ents = self.ft[src, 2 - self.ent_col]
        return torch.from_numpy(np.stack([rows, ents], axis=1))


def get_entity_filter(
    triples: Union[torch.Tensor, NDArray[np.int64]],
    filter_triples: Union[torch.Tensor, NDArray[np.int64]],
    filter_mode: str,
) -> torch.Tensor:
    """For each triple (h, r, t) of `triples` [x, 3]: the entities e such that
    (h, r, e) — filter_mode "t" — or (e, r, t) — "h" — appears in `filter_triples`
    [y, 3].  Returns the sparse filter [z, 2]: rows (i, e), i = index in `triples`.

    Same result AND ROW ORDER as the reference (utils.py:36-69), which builds the dense
    [x, y] comparison and calls `nonzero` (row-major: by i, then by position in
    `filter_triples`).  Here it is a sort-merge join — O((x + y) log y) instead of
    O(x * y): `filter_triples` is ordered once by the key (entity, relation) with a stable
    sort, and every query reads the contiguous run of its key, whose members are in
    increasing original position."""
    return EntityFilterIndex(filter_triples, filter_mode).query(triples)
