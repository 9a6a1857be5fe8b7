"""Device-resident sampling feed.

`ShardedBatchSampler.__getitem__` (reference: batch_sampler.py:138-196) does two things on
the host for every step: it draws indices with numpy's generator, and it fancy-indexes large
arrays with them — `triples[sample_idx]` (:159-162) and, for predetermined candidates,
`padded_negatives[sample_idx]` / `mask[sample_idx]` (negative_sampler.py:394-401).  The
second part is what caps end-to-end throughput (≈ 6-8 M triples/s on one host thread for the
wikikg2 shape, and 5 MB of host->device traffic per step for 2048 queries x 500 candidates).

`DeviceBatchFeed` keeps the indexed arrays in HBM and moves only the drawn indices:
  host   : the sampler's OWN `sample_triples` / negative-sampler RNG calls, in the same order, so
           the random stream — and therefore every batch — is bit-identical to `sampler[idx]`;
  device : `bess_gather_rows` over 16-byte rows of the (h, r, t, 0) triple table and of the
           padded candidate / mask tables, then strided views for the reference's layouts
           (tail's shard axes transposed, batch_sampler.py:163-167; candidates regrouped by the
           shard that stores them, negative_sampler.py `ent_rearrange_pattern`).
The result is a dict of CUDA tensors with the keys, shapes, dtypes and values of `sampler[idx]`;
the BESS modules accept it unchanged.  Cases that only need index-sized host work keep the
sampler's host code (random negatives, flat candidate lists, the "ht" candidate split,
frequency weights).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import einops
import numpy as np
import torch

from . import kernels as K
from ._lib import BessLibraryError
from .batch_sampler import ShardedBatchSampler
from .negative_sampler import TripleBasedShardedNegativeSampler


def _pad_cols(a: np.ndarray, multiple: int) -> np.ndarray:
    """[rows, cols] -> [rows, cols rounded up to `multiple`], zero padded."""
    cols = a.shape[1]
    want = (cols + multiple - 1) // multiple * multiple
    if want == cols:
        return np.ascontiguousarray(a)
    out = np.zeros((a.shape[0], want), dtype=a.dtype)
    out[:, :cols] = a
    return out


class DeviceBatchFeed:
    """`feed[idx]` == `sampler[idx]` (same RNG stream), as CUDA tensors."""

    def __init__(self, sampler: ShardedBatchSampler, device: torch.device) -> None:
        device = torch.device(device)
        if device.type != "cuda":
            raise BessLibraryError("DeviceBatchFeed needs a CUDA device (there is no CPU path)")
        self.sampler = sampler
        self.device = device
        trip = np.zeros((sampler.triples.shape[0], 4), dtype=np.int32)
        trip[:, :3] = sampler.triples
        self._triples = torch.from_numpy(trip).to(device)
        ns = sampler.negative_sampler
        self._cand: Optional[torch.Tensor] = None
        self._mask: Optional[torch.Tensor] = None
        self._cand_on_device = (
            isinstance(ns, TripleBasedShardedNegativeSampler)
            and not ns.flat_negative_format
            and ns.corruption_scheme in ("h", "t")
        )
        if self._cand_on_device:
            n, L = ns.padded_negatives.shape[1], ns.padded_negatives.shape[2]
            self._n_neg_shard, self._L = n, L
            # rows of n * L int32 ids / n * L mask bytes, padded to 16-byte multiples
            self._cand = torch.from_numpy(_pad_cols(
                ns.padded_negatives.reshape(-1, n * L).astype(np.int32), 4)).to(device)
            self._mask = torch.from_numpy(_pad_cols(
                ns.mask.reshape(-1, n * L).astype(np.uint8), 16)).to(device)

    # ------------------------------------------------------------------
    def _gather(self, table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """rows `idx` of a device table whose rows are 16-byte multiples (bit copy)."""
        out = torch.empty(idx.numel(), table.shape[1], dtype=table.dtype, device=self.device)
        # the gather kernel moves 128-bit words; int32 / uint8 rows are passed as fp32 words
        K.gather_rows(table.view(torch.float32), idx, out.view(torch.float32))
        return out

    def __getitem__(self, idx: List[int]) -> Dict[str, torch.Tensor]:
        sm = self.sampler
        ns = sm.negative_sampler
        dev = self.device
        sampled = sm.sample_triples(idx)
        if sm.duplicate_batch:  # batch_sampler.py:150-157
            sampled = {
                k: einops.repeat(v, "step shard ... triple -> step shard ... (2 triple)")
                for k, v in sampled.items()
            }
        sample_idx = sampled.pop("sample_idx")
        shape = sample_idx.shape
        si = torch.from_numpy(np.ascontiguousarray(sample_idx.astype(np.int32))).pin_memory().to(
            dev, non_blocking=True).view(-1)
        hrt = self._gather(self._triples, si)  # [M, 4]
        head = hrt[:, 0].reshape(shape)
        relation = hrt[:, 1].reshape(shape)
        tail = hrt[:, 2].reshape(shape)
        if sm.triple_partition_mode == "ht_shardpair":
            tail = tail.transpose(1, 2)  # block (shard_t, shard_h) is gathered on shard_t
        batch: Dict[str, torch.Tensor] = {
            "head": head.contiguous(), "relation": relation.contiguous(), "tail": tail.contiguous()}
        for k, v in sampled.items():  # triple_mask of the rigid sampler
            batch[k] = torch.from_numpy(np.ascontiguousarray(v)).to(dev, non_blocking=True)

        if self._cand_on_device:
            n, L = self._n_neg_shard, self._L
            bps, n_sh = shape[0], shape[1]
            ents = self._gather(self._cand, si)[:, :n * L]
            mask = self._gather(self._mask, si)[:, :n * L]
            # [step, shard, (... triple), shard_neg, idx_neg]
            ents = ents.reshape(bps, n_sh, -1, n, L)
            mask = mask.reshape(bps, n_sh, -1, n, L)
            # entities are consumed on the shard that stores them: step shard_neg shard triple idx
            batch["negative"] = ents.permute(0, 3, 1, 2, 4).contiguous()
            if ns.mask_on_gather:
                batch["negative_mask"] = mask.permute(0, 3, 1, 2, 4).contiguous().to(torch.bool)
            else:
                batch["negative_mask"] = mask.contiguous().to(torch.bool)
            if ns.return_sort_idx:
                sort_idx = ns.sort_neg_idx[sample_idx]
                batch["negative_sort_idx"] = torch.from_numpy(np.ascontiguousarray(
                    sort_idx.reshape(bps, n_sh, -1, sort_idx.shape[-1]))).to(dev, non_blocking=True)
        else:
            neg = ns(sample_idx)
            if "negative_entities" in neg:
                batch["negative"] = torch.from_numpy(
                    neg.pop("negative_entities").astype(np.int32)).to(dev, non_blocking=True)
            for k, v in neg.items():
                batch[k] = torch.from_numpy(np.ascontiguousarray(v)).to(dev, non_blocking=True)

        if sm.dummy in ("head", "tail"):
            batch.pop(sm.dummy)
        if sm.hrt_freq_weighting:  # batch_sampler.py:184-191 (index-sized host work)
            w = sm.hrt_weights[sample_idx].reshape(shape[0], shape[1], -1)
            w = w / np.sum(w, axis=-1, keepdims=True) * sm.shard_bs
            batch["triple_weight"] = torch.from_numpy(w.astype(np.float32)).to(dev, non_blocking=True)
        if sm.return_triple_idx:
            batch["triple_idx"] = torch.from_numpy(np.ascontiguousarray(sample_idx)).to(
                dev, non_blocking=True)
        return batch
