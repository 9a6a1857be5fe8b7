"""BESS distribution modules — drop-in for reference `besskge/bess.py`
(`EmbeddingMovingBessKGE`, `ScoreMovingBessKGE`, `TopKQueryBessKGE`), built
B200-first.

What the reference expresses as torch ops traced by PopTorch for IPUs
(`entity_embedding[cat(head, tail, negative)]` -> `all_to_all` -> score ->
loss, bess.py:322-468) is here one stream-ordered sequence of hand-written
sm_100a kernels per micro-batch, called through the C-ABI:

  routed gather  ->  (exchange)  ->  score_triple  ->  query prologue +
  negative scoring  ->  masks  ->  fused loss / dL/dscore  ->  backward
  kernels  ->  (reverse exchange)  ->  radix sort + deterministic segmented
  scatter-add fused with the optimizer.

Replica placement (the reference is single-controller with PopTorch replicas):
  * local mode      — all `n_shard` shards live on THIS process's GPU; the
    AllToAll is folded into the gather, which writes every row straight into
    the receive buffer of the replica that scores it;
  * distributed mode — one process per GPU (`torch.distributed` for rendezvous
    and the peer mapping: torch symmetric memory or this library's CUDA-IPC
    calls); this process owns shard `rank`; tail / negative rows reach the
    destination GPU's receive buffer over NVLink either by remote stores of
    the gather kernel or by copy-engine transfers of a local send buffer
    (`_use_copy_engine`), gradients travel back the same way, and flag
    handshakes replace the collectives (csrc/peer.cu).  The whole distributed
    training call is one CUDA graph per rank.  NCCL `all_to_all_single`
    remains selectable for A/B runs (`USE_PEER_EXCHANGE`).
Inputs keep the reference layout: a leading `batches_per_step * n_shard` axis
(step-major, shard-minor); outputs are concatenated in the same order
(tests/test_bess.py:181-196 of the reference).
"""
from __future__ import annotations

import dataclasses
import os
from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib as L
from . import kernels as K
from .kernels import BAD_NEGATIVE_SCORE
from .loss import BaseLossFunction
from .metric import Evaluation
from .negative_sampler import (
    PlaceholderNegativeSampler,
    ShardedNegativeSampler,
    TripleBasedShardedNegativeSampler,
)
from .optim import SGD, AdamW
from .scoring import BaseScoreFunction

__all__ = [
    "BAD_NEGATIVE_SCORE",
    "BessKGE",
    "EmbeddingMovingBessKGE",
    "ScoreMovingBessKGE",
    "TopKQueryBessKGE",
    "AllScoresBESS",
    "ScatterAllToAllBessKGE",
    "AllScatterAllGatherBessKGE",
    "TrainingModel",
    "training_model",
]


# ---------------------------------------------------------------------------
# replica placement
# ---------------------------------------------------------------------------
# Tests only: run all shards on this process's GPU even though torch.distributed
# is initialised (the local-mode result is the reference for the distributed one).
FORCE_LOCAL = False


class _Placement:
    """Which shards this process computes and how blocks are exchanged."""

    def __init__(self, n_shard: int) -> None:
        self.n_shard = n_shard
        dist = torch.distributed
        self.distributed = bool(
            not FORCE_LOCAL and dist.is_available() and dist.is_initialized()
            and dist.get_world_size() > 1
        )
        if self.distributed:
            if dist.get_world_size() != n_shard:
                raise ValueError(
                    f"distributed BESS needs world_size == n_shard ({dist.get_world_size()} vs {n_shard})"
                )
            self.rank = dist.get_rank()
            self.shards = [self.rank]
        else:
            self.rank = 0
            self.shards = list(range(n_shard))

    @property
    def n_local(self) -> int:
        return len(self.shards)

    def all_to_all(self, recv: torch.Tensor, send: torch.Tensor) -> None:
        """recv[j] <- send_j[rank]: equal-split AllToAll (bess.py:348-350)."""
        torch.distributed.all_to_all_single(recv.view(-1), send.view(-1))


# Distributed mode exchanges blocks with this process's own kernels over peer
# (symmetric) memory; the NCCL path is kept selectable for A/B tests.
USE_PEER_EXCHANGE = True

_ALIGN = 256


def _up(x: int, a: int = _ALIGN) -> int:
    return (x + a - 1) // a * a


class _PeerExchange:
    """This rank's symmetric-memory receive buffers and flag rows.

    Layout (identical on every rank, so peer j's address of a region is
    `ptrs[j] + offset`): 3 flag rows (forward rows / gradient rows / relation
    partials) | TN receive buffer [n, per, W] (table dtype) | gradient receive
    buffer [n, per, W] fp32 | relation partial slots [n, count] fp32.
    Allocation and (re)sizing are collective: every rank calls `ensure` with
    the same sizes at the same point of the step."""

    FLAG_BYTES = 3 * 64 * 4

    def __init__(self, device: torch.device, n: int, rank: int) -> None:
        self.device, self.n, self.rank = device, n, rank
        self.sizes = (0, 0, 0)
        self.buf: Optional[torch.Tensor] = None
        self._keep: List[Any] = []
        # sequence counters are monotonic for the life of the exchange: a re-allocation starts
        # with zeroed flag rows, and every later signal carries a larger number than any earlier
        # one, so graphs captured against an older (still mapped) buffer stay self-consistent
        self.counters = torch.zeros(3, dtype=torch.int32, device=device)
        self.generation = 0  # moves on every re-allocation (captured graphs must be dropped)

    def ensure(self, tn_bytes: int, grad_bytes: int, rel_bytes: int) -> None:
        want = (_up(tn_bytes), _up(grad_bytes), _up(rel_bytes))
        if self.buf is not None and all(w <= h for w, h in zip(want, self.sizes)):
            return
        sizes = tuple(max(w, h) for w, h in zip(want, self.sizes))
        total = _up(self.FLAG_BYTES) + sum(sizes)
        torch.cuda.synchronize(self.device)
        torch.distributed.barrier()
        if self._transport() == "ipc":
            # this library's own mapping: cudaMalloc + CUDA IPC handles (csrc/peer.cu), the
            # handles exchanged through the process group's object collective
            pbuf = K.PeerBuffer(total, self.device)
            handles: List[Any] = [None] * self.n
            torch.distributed.all_gather_object(handles, pbuf.export())
            ptrs = [pbuf.ptr if j == self.rank else K.peer_import(handles[j], self.device)
                    for j in range(self.n)]
            buf, hdl = pbuf.tensor, pbuf
        else:
            import torch.distributed._symmetric_memory as symm

            group = torch.distributed.group.WORLD
            try:
                symm.enable_symm_mem_for_group(group.group_name)
            except Exception:  # newer torch enables it implicitly
                pass
            buf = symm.empty(total, dtype=torch.uint8, device=self.device)
            hdl = symm.rendezvous(buf, group)
            buf.zero_()
            ptrs = [int(p) for p in hdl.buffer_ptrs]
        self.generation += 1
        torch.cuda.synchronize(self.device)
        torch.distributed.barrier()  # nobody signals before every flag row is zero
        self._keep.append((self.buf, getattr(self, "hdl", None)))  # peers may still map it
        self.buf, self.hdl, self.sizes = buf, hdl, sizes
        self.ptrs = ptrs
        self.off_tn = _up(self.FLAG_BYTES)
        self.off_grad = self.off_tn + sizes[0]
        self.off_rel = self.off_grad + sizes[1]

    def _transport(self) -> str:
        """"symm": torch symmetric memory (default on multi-GPU nodes); "ipc": this library's
        CUDA-IPC mapping — chosen by BESS_PEER_TRANSPORT, or automatically when two ranks share
        a device (symmetric memory refuses that) or the process group has no NCCL backend."""
        if getattr(self, "_transport_name", None) is None:
            name = os.environ.get("BESS_PEER_TRANSPORT", "auto")
            if name == "auto":
                ids: List[Any] = [None] * self.n
                me = (os.uname().nodename, os.environ.get("CUDA_VISIBLE_DEVICES", ""),
                      torch.cuda.current_device())
                torch.distributed.all_gather_object(ids, me)
                backend = str(torch.distributed.get_backend())
                name = "ipc" if len(set(ids)) < self.n or "nccl" not in backend else "symm"
            if name not in ("symm", "ipc"):
                raise ValueError(f"BESS_PEER_TRANSPORT={name!r}: expected auto, symm or ipc")
            self._transport_name = name
        return self._transport_name

    def view(self, offset: int, shape: Sequence[int], dtype: torch.dtype) -> torch.Tensor:
        nbytes = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        return self.buf[offset:offset + nbytes].view(dtype).view(*shape)

    def flags(self, channel: int) -> torch.Tensor:
        return self.view(channel * 64 * 4, (self.n,), torch.int32)

    def signal(self, channel: int) -> None:
        """tell every rank that this rank's stores of the current phase are complete."""
        K.peer_signal(self.counters[channel:channel + 1],
                      [p + channel * 64 * 4 for p in self.ptrs], self.rank)

    def wait(self, channel: int) -> None:
        """wait until every rank has signalled the current phase (after `signal`)."""
        K.peer_wait(self.counters[channel:channel + 1], self.flags(channel), self.n)

    def handshake(self, channel: int) -> None:
        """signal every rank, then wait for every rank, on one channel."""
        self.signal(channel)
        self.wait(channel)


class _Staged:
    """Device-resident inputs of one call (see EmbeddingMovingBessKGE.stage)."""

    dims: Tuple[int, int, int, int, int]
    gidx: torch.Tensor
    rel: torch.Tensor
    tw: Optional[torch.Tensor]
    nmask: Optional[torch.Tensor]
    tmask: Optional[torch.Tensor]
    h2d_bytes: int


@dataclasses.dataclass
class _Pass:
    """One launch group of score_heads / score_tails (bess.py:368-466)."""

    mode: int
    n_query: int
    qmap: L.RowMap  # query q -> position in the [S] micro-batch order
    fixed_from_head: bool  # fixed entity rows come from H (else from received tails)
    fixed_map: L.RowMap
    shared: bool
    n_cand: int  # shared: size of the candidate list; else candidates per query
    cand_map: L.RowMap
    q_stride: int
    col0: int
    cand_from_head: bool = False  # candidates live in the local head buffer
    aug: bool = False  # candidates are the micro-batch's own heads / tails


# Shared-negative DistMult / ComplEx contractions run on the tcgen05 GEMM
# (csrc/gemm_tc.cu).  The exact-fp32 CUDA-core tile kernel (csrc/pair.cu)
# computes the same thing and is kept selectable for A/B parity tests only.
USE_TENSOR_CORES = True
# profiling aid: %globaltimer stamps between the stages of a (captured) training step
STAGE_STAMPS = os.environ.get("BESS_STAGE_STAMPS", "0") == "1"
STAGE_NAMES = ["start", "gather+local prologue", "rows arrived", "scores", "loss",
               "bwd: 1st contraction", "bwd: 2nd contraction", "backward",
               "gradients exchanged", "updated"]
SM_STAGE_NAMES = ["start", "query rows stored", "query rows arrived", "scored", "scores arrived",
                  "masks+metrics"]  # ScoreMoving distributed inference over peer memory
# fp32 tables on the tensor cores: 3xFP16 (scaled fp16 hi / lo pairs, csrc/gemm_tc.cu) instead of
# 3xTF32 for the DistMult / ComplEx training contractions.  BESS_F16X3=0 selects 3xTF32.
USE_F16X3 = os.environ.get("BESS_F16X3", "1") != "0"
# distributed training: the gather of the rows that travel to other GPUs runs on a second
# stream, next to the gather of the local head rows and the query prologue (A/B: BESS_SPLIT_GATHER=0)
SPLIT_EXCHANGE_GATHER = os.environ.get("BESS_SPLIT_GATHER", "1") != "0"
# ... and the blocks cross NVLink on the copy engines instead of by SM stores: the gather writes
# a local send buffer, one memcpy per destination (spread over COPY_STREAMS streams) delivers it,
# and the gradient blocks travel back the same way — no SM cycles, so the contractions that run
# next to the transfer keep their speed (A/B: BESS_COPY_ENGINE=0 selects the remote-store kernels)
# Measured (cfg 2, 16.5 MB per rank and direction; ms per step, copy engines vs remote stores):
# 2 ranks 0.394 vs 0.409, 4 ranks 0.395 vs 0.407, 8 ranks 0.438 vs 0.420 — seven 2.4 MB copies
# take ~90 us (~180 GB/s aggregate) and are no longer hidden under the contractions, while the
# remote-store kernels deliver the same bytes in ~35 us at the price of ~20 us of contention.
# "auto" picks by the number of ranks.
COPY_ENGINE = os.environ.get("BESS_COPY_ENGINE", "auto")
COPY_ENGINE_MAX_RANKS = int(os.environ.get("BESS_COPY_ENGINE_MAX_RANKS", "4"))
COPY_STREAMS = max(1, int(os.environ.get("BESS_COPY_STREAMS", "4")))


def _use_copy_engine(n_rank: int, overlapped: bool) -> bool:
    """`overlapped`: the step has contractions to hide the copies under (the tensor-core step
    that pushes its gradients early).  Where the transfer is exposed — e.g. the CUDA-core
    distance families, whose push follows the whole backward — the remote-store kernels are the
    faster way to move the same bytes (wikikg2 TransE-L1 at 2 ranks: 0.40 vs 0.42 ms)."""
    if COPY_ENGINE == "auto":
        return overlapped and n_rank <= COPY_ENGINE_MAX_RANKS
    return COPY_ENGINE != "0"
# score_triple fwd / bwd and the relation-table reduce on a second stream, concurrent with the
# negative-scoring / contraction kernels (A/B switch: BESS_OVERLAP=0)
OVERLAP_SMALL_KERNELS = os.environ.get("BESS_OVERLAP", "1") != "0"
# TransE / RotatE with scoring_norm=2 against shared negatives: norm-expanded distance
# ||q||^2 + ||c||^2 - 2 q.c with the q.c block (and both backward contractions) on the tcgen05
# GEMM (csrc/l2.cu).  The expansion carries an absolute error of ~1e-6 (||q||^2 + ||c||^2) in the
# squared distance — the same cancellation torch.cdist's own matmul path has — which only matters
# for pairs closer than ~1e-2 of their norms; False selects the exact CUDA-core tile kernels.
USE_L2_TENSOR_CORES = os.environ.get("BESS_L2_TC", "1") != "0"
# ... and only where it pays: the expansion adds ~a dozen small kernels (operand split, norms,
# coefficient sums, row scalings) around three GEMMs, which beats the three CUDA-core tile kernels
# once a pass has about this many (query, candidate, coordinate) elements (measured: 8192 x 64 x
# 128 = 6.7e7 is 17 % slower on the GEMM path, 0.202 vs 0.172 ms per step).
L2_TC_MIN_WORK = int(os.environ.get("BESS_L2_TC_MIN_WORK", str(1 << 27)))


def _pad8(x: int) -> int:
    return (x + 7) // 8 * 8


def _operand_format(table_dtype: torch.dtype, allow_f16x3: bool = True) -> int:
    """GEMM operand format of a table dtype: halves as they are; fp32 as scaled fp16 (hi, lo)
    pairs (3xFP16, twice the tensor-core rate of 3xTF32 at the same 22-bit products) or, when
    that is switched off, tf32 pairs."""
    if table_dtype != torch.float32:
        return L.dtype_code(table_dtype)
    return L.F16X3 if (USE_F16X3 and allow_f16x3) else L.F32


class _TcOperand:
    """Dense K-major operand arrays of one matrix for bess_dot_gemm: hi (+ lo for the split
    formats 3xTF32 / 3xFP16) [rows, ld] and, when `transpose`, hiT (+ loT) [width, ldt].
    `fmt` is the operand format (L.F32 = tf32 pairs, L.F16X3 = scaled fp16 pairs with the
    {s, 1 / s} pair in `scale`, L.F16 / L.BF16 = plain halves)."""

    def __init__(self, ws: "K.Workspace", tag: str, n_rows: int, width: int,
                 dtype: torch.dtype, transpose: bool, fmt: Optional[int] = None) -> None:
        self.fmt = L.dtype_code(dtype) if fmt is None else fmt
        two = self.fmt in (L.F32, L.F16X3)
        adt = torch.float16 if self.fmt == L.F16X3 else dtype
        tag = f"{tag}_x3" if self.fmt == L.F16X3 else tag
        self.n_rows, self.width = n_rows, width
        self.ld, self.ldt = _pad8(width), _pad8(n_rows)
        self.hi = ws.get(f"tc_{tag}_hi", (n_rows, self.ld), adt)
        self.lo = ws.get(f"tc_{tag}_lo", (n_rows, self.ld), adt) if two else None
        self.hit = self.lot = None
        if transpose:
            self.hit = ws.get(f"tc_{tag}_hiT", (width, self.ldt), adt)
            self.lot = ws.get(f"tc_{tag}_loT", (width, self.ldt), adt) if two else None
        self.scale = self.state = None
        if self.fmt == L.F16X3:
            self.scale = ws.get(f"tc_{tag}_scale", (2,), torch.float32)
            self.state = ws.get_zeroed(f"tc_{tag}_state", (2,), torch.int32)

    def fill(self, src_dt: int, src: L.Rows, out_dt: int, scale: Optional[torch.Tensor],
             device: torch.device) -> None:
        """`out_dt` is kept for call-site symmetry; the operand's own format decides."""
        if self.fmt == L.F16X3:
            K.operand_scale(src_dt, src, self.n_rows, self.width, scale, 1.0, self.scale,
                            self.state)
        K.split_operand(src_dt, src, self.n_rows, self.width, scale, self.fmt, self.hi, self.lo,
                        self.ld, self.hit, self.lot, self.ldt, device, self.scale)


def _as_i32(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.int32 else t.to(torch.int32)


class BessKGE(torch.nn.Module, ABC):
    """Base class (reference: bess.py:34-305): validation, input staging, masks,
    loss, metrics.  Subclasses define how negatives are scored."""

    def __init__(
        self,
        negative_sampler: ShardedNegativeSampler,
        score_fn: BaseScoreFunction,
        loss_fn: Optional[BaseLossFunction] = None,
        evaluation: Optional[Evaluation] = None,
        return_scores: bool = False,
        augment_negative: bool = False,
    ) -> None:
        super().__init__()
        self.sharding = score_fn.sharding
        self.negative_sampler = negative_sampler
        self.score_fn = score_fn
        self.loss_fn = loss_fn
        self.evaluation = evaluation
        self.return_scores = return_scores
        self.augment_negative = augment_negative
        if not (loss_fn or evaluation or return_scores):
            raise ValueError(
                "Nothing to return. At least one of loss_fn,"
                " evaluation or return_scores needs to be != None"
            )
        if self.augment_negative:
            assert (
                score_fn.negative_sample_sharing
            ), "Negative augmentation requires negative sample sharing"
            assert not isinstance(
                self, ScoreMovingBessKGE
            ), "ScoreMovingBessKGE does not support negative augmentation"
        if negative_sampler.flat_negative_format:
            assert (
                score_fn.negative_sample_sharing
            ), "Using flat negative format requires negative sample sharing"
        elif score_fn.negative_sample_sharing and isinstance(
            self.negative_sampler, TripleBasedShardedNegativeSampler
        ):
            raise ValueError(
                "Negative sample sharing cannot be used with non-flat triple-specific negatives"
            )
        self.entity_embedding = self.score_fn.entity_embedding
        self.entity_embedding_size: int = self.score_fn.entity_embedding.shape[-1]
        self._ws: Optional[K.Workspace] = None
        self._placement: Optional[_Placement] = None
        self._px: Optional[_PeerExchange] = None
        self._opt_state: Dict[str, Any] = {}
        # gradient accumulation over micro-batches (set by TrainingModel): an optimizer step is
        # applied on every k-th micro-batch, `_accum_phase0` = position of the current call's
        # first micro-batch in its cycle
        self._accum_k = 1
        self._accum_scale = 1.0
        self._accum_phase0 = 0

    # ------------------------------------------------------------------ misc
    @property
    def n_embedding_parameters(self) -> int:
        return self.score_fn.entity_embedding.numel() + self.score_fn.relation_embedding.numel()

    def _device(self) -> torch.device:
        dev = self.score_fn.entity_embedding.device
        if dev.type != "cuda":
            raise L.BessLibraryError(
                "besskge_b200 has no CPU path: move the module to a CUDA device first "
                "(model.cuda()); on a machine without a GPU only the host-side samplers run"
            )
        return dev

    def _tables(self) -> Tuple[torch.Tensor, torch.Tensor]:
        ent = self.score_fn.entity_embedding.data
        rel = self.score_fn.relation_embedding.data
        if rel.dtype != ent.dtype:
            raise TypeError("entity and relation tables must share one dtype")
        return ent, rel

    def _setup(self) -> Tuple[K.Workspace, _Placement]:
        dev = self._device()
        if self._ws is None or self._ws.device != dev:
            self._ws = K.Workspace(dev)
        if self._placement is None:
            self._placement = _Placement(self.sharding.n_shard)
        return self._ws, self._placement

    def _hyper(self, bps: int) -> torch.Tensor:
        """[bps, HYPER_COUNT] fp32 on the device: optimizer hyper-parameters of each
        micro-batch of the current call (csrc/scatter.cu reads them, so they are not frozen
        into a captured graph)."""
        return self._ws.get("opt_hyper", (bps, L.HYPER_COUNT), torch.float32)

    def _begin_steps(self, opt, bps: int) -> None:
        """Advance the step counter by `bps` optimizer steps and write their
        hyper-parameters (current lr / momentum / ..., Adam bias corrections of each step)
        to the device array on the current stream — OUTSIDE any captured graph."""
        self._setup()
        hyper = self._hyper(bps)
        b1, b2 = getattr(opt, "betas", (0.0, 0.0))
        k = self._accum_k
        self._accum_phase0 = self._opt_state.get("micro", 0) % k
        for s in range(bps):
            if (self._accum_phase0 + s) % k != k - 1:
                continue  # gradients of this micro-batch are only accumulated
            self._opt_state["step"] = self._opt_state.get("step", 0) + 1
            K.set_hyper(hyper[s], opt.lr, opt.momentum, opt.dampening, b1, b2,
                        getattr(opt, "eps", 0.0), opt.weight_decay, self._opt_state["step"])
        self._opt_state["micro"] = self._opt_state.get("micro", 0) + bps

    def _side_stream(self, dev: torch.device) -> "torch.cuda.Stream":
        if getattr(self, "_side", None) is None or self._side.device != dev:
            self._side = torch.cuda.Stream(dev)
        return self._side

    def _aux_stream(self, dev: torch.device) -> "torch.cuda.Stream":
        """Second side stream: small HBM-bound kernels that are independent of the big
        scoring / contraction kernels (score_triple forward and backward, the relation-table
        reduce) run here, concurrently with them (fork / join is part of the captured graph)."""
        if getattr(self, "_aux", None) is None or self._aux.device != dev:
            self._aux = torch.cuda.Stream(dev)
        return self._aux

    def _copy_streams(self, dev: torch.device) -> List["torch.cuda.Stream"]:
        sts = getattr(self, "_ce_streams", None)
        if sts is None or sts[0].device != dev or len(sts) != COPY_STREAMS:
            sts = self._ce_streams = [torch.cuda.Stream(dev) for _ in range(COPY_STREAMS)]
        return sts

    def _ce_blocks(self, dev: torch.device, src_ptr: int, stride: int, dst_ptrs: Sequence[int],
                   nbytes: int, rank: int) -> None:
        """dst_ptrs[j] <- src_ptr + j * stride (nbytes each) on the copy engines: forked from the
        current stream over the copy streams and joined back into it (under capture: parallel
        memcpy nodes).  Destinations are visited starting after this rank, so that at any moment
        the ranks write to different peers."""
        cur = torch.cuda.current_stream(dev)
        n = len(dst_ptrs)
        sts = self._copy_streams(dev)[:max(1, min(COPY_STREAMS, n))]
        for st in sts:
            st.wait_stream(cur)
        for t in range(n):
            j = (rank + 1 + t) % n
            if dst_ptrs[j]:  # 0: nothing to deliver to this rank (block already in place)
                K.peer_copy(src_ptr + j * stride, dst_ptrs[j], nbytes, sts[t % len(sts)])
        for st in sts:
            cur.wait_stream(st)

    # -------------------------------------------------------------- forward
    def forward(
        self,
        head: torch.Tensor,
        relation: torch.Tensor,
        tail: torch.Tensor,
        negative: torch.Tensor,
        triple_mask: Optional[torch.Tensor] = None,
        triple_weight: Optional[torch.Tensor] = None,
        negative_mask: Optional[torch.Tensor] = None,
    ) -> Dict[str, Any]:
        """Score (and, under `training_model`, train on) `batches_per_step`
        micro-batches.  Shapes as the reference (bess.py:117-161) with a leading
        `bps * n_shard` axis: head/relation/tail [L, n, p], negative [L, n, B, Nn],
        triple_weight [L, S], negative_mask [L, B, n, Nn], triple_mask [L, n, p]."""
        return self._run(head, relation, tail, negative, triple_mask, triple_weight,
                         negative_mask, optimizer=None)

    @abstractmethod
    def _run(self, head, relation, tail, negative, triple_mask, triple_weight, negative_mask,
             optimizer) -> Dict[str, Any]:
        raise NotImplementedError

    # ------------------------------------------------------ shared plumbing
    def _stage(self, name: str, t: Optional[torch.Tensor], dtype: torch.dtype,
               dev: torch.device) -> Optional[torch.Tensor]:
        """Host->device copy into a persistent buffer (no per-step allocation)."""
        if t is None:
            return None
        if t.dtype != dtype:
            t = t.to(dtype)
        buf = self._ws.get("in_" + name, tuple(t.shape), dtype)
        buf.copy_(t, non_blocking=True)
        return buf

    _PIN_SLOTS = 4

    def _pinned(self, name: str, shape: Tuple[int, ...], dtype: torch.dtype) -> torch.Tensor:
        """Next slot of the pinned staging ring of `name` (host side of an async H2D copy).
        A slot is reused only after the copy that last read it has completed."""
        ring = self.__dict__.setdefault("_pin_rings", {})
        ent = ring.get(name)
        n = int(np.prod(shape)) if len(shape) else 1
        if ent is None or ent["dtype"] != dtype or ent["numel"] < n:
            ent = ring[name] = dict(dtype=dtype, numel=max(n, 1), pos=-1, used=False,
                                    bufs=[torch.empty(max(n, 1), dtype=dtype, pin_memory=True)
                                          for _ in range(self._PIN_SLOTS)],
                                    events=[None] * self._PIN_SLOTS)
        ent["pos"] = (ent["pos"] + 1) % self._PIN_SLOTS
        ev = ent["events"][ent["pos"]]
        if ev is not None:
            ev.synchronize()
        ent["used"] = True
        return ent["bufs"][ent["pos"]][:n].view(*shape)

    def _pinned_done(self, name: str) -> None:
        """Record, on the current stream, that the H2D copy out of the current slot of `name`
        has been enqueued."""
        ent = self.__dict__.get("_pin_rings", {}).get(name)
        if ent is None or not ent["used"]:
            return
        ev = ent["events"][ent["pos"]]
        if ev is None:
            ev = ent["events"][ent["pos"]] = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self._ws.device))
        ent["used"] = False

    def _step_rows(self, placement: _Placement, bps: int) -> List[List[int]]:
        """rows of the STAGED inputs used at each step, one per local replica
        (distributed mode stages only this rank's rows)."""
        n = placement.n_shard
        if placement.distributed:
            return [[s] for s in range(bps)]
        return [[s * n + r for r in placement.shards] for s in range(bps)]

    def _finish_metrics(self, out: Dict[str, Any], pos: torch.Tensor, neg: torch.Tensor,
                        triple_mask: Optional[torch.Tensor], acc: Dict[str, List]) -> None:
        if self.evaluation is None:
            return
        rank = self.evaluation.ranks_from_scores(pos, neg)
        if self.evaluation.return_ranks:
            acc.setdefault("ranks", []).append(rank)
        acc.setdefault("metrics", []).append(
            self.evaluation.stacked_metrics_from_ranks(rank, triple_mask)
        )


# ---------------------------------------------------------------------------
# EmbeddingMoving
# ---------------------------------------------------------------------------
class EmbeddingMovingBessKGE(BessKGE):
    """Negative (and tail) embeddings travel to the shard that scores the
    positive triple (reference: bess.py:308-468).  One balanced AllToAll of
    `[n, p + B*Nn, D]` per micro-batch; in local mode it is folded into the
    gather kernel."""

    # ---- plan --------------------------------------------------------------
    def _plan(self, n: int, p: int, B: int, Nn: int) -> Tuple[List[_Pass], int]:
        scheme = self.negative_sampler.corruption_scheme
        flat = self.negative_sampler.flat_negative_format
        shared = bool(self.score_fn.negative_sample_sharing)
        local = bool(self.negative_sampler.local_sampling)
        S = n * p
        per = p if local else p + B * Nn
        # where negative block (j, b, k) lives: received buffer TN[j][p + b*Nn + k] or,
        # with local sampling, the local buffer after the S head rows
        if local:
            neg_off, neg_stride = S, B * Nn
        else:
            neg_off, neg_stride = p, per
        rm = L.rowmap
        passes: List[_Pass] = []
        if scheme in ("h", "t"):
            mode = L.MODE_TAILS if scheme == "t" else L.MODE_HEADS
            from_head = scheme == "t"
            fixed_map = L.IDENT if from_head else rm(p, per, 0)
            if flat:
                passes.append(_Pass(mode, S, L.IDENT, from_head, fixed_map, True, n * Nn,
                                    rm(Nn, neg_stride, neg_off), 0, 0))
                n_col = n * Nn
            elif shared:
                passes.append(_Pass(mode, S, L.IDENT, from_head, fixed_map, True, S * n * Nn,
                                    rm(Nn, neg_stride, neg_off, n * Nn, Nn), 0, 0))
                n_col = S * n * Nn
            else:
                passes.append(_Pass(mode, S, L.IDENT, from_head, fixed_map, False, n * Nn,
                                    rm(Nn, neg_stride, neg_off), Nn, 0))
                n_col = n * Nn
        elif scheme == "ht":
            half = p // 2
            for k, (mode, from_head) in enumerate(((L.MODE_HEADS, False), (L.MODE_TAILS, True))):
                qmap = rm(half, p, k * half)
                fixed_map = rm(half, p, k * half) if from_head else rm(half, per, k * half)
                if not flat and shared:
                    # every query of the group against the negatives of ALL its queries
                    # (bess.py:423-429 + the shared broadcast of scoring.py): candidate
                    # ((i, q'), j, kk) is the kk-th negative from shard j of query b = i*p + k*half
                    # + q'.  One pass per partition i keeps the candidate map two-level.
                    blk = half * n * Nn
                    for i in range(n):
                        cmap = rm(Nn, neg_stride, neg_off + (i * p + k * half) * Nn, n * Nn, Nn)
                        passes.append(_Pass(mode, n * half, qmap, from_head, fixed_map, True, blk,
                                            cmap, 0, i * blk))
                    continue
                if flat:
                    passes.append(_Pass(mode, n * half, qmap, from_head, fixed_map, True, n * Nn,
                                        rm(Nn, neg_stride, neg_off + k * Nn), 0, 0))
                else:
                    passes.append(_Pass(mode, n * half, qmap, from_head, fixed_map, False, n * Nn,
                                        rm(Nn, neg_stride, neg_off), Nn, 0))
            n_col = n * half * n * Nn if (not flat and shared) else n * Nn
        else:
            raise ValueError(f"unknown corruption scheme {scheme}")
        if self.augment_negative:
            if not flat or local:
                raise NotImplementedError(
                    "augment_negative needs flat_negative_format=True and local_sampling=False"
                )
            # the micro-batch's own heads / tails come first (bess.py:369-394, 430-448)
            aug: List[_Pass] = []
            for ps in passes:
                g = ps.qmap.group
                if ps.mode == L.MODE_TAILS:  # candidates: received tails of the same queries
                    cmap = rm(g, per, ps.qmap.offset) if g > 0 else rm(p, per, 0)
                    aug.append(_Pass(ps.mode, ps.n_query, ps.qmap, ps.fixed_from_head,
                                     ps.fixed_map, True, ps.n_query, cmap, 0, 0, False, True))
                else:  # candidates: local heads of the same queries
                    aug.append(_Pass(ps.mode, ps.n_query, ps.qmap, ps.fixed_from_head,
                                     ps.fixed_map, True, ps.n_query, ps.qmap, 0, 0, True, True))
                ps.col0 = ps.n_query
            n_col = passes[0].n_query + n * Nn
            passes = aug + passes
        return passes, n_col

    # ---- run ---------------------------------------------------------------
    def stage(self, head, relation, tail, negative, triple_mask=None, triple_weight=None,
              negative_mask=None, persistent: bool = False) -> "_Staged":
        """Pack the step's index tensors and copy them to the device (one H2D copy
        per tensor from the caller's — ideally pinned — host memory).  With
        `persistent=True` the device copies are owned by the returned object and
        can be replayed any number of times (inputs resident in HBM)."""
        ws, pl = self._setup()
        dev = ws.device
        n = self.sharding.n_shard
        L_rows = head.shape[0]
        if L_rows % n != 0:
            raise ValueError(f"leading axis {L_rows} is not a multiple of n_shard={n}")
        p = head.shape[-1]
        B, Nn = negative.shape[-2], negative.shape[-1]
        S = n * p
        local = bool(self.negative_sampler.local_sampling)
        # distributed mode: this rank needs (and packs, and copies) only its own rows
        mine = slice(pl.rank, None, n) if pl.distributed else slice(None)
        h2 = _as_i32(head).reshape(L_rows, S)[mine]
        t3 = _as_i32(tail).reshape(L_rows, n, p)[mine]
        n3 = _as_i32(negative).reshape(L_rows, n, B * Nn)[mine]
        rows = h2.shape[0]
        G = S + n * (p + B * Nn)
        # the packed gather list is assembled directly in pinned memory (one slot of a small
        # ring, so the asynchronous H2D copy of a previous call is never overwritten)
        gidx_h = self._pinned("gidx", (rows, G), torch.int32)
        gidx_h[:, :S] = h2
        if local:
            gidx_h[:, S:S + n * B * Nn] = n3.reshape(rows, -1)
            gidx_h[:, S + n * B * Nn:] = t3.reshape(rows, -1)
        else:
            body = gidx_h[:, S:].view(rows, n, p + B * Nn)
            body[:, :, :p] = t3
            body[:, :, p:] = n3

        def put(name, t, dtype, sliced=False):
            if t is None:
                return None
            if not sliced:
                t = t[mine]
            if persistent:
                return t.to(device=dev, dtype=dtype, non_blocking=True).contiguous()
            if not t.is_pinned() or t.dtype != dtype or not t.is_contiguous():
                host = self._pinned(name, tuple(t.shape), dtype)
                host.copy_(t)
                t = host
            return self._stage(name, t, dtype, dev)

        st = _Staged()
        st.dims = (L_rows, n, p, B, Nn)
        st.gidx = put("gidx", gidx_h, torch.int32, sliced=True)
        self._pinned_done("gidx")
        st.rel = put("rel", relation.reshape(L_rows, S), torch.int32)
        st.tw = put("tw", None if triple_weight is None else triple_weight.reshape(L_rows, -1),
                    torch.float32)
        st.nmask = None
        if negative_mask is not None:
            st.nmask = put("nmask", negative_mask.reshape(L_rows, negative_mask.shape[1], -1),
                           torch.bool if negative_mask.dtype == torch.bool else torch.uint8)
        st.tmask = put("tmask", None if triple_mask is None else triple_mask.reshape(L_rows, -1),
                       torch.bool)
        for name in ("rel", "tw", "nmask", "tmask"):
            self._pinned_done(name)
        st.h2d_bytes = sum(t.numel() * t.element_size()
                           for t in (st.gidx, st.rel, st.tw, st.nmask, st.tmask) if t is not None)
        return st

    def _restage(self, staged: "_Staged") -> "_Staged":
        """Device-to-device copy of staged inputs into the persistent input
        buffers (fixed addresses: what a captured CUDA graph reads)."""
        ws, _ = self._setup()
        st = _Staged()
        st.dims, st.h2d_bytes = staged.dims, staged.h2d_bytes
        for name in ("gidx", "rel", "tw", "nmask", "tmask"):
            t = getattr(staged, name)
            if t is not None:
                buf = ws.get("in_" + name, tuple(t.shape), t.dtype)
                if buf.data_ptr() != t.data_ptr():
                    buf.copy_(t, non_blocking=True)
                t = buf
            setattr(st, name, t)
        return st

    def _run(self, head, relation, tail, negative, triple_mask, triple_weight, negative_mask,
             optimizer, staged: Optional["_Staged"] = None) -> Dict[str, Any]:
        ws, pl = self._setup()
        dev = ws.device
        ent, rel_table = self._tables()
        n = self.sharding.n_shard
        if staged is None:
            staged = self.stage(head, relation, tail, negative, triple_mask, triple_weight,
                                negative_mask)
        L_rows, _, p, B, Nn = staged.dims
        bps = L_rows // n
        S = n * p
        local = bool(self.negative_sampler.local_sampling)
        per = p if local else p + B * Nn
        n_loc_rows = S + (n * B * Nn if local else 0)
        passes, N = self._plan(n, p, B, Nn)
        W, Wr = ent.shape[-1], rel_table.shape[-1]
        cfg = self.score_fn.kernel_cfg()
        dt = L.dtype_code(ent.dtype)
        tdt = ent.dtype
        train = optimizer is not None
        if train and self.loss_fn is None:
            raise ValueError("training needs a loss_fn")

        # ---- inputs: packed gather list gidx = [heads | (negs if local) | per dst: tails, negs]
        if staged is None:
            staged = self.stage(head, relation, tail, negative, triple_mask, triple_weight,
                                negative_mask)
        gidx, rel, tw, nmask, tmask = (staged.gidx, staged.rel, staged.tw, staged.nmask,
                                       staged.tmask)
        one = ws.get("one", (1,), torch.float32)
        one.fill_(1.0)

        R = pl.n_local
        G = gidx.shape[1]
        # ---- persistent buffers
        H = ws.get("H", (R, n_loc_rows, W), tdt)
        px: Optional[_PeerExchange] = None
        if pl.distributed and USE_PEER_EXCHANGE:
            if self._px is None:
                self._px = _PeerExchange(dev, n, pl.rank)
            px = self._px
            rel_cnt = _up(rel_table.numel() * 4, 16) // 4
            px.ensure(n * per * W * ent.element_size(), n * per * W * 4, n * rel_cnt * 4)
            TN = px.view(px.off_tn, (1, n, per, W), tdt)
        else:
            TN = ws.get("TN", (R, n, per, W), tdt)
        SEND = ws.get("SEND", (n, per, W), tdt) if pl.distributed and px is None else None
        pos_ws = ws.get("pos", (R, S), torch.float32)
        nvec = K.call("bess_query_nvec", L.C.byref(cfg))
        qv = ws.get("qv", (S, nvec, W), torch.float32)
        need_aux = cfg.family == L.BOXE and cfg.norm_p == 2
        need_scale = K.needs_cand_scale(cfg)
        aux = ws.get("aux", (R, S, N), torch.float32) if need_aux else None

        n_out = bps * R
        want_scores = self.return_scores
        pos_out = torch.empty(n_out * S, dtype=torch.float32, device=dev) if want_scores else None
        neg_out = (torch.empty(n_out * S, N, dtype=torch.float32, device=dev)
                   if want_scores else None)
        neg_ws = None if want_scores else ws.get("neg", (R, S, N), torch.float32)
        loss_out = (torch.empty(n_out, dtype=torch.float32, device=dev)
                    if self.loss_fn is not None else None)
        acc: Dict[str, List] = {}

        if self.loss_fn is not None:
            lp = self.loss_fn.kernel_params()
            row_loss = ws.get("row_loss", (S,), torch.float32)
            d_pos = ws.get("d_pos", (R, S), torch.float32)
            d_neg = ws.get("d_neg", (R, S, N), torch.float32)
            ce_copy = lp["kind"] == L.LOSS_SOFTMAX_CE and train and cfg.norm_p == 2 and \
                cfg.family in (L.TRANSE, L.ROTATE, L.PAIRRE, L.BOXE, L.TRIPLERE)
            neg_l = ws.get("neg_l", (S, N), torch.float32) if ce_copy else None
        if train:
            dH = ws.get("dH", (R, n_loc_rows, W), torch.float32)
            dTN = ws.get("dTN", (R, n, per, W), torch.float32)
            dBACK = None
            if px is not None:
                dBACK = px.view(px.off_grad, (n, per, W), torch.float32)
            elif pl.distributed:
                dBACK = ws.get("dBACK", (n, per, W), torch.float32)
            dRq = ws.get("dRq", (R, S, Wr), torch.float32)
            d_qv = ws.get("d_qv", (S, nvec, W), torch.float32)
            max_cand = max(ps.n_cand for ps in passes if ps.shared) if any(
                ps.shared for ps in passes) else 0
            cand_ws = None
            if max_cand:
                nbytes = max(K.shared_bwd_cand_workspace(cfg, ps.n_query, ps.n_cand)
                             for ps in passes if ps.shared)
                cand_ws = ws.get("cand_ws", (max(nbytes // 4, 1),), torch.float32)
            sk = ws.get("sort_keys", (R, G), torch.int32)
            sp = ws.get("sort_perm", (R, G), torch.int32)
            sort_ws = ws.get("sort_ws", (K.sort_workspace(max(G, R * S)) // 4 + 64,), torch.int32)
            rk = ws.get("rel_keys", (R * S,), torch.int32)
            rp = ws.get("rel_perm", (R * S,), torch.int32)
            d_rel_table = ws.get("d_rel_table", (_up(rel_table.numel() * 4, 16) // 4,),
                                 torch.float32)[:rel_table.numel()].view(rel_table.shape)
            key_bits = max(1, int(ent.shape[1] - 1).bit_length())
            rel_bits = max(1, int(rel_table.shape[0] - 1).bit_length())
        scale_buf = ws.get("cand_scale", (2 * max(ps.n_cand for ps in passes),), torch.float32) \
            if need_scale else None

        use_tc = USE_TENSOR_CORES and cfg.family in (L.DISTMULT, L.COMPLEX)
        use_l2tc = (USE_TENSOR_CORES and USE_L2_TENSOR_CORES and cfg.norm_p == 2
                    and cfg.family in (L.TRANSE, L.ROTATE))
        # per pass: tensor cores only where the pass is large enough to amortise the extra kernels
        l2tc = [bool(use_l2tc and ps.shared and ps.n_query * ps.n_cand * W >= L2_TC_MIN_WORK)
                for ps in passes]
        use_l2tc = any(l2tc)
        tc_q: Dict[int, _TcOperand] = {}
        tc_c: Dict[int, _TcOperand] = {}
        gemm_ws = None
        if (use_tc or use_l2tc) and any(ps.shared for ps in passes):
            nbytes = 0
            for ps in passes:
                if ps.shared:
                    nbytes = max(nbytes, K.dot_gemm_workspace(ps.n_query, ps.n_cand, W))
                    if train:
                        nbytes = max(nbytes, K.dot_gemm_workspace(ps.n_query, W, ps.n_cand),
                                     K.dot_gemm_workspace(ps.n_cand, W, ps.n_query))
            gemm_ws = ws.get("gemm_ws", (max(nbytes // 4, 1),), torch.float32)
        # The loss kernel can emit dL/dscore directly as GEMM operand arrays (no fp32 [S, N]
        # gradient, no split / transpose pass) when every pass is a tensor-core pass over
        # all S rows in micro-batch order and its column block is 16-byte aligned.
        ldN = _pad8(N)
        direct_ds = bool(
            use_tc and train and passes
            and all(ps.shared and ps.qmap.group <= 0 and ps.qmap.offset == 0 and ps.n_query == S
                    and ps.col0 % 8 == 0 for ps in passes))
        # Distributed 't'-type steps: the gradient rows that travel back (tails from
        # score_triple, negatives from the dC contractions) are complete before the dQ
        # contractions and the prologue backward run, so their push over NVLink and the
        # handshake go to the side stream and overlap that remaining compute.
        early_push = bool(px is not None and direct_ds and R == 1
                          and all(ps.fixed_from_head for ps in passes))
        use_ce = bool(px is not None and _use_copy_engine(n, early_push))
        # operand format of the DOT contractions (fp32 tables: scaled fp16 pairs, 3xFP16); the
        # norm-expanded L2 path keeps tf32 pairs
        gdt = _operand_format(tdt) if use_tc else dt
        ds_hi = ds_lo = ds_scale = ds_state = None
        if direct_ds:
            ds_dt = torch.float16 if gdt == L.F16X3 else tdt
            sfx = "_x3" if gdt == L.F16X3 else ""
            ds_hi = ws.get("tc_dsd_hi" + sfx, (S, ldN), ds_dt)
            ds_lo = (ws.get("tc_dsd_lo" + sfx, (S, ldN), ds_dt)
                     if gdt in (L.F32, L.F16X3) else None)
            if gdt == L.F16X3:
                ds_scale = ws.get("tc_dsd_scale", (2,), torch.float32)
                ds_state = ws.get_zeroed("tc_dsd_state", (2,), torch.int32)

        flat = self.negative_sampler.flat_negative_format
        scheme = self.negative_sampler.corruption_scheme

        # optional per-stage device timestamps (BESS_STAGE_STAMPS=1; bench --stage-timing)
        stamps = None
        if STAGE_STAMPS and train:
            stamps = ws.get_zeroed("stage_stamps", (bps, len(STAGE_NAMES)), torch.int64)

        def stamp(s_, name):
            if stamps is not None:
                K.stamp(stamps[s_, STAGE_NAMES.index(name)])

        side = self._side_stream(dev) if train else None
        # score_triple (forward and backward) only touches rows that the negative-scoring
        # kernels never write — unless augment_negative makes the micro-batch's own heads /
        # tails candidates as well — so it runs on a second stream next to them
        overlap = OVERLAP_SMALL_KERNELS and not self.augment_negative
        aux_st = self._aux_stream(dev) if overlap else None
        # The relation-table update of micro-batch s (reduce -> all-reduce over peer memory ->
        # optimizer, all on the side stream) is joined only where micro-batch s+1 first reads the
        # table, so its tail runs under the next gather instead of extending this step.
        rel_pending: List[Any] = []

        def join_relation_update():
            while rel_pending:  # an event, not the stream: the side stream has newer work queued
                torch.cuda.current_stream(dev).wait_event(rel_pending.pop())

        for s, step_rows in enumerate(self._step_rows(pl, bps)):
            if train:
                # The sort permutations of the scatter depend only on the step's indices:
                # compute them on a side stream while the main stream gathers and scores
                # (fork / join is captured into the CUDA graph like any other dependency).
                main = torch.cuda.current_stream(dev)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    for li, row in enumerate(step_rows):
                        K.sort_keys(gidx[row], G, key_bits, sk[li], sp[li], sort_ws)
                    rows_this_step = rel[step_rows[0]:step_rows[-1] + 1]
                    if not rows_this_step.is_contiguous() or len(step_rows) != R:
                        rows_this_step = rows_this_step.contiguous()
                    K.sort_keys(rows_this_step.view(-1), R * S, rel_bits, rk, rp, sort_ws)
            stamp(s, "start")
            # ================= gather (+ exchange) =================
            xch_st = None
            for li, (row, shard) in enumerate(zip(step_rows, pl.shards)):
                table = ent[shard]
                if px is not None:  # rows go straight into every peer's receive buffer
                    dst = [p + px.off_tn for p in px.ptrs]
                    slot = pl.rank
                elif pl.distributed:
                    dst = [SEND[j].data_ptr() for j in range(n)]
                    slot = 0
                else:
                    dst = [TN[j].data_ptr() for j in range(n)]
                    slot = li
                if px is not None and SPLIT_EXCHANGE_GATHER:
                    # The rows that travel (remote stores over NVLink: latency-bound, few SM
                    # cycles) go to a second stream together with their signal; the local head
                    # rows and everything that only needs them stay on this one.
                    main = torch.cuda.current_stream(dev)
                    xch_st = self._aux_stream(dev)
                    xch_st.wait_stream(main)
                    with torch.cuda.stream(xch_st):
                        if use_ce:
                            rb = per * W * ent.element_size()
                            sendx = ws.get("SENDX", (n, per, W), tdt)
                            stage_dst = [sendx[j].data_ptr() for j in range(n)]
                            stage_dst[pl.rank] = dst[pl.rank] + pl.rank * rb  # own block: in place
                            K.gather_route(table, gidx[row][n_loc_rows:], 0, per, None, stage_dst, 0)
                            self._ce_blocks(dev, sendx.data_ptr(), rb,
                                            [dst[j] + pl.rank * rb if j != pl.rank else 0
                                             for j in range(n)], rb, pl.rank)
                        else:
                            K.gather_route(table, gidx[row][n_loc_rows:], 0, per, None, dst, slot)
                        px.signal(0)
                    K.gather_route(table, gidx[row][:n_loc_rows], n_loc_rows, 0, H[li], [], 0)
                else:
                    K.gather_route(table, gidx[row], n_loc_rows, per, H[li], dst, slot)
            join_relation_update()
            pre_done = False
            if px is not None:
                # Signal at once, wait as late as possible: the query prologue and its operand
                # split read only LOCAL rows (own heads, replicated relation table), so they run
                # while the peers' tail / negative rows are still arriving over NVLink.
                if xch_st is None:
                    px.signal(0)
                if R == 1 and len(passes) == 1 and passes[0].fixed_from_head:
                    ps0, row0 = passes[0], step_rows[0]
                    K.prologue_fwd(cfg, dt, ps0.mode, L.rows(H[0], rmap=ps0.fixed_map), rel_table,
                                   rel[row0], ps0.qmap, ps0.n_query, qv)
                    if ps0.shared and (use_tc or l2tc[0]):
                        q0 = tc_q[0] = _TcOperand(ws, "q0", ps0.n_query, W, tdt, train, gdt)
                        q0.fill(L.F32, L.rows(qv.view(-1, W)), dt, None, dev)
                    pre_done = True
                stamp(s, "gather+local prologue")
                if xch_st is not None:
                    torch.cuda.current_stream(dev).wait_stream(xch_st)  # own block of TN stored
                px.wait(0)
            elif pl.distributed:
                stamp(s, "gather+local prologue")
                pl.all_to_all(TN[0], SEND)
            else:
                stamp(s, "gather+local prologue")
            stamp(s, "rows arrived")

            for li, row in enumerate(step_rows):
                o = s * R + li
                pos = pos_out[o * S:(o + 1) * S] if want_scores else pos_ws[li]
                neg = neg_out[o * S:(o + 1) * S] if want_scores else neg_ws[li]
                Hl, TNl = H[li], TN[li].view(n * per, W)
                head_rows = L.rows(Hl)
                tail_rows = L.rows(TNl, rmap=L.rowmap(p, per, 0))
                # ================= positive scores =================
                if overlap:
                    main = torch.cuda.current_stream(dev)
                    aux_st.wait_stream(main)
                    with torch.cuda.stream(aux_st):
                        K.triple_fwd(cfg, dt, head_rows, tail_rows, rel_table, rel[row], L.IDENT, S,
                                     pos, L.IDENT)
                else:
                    K.triple_fwd(cfg, dt, head_rows, tail_rows, rel_table, rel[row], L.IDENT, S,
                                 pos, L.IDENT)
                # ================= negative scores =================
                for pi, ps in enumerate(passes):
                    fixed = (L.rows(Hl, rmap=ps.fixed_map) if ps.fixed_from_head
                             else L.rows(TNl, rmap=ps.fixed_map))
                    if not pre_done:
                        K.prologue_fwd(cfg, dt, ps.mode, fixed, rel_table, rel[row], ps.qmap,
                                       ps.n_query, qv)
                    cand = self._cand_rows(ps, Hl, TNl, local)
                    if ps.shared and use_tc:
                        # scores = Q C^T on the tensor cores (scoring.py:252)
                        if pre_done:
                            q_op = tc_q[pi]
                        else:
                            q_op = tc_q[pi] = _TcOperand(ws, f"q{pi}", ps.n_query, W, tdt, train,
                                                         gdt)
                            q_op.fill(L.F32, L.rows(qv.view(-1, W)), dt, None, dev)
                        c_op = tc_c[pi] = _TcOperand(ws, f"c{pi}", ps.n_cand, W, tdt, train, gdt)
                        c_op.fill(dt, cand, dt, None, dev)
                        K.dot_gemm(gdt, q_op.hi, q_op.lo, q_op.ld, c_op.hi, c_op.lo, c_op.ld,
                                   ps.n_query, ps.n_cand, W, neg, ps.qmap, N, ps.col0, False,
                                   gemm_ws, a_scale=q_op.scale, b_scale=c_op.scale)
                    elif l2tc[pi]:
                        # -||q - c||_2 from ||q||^2 + ||c||^2 - 2 q.c, q.c on the tensor cores
                        if pre_done:
                            q_op = tc_q[pi]
                        else:
                            q_op = tc_q[pi] = _TcOperand(ws, f"q{pi}", ps.n_query, W, tdt, train)
                            q_op.fill(L.F32, L.rows(qv.view(-1, W)), dt, None, dev)
                        c_op = tc_c[pi] = _TcOperand(ws, f"c{pi}", ps.n_cand, W, tdt, train)
                        c_op.fill(dt, cand, dt, None, dev)
                        K.dot_gemm(dt, q_op.hi, q_op.lo, q_op.ld, c_op.hi, c_op.lo, c_op.ld,
                                   ps.n_query, ps.n_cand, W, neg, ps.qmap, N, ps.col0, False,
                                   gemm_ws)
                        qn = ws.get(f"l2_qn{pi}", (ps.n_query,), torch.float32)
                        cn = ws.get(f"l2_cn{pi}", (ps.n_cand,), torch.float32)
                        if tdt == torch.float32:  # 3xTF32 products are fp32-grade: norms of q itself
                            K.row_sqnorm(L.F32, L.rows(qv.view(-1, W)), ps.n_query, W, qn)
                        else:  # norms of the rounded operand the MMA multiplied
                            K.row_sqnorm(dt, L.rows(q_op.hi), ps.n_query, W, qn)
                        K.row_sqnorm(dt, cand, ps.n_cand, W, cn)
                        K.l2_from_dot(neg, ps.qmap, N, ps.col0, ps.n_query, ps.n_cand, qn, cn)
                    elif ps.shared:
                        scale = None
                        if need_scale:
                            scale = K.cand_scales(cfg, dt, cand, ps.n_cand, W, scale_buf)
                        K.shared_fwd(cfg, dt, ps.mode, qv, ps.n_query, cand, scale, ps.n_cand,
                                     neg, ps.qmap, N, ps.col0, None if aux is None else aux[li])
                    else:
                        K.pertriple_fwd(cfg, dt, ps.mode, qv, ps.n_query, cand, ps.q_stride,
                                        ps.n_cand, neg, ps.qmap, N, ps.col0,
                                        None if aux is None else aux[li])
                if overlap:
                    torch.cuda.current_stream(dev).wait_stream(aux_st)  # positive scores ready
                stamp(s, "scores")
                # ================= masks (bess.py:182-245) =================
                self._apply_masks(neg, S, N, p, B, Nn, n, nmask[row] if nmask is not None else None,
                                  flat, scheme)
                # ================= loss =================
                if self.loss_fn is not None:
                    w = tw[row] if tw is not None else one
                    neg_for_loss = neg
                    if ce_copy:
                        neg_l.copy_(neg)
                        neg_for_loss = neg_l
                    if direct_ds:
                        if gdt == L.F16X3:
                            # |dL/dscore| <= loss_scale * max(weight) for every loss here: the
                            # power-of-two scale of the gradient operand comes from that bound
                            K.operand_scale(L.F32, L.rows(w.view(1, -1)), 1, w.numel(), None,
                                            abs(float(lp["loss_scale"])), ds_scale, ds_state)
                        K.loss_fwd_bwd_operand(lp["kind"], lp["margin"], lp["adversarial"],
                                               lp["adv_scale"], lp["loss_scale"], lp["n_entity"],
                                               pos, neg_for_loss, S, N, N, w, row_loss, d_pos[li],
                                               gdt, ds_hi, ds_lo, ldN, ds_scale)
                    else:
                        K.loss_fwd_bwd(lp["kind"], lp["margin"], lp["adversarial"],
                                       lp["adv_scale"], lp["loss_scale"], lp["n_entity"], pos,
                                       neg_for_loss, S, N, N, w, row_loss, d_pos[li], d_neg[li])
                    K.sum_f32(row_loss, S, loss_out[o:o + 1])
                    if ce_copy and want_scores:
                        neg_saved = ws.get("neg_unshift", (R, S, N), torch.float32)
                        neg_saved[li].copy_(neg)
                        neg.copy_(neg_l)
                if self.evaluation is not None:
                    self._finish_metrics({}, pos, neg, tmask[row] if tmask is not None else None,
                                         acc)

                stamp(s, "loss")
                # ================= backward =================
                if train:
                    score_for_bwd = neg
                    if ce_copy and want_scores:
                        score_for_bwd = ws.get("neg_unshift", (R, S, N), torch.float32)[li]
                    dHl, dTNl = dH[li], dTN[li].view(n * per, W)
                    main = torch.cuda.current_stream(dev)
                    if overlap:
                        aux_st.wait_stream(main)
                    with torch.cuda.stream(aux_st if overlap else main):
                        K.triple_bwd(cfg, dt, head_rows, tail_rows, rel_table, rel[row], L.IDENT, S,
                                     pos, d_pos[li], L.IDENT, L.rows(dHl),
                                     L.rows(dTNl, rmap=L.rowmap(p, per, 0)), dRq[li], False, False,
                                     False)
                    tb_joined = not overlap

                    def join_triple_bwd():
                        # score_triple's gradients (dH, tail rows of dTN, dRq) are complete
                        nonlocal tb_joined
                        if not tb_joined:
                            torch.cuda.current_stream(dev).wait_stream(aux_st)
                            tb_joined = True
                    if early_push:
                        for pi, ps in enumerate(passes):
                            d_cand = self._cand_rows(ps, dHl, dTNl, local)
                            K.dot_gemm(gdt, ds_hi, ds_lo, ldN, tc_q[pi].hit, tc_q[pi].lot,
                                       tc_q[pi].ldt, ps.n_cand, W, ps.n_query, d_qv, d_cand.map,
                                       d_cand.pitch, 0, ps.aug, gemm_ws, out_ptr=d_cand.base,
                                       a_mn_major=True, a_offset_elems=ps.col0,
                                       a_scale=ds_scale, b_scale=tc_q[pi].scale)
                        join_triple_bwd()
                        stamp(s, "bwd: 1st contraction")  # dC (+ score_triple's backward)
                        main = torch.cuda.current_stream(dev)
                        side.wait_stream(main)
                        with torch.cuda.stream(side):
                            self._push_gradients(dev, px, pl, dTN, per * W * 4, use_ce)
                            px.handshake(1)
                    for pi, ps in enumerate(passes):
                        fixed = (L.rows(Hl, rmap=ps.fixed_map) if ps.fixed_from_head
                                 else L.rows(TNl, rmap=ps.fixed_map))
                        d_fixed = (L.rows(dHl, rmap=ps.fixed_map) if ps.fixed_from_head
                                   else L.rows(dTNl, rmap=ps.fixed_map))
                        cand = self._cand_rows(ps, Hl, TNl, local)
                        d_cand = self._cand_rows(ps, dHl, dTNl, local)
                        a = None if aux is None else aux[li]
                        if ps.shared and use_tc:
                            # dQ = dS C and dC = dS^T Q on the tensor cores; the operand
                            # arrays of Q and C (and their transposes) survive from forward
                            q_op, c_op = tc_q[pi], tc_c[pi]
                            if direct_ds:
                                # dS operand arrays came straight from the loss kernel; dC reads
                                # them MN-major (dS^T without a transposed copy)
                                K.dot_gemm(gdt, ds_hi, ds_lo, ldN, c_op.hit, c_op.lot, c_op.ldt,
                                           ps.n_query, W, ps.n_cand, d_qv, L.IDENT, W, 0, False,
                                           gemm_ws, a_offset_elems=ps.col0, a_scale=ds_scale,
                                           b_scale=c_op.scale)
                                stamp(s, "bwd: 2nd contraction" if early_push
                                      else "bwd: 1st contraction")  # dQ
                                if not early_push:
                                    K.dot_gemm(gdt, ds_hi, ds_lo, ldN, q_op.hit, q_op.lot, q_op.ldt,
                                               ps.n_cand, W, ps.n_query, d_qv, d_cand.map,
                                               d_cand.pitch, 0, ps.aug, gemm_ws,
                                               out_ptr=d_cand.base, a_mn_major=True,
                                               a_offset_elems=ps.col0, a_scale=ds_scale,
                                               b_scale=q_op.scale)
                                    stamp(s, "bwd: 2nd contraction")  # dC
                            else:
                                ds = _TcOperand(ws, "ds", ps.n_query, ps.n_cand, tdt, True, gdt)
                                ds.fill(L.F32, L.rows(d_neg[li], rmap=ps.qmap, pitch=N,
                                                      offset_elems=ps.col0), dt, None, dev)
                                K.dot_gemm(gdt, ds.hi, ds.lo, ds.ld, c_op.hit, c_op.lot, c_op.ldt,
                                           ps.n_query, W, ps.n_cand, d_qv, L.IDENT, W, 0, False,
                                           gemm_ws, a_scale=ds.scale, b_scale=c_op.scale)
                                K.dot_gemm(gdt, ds.hit, ds.lot, ds.ldt, q_op.hit, q_op.lot,
                                           q_op.ldt, ps.n_cand, W, ps.n_query, d_qv, d_cand.map,
                                           d_cand.pitch, 0, ps.aug, gemm_ws, out_ptr=d_cand.base,
                                           a_scale=ds.scale, b_scale=q_op.scale)
                            join_triple_bwd()
                            K.prologue_bwd(cfg, dt, ps.mode, fixed, rel_table, rel[row], ps.qmap,
                                           ps.n_query, d_qv, d_fixed, dRq[li], True, True)
                            continue
                        K.prologue_fwd(cfg, dt, ps.mode, fixed, rel_table, rel[row], ps.qmap,
                                       ps.n_query, qv)
                        if l2tc[pi]:
                            # b = dL/dscore / dist; dQ = B C - rb * Q, dC = B^T Q - cb * C
                            q_op, c_op = tc_q[pi], tc_c[pi]
                            nq_, nc_ = ps.n_query, ps.n_cand
                            ldc = _pad8(nc_)
                            coef = ws.get("l2_coef", (nq_, ldc), torch.float32)
                            rb = ws.get("l2_rb", (nq_,), torch.float32)
                            cb = ws.get("l2_cb", (nc_,), torch.float32)
                            cws = ws.get("l2_ws", (max(K.l2_coef_workspace(nq_, nc_) // 4, 1),),
                                         torch.float32)
                            K.l2_coef(d_neg[li], score_for_bwd, ps.qmap, N, ps.col0, nq_, nc_, coef,
                                      ldc, rb, cb, cws)
                            ds = _TcOperand(ws, "ds", nq_, nc_, tdt, True)
                            ds.fill(L.F32, L.rows(coef), dt, None, dev)
                            K.dot_gemm(dt, ds.hi, ds.lo, ds.ld, c_op.hit, c_op.lot, c_op.ldt, nq_, W,
                                       nc_, d_qv, L.IDENT, W, 0, False, gemm_ws)
                            K.rows_axpy(L.F32, rb, -1.0, L.rows(qv.view(-1, W)),
                                        L.rows(d_qv.view(-1, W)), nq_, W)
                            K.dot_gemm(dt, ds.hit, ds.lot, ds.ldt, q_op.hit, q_op.lot, q_op.ldt, nc_,
                                       W, nq_, d_qv, d_cand.map, d_cand.pitch, 0, ps.aug, gemm_ws,
                                       out_ptr=d_cand.base)
                            K.rows_axpy(dt, cb, -1.0, cand, d_cand, nc_, W)
                            join_triple_bwd()
                            K.prologue_bwd(cfg, dt, ps.mode, fixed, rel_table, rel[row], ps.qmap,
                                           nq_, d_qv, d_fixed, dRq[li], True, True)
                            continue
                        if ps.shared:
                            scale = None
                            if need_scale:
                                scale = K.cand_scales(cfg, dt, cand, ps.n_cand, W, scale_buf)
                            # dQ and dC are independent: the candidate-gradient kernel goes
                            # to the second stream (after score_triple's backward there)
                            cand_st = aux_st if (overlap and not need_scale) else main
                            if cand_st is not main:
                                cand_st.wait_stream(torch.cuda.current_stream(dev))
                            # augmented candidates are the micro-batch's own head / tail rows,
                            # whose gradient rows also collect other terms: with a candidate
                            # normalisation, this pass's term goes through the chain rule in a
                            # buffer of its own before it is added
                            via_tmp = bool(ps.aug and need_scale)
                            d_target = d_cand
                            if via_tmp:
                                d_cand = L.rows(ws.get("aug_dcand", (ps.n_cand, W), torch.float32))
                            with torch.cuda.stream(cand_st):
                                K.shared_bwd_cand(cfg, dt, ps.mode, qv, ps.n_query, cand, scale,
                                                  ps.n_cand, score_for_bwd, d_neg[li], ps.qmap, N,
                                                  ps.col0, a, d_cand, cand_ws,
                                                  add=ps.aug and not via_tmp)
                            K.shared_bwd_query(cfg, dt, ps.mode, qv, ps.n_query, cand, scale,
                                               ps.n_cand, score_for_bwd, d_neg[li], ps.qmap, N,
                                               ps.col0, a, d_qv)
                            if cand_st is not main:
                                tb_joined = False  # more work on the second stream: join again
                            if need_scale:
                                K.cand_scales_bwd(cfg, dt, cand, ps.n_cand, W, scale, d_cand)
                            if via_tmp:
                                ones = ws.get("ones_rows", (ps.n_cand,), torch.float32)
                                ones.fill_(1.0)
                                K.rows_axpy(L.F32, ones, 1.0, d_cand, d_target, ps.n_cand, W)
                        else:
                            K.pertriple_bwd(cfg, dt, ps.mode, qv, ps.n_query, cand, ps.q_stride,
                                            ps.n_cand, score_for_bwd, d_neg[li], ps.qmap, N,
                                            ps.col0, a, d_qv, d_cand)
                        join_triple_bwd()
                        K.prologue_bwd(cfg, dt, ps.mode, fixed, rel_table, rel[row], ps.qmap,
                                       ps.n_query, d_qv, d_fixed, dRq[li], True, True)
                    join_triple_bwd()
                    if cfg.family == L.BOXE:
                        K.boxe_rel_finalize(cfg, dt, rel_table, rel[row], S, dRq[li])

            stamp(s, "backward")
            # ================= reverse exchange + update =================
            if px is not None and not train:
                px.handshake(1)  # every rank is done reading its receive buffer
            if train:
                if px is not None and not early_push:
                    self._push_gradients(dev, px, pl, dTN, per * W * 4, use_ce)
                    px.handshake(1)
                elif pl.distributed and px is None:
                    pl.all_to_all(dBACK, dTN[0])
                hyper = self._hyper(bps)[s]  # filled by TrainingModel before this call / replay
                main = torch.cuda.current_stream(dev)
                main.wait_stream(side)  # join: permutations ready (and the early gradient push)
                stamp(s, "gradients exchanged")
                # relation table (replicated): reduce per-query rows of all local replicas, then
                # its all-reduce over peer memory — on the side stream, under the entity scatter
                mean = getattr(optimizer, "relation_grad_reduction", "mean") == "mean"
                rel_st = side if (px is not None or overlap) else main
                if rel_st is not main:
                    rel_st.wait_stream(main)
                with torch.cuda.stream(rel_st):
                    K.relation_grad_reduce(dRq.view(R * S, Wr), Wr, rk, rp, R * S,
                                           rel_table.shape[0], d_rel_table)
                    if px is not None:
                        # every rank pushes its partial into slot [rank] of every rank, then sums
                        # the n slots in rank order (bit-identical tables on all ranks)
                        cnt = _up(rel_table.numel() * 4, 16) // 4
                        K.peer_push(d_rel_table, 0,
                                    [q + px.off_rel + pl.rank * cnt * 4 for q in px.ptrs], cnt * 4)
                        px.handshake(2)
                        K.peer_reduce(px.view(px.off_rel, (n, cnt), torch.float32), n, cnt,
                                      1.0 / n if mean else 1.0, d_rel_table)
                    else:
                        if pl.distributed:
                            torch.distributed.all_reduce(d_rel_table)
                        if mean and n > 1:
                            d_rel_table.mul_(1.0 / n)
                acc_k = self._accum_k
                acc_phase = (self._accum_phase0 + s) % acc_k
                defer_rel = rel_st is not main and acc_k == 1
                if defer_rel:
                    with torch.cuda.stream(rel_st):
                        self._update_relation(optimizer, rel_table, d_rel_table, hyper, ws)
                        done = torch.cuda.Event()
                        done.record(rel_st)
                    rel_pending.append(done)
                for li, (row, shard) in enumerate(zip(step_rows, pl.shards)):
                    if pl.distributed:
                        g_dst, stride_rows = dBACK.data_ptr(), per
                    else:
                        g_dst = dTN.data_ptr() + li * per * W * 4
                        stride_rows = n * per
                    if acc_k > 1:
                        self._accumulate_entity(optimizer, ent[shard], shard, sk[li], sp[li], G,
                                                n_loc_rows, per, dH[li], g_dst, stride_rows, hyper,
                                                acc_phase == acc_k - 1)
                    else:
                        self._update_entity(optimizer, ent[shard], shard, sk[li], sp[li], G,
                                            n_loc_rows, per, dH[li], g_dst, stride_rows, hyper, ws)
                if defer_rel:
                    pass  # joined by the next micro-batch (or after the last one)
                elif rel_st is not main:
                    main.wait_stream(rel_st)  # relation gradient reduced (and all-reduced)
                if defer_rel:
                    pass
                elif acc_k > 1:
                    # relation table: one slot per micro-batch of the cycle, summed in slot
                    # order on the last one (deterministic), then one optimizer step
                    cnt = d_rel_table.numel()
                    slots = self._opt_state.get("rel_acc")
                    if slots is None or slots.shape != (acc_k, cnt):
                        slots = self._opt_state["rel_acc"] = torch.zeros(
                            acc_k, cnt, dtype=torch.float32, device=dev)
                    slots[acc_phase].copy_(d_rel_table.view(-1))
                    if acc_phase == acc_k - 1:
                        K.peer_reduce(slots, acc_k, cnt, self._accum_scale, d_rel_table)
                        self._update_relation(optimizer, rel_table, d_rel_table, hyper, ws)
                else:
                    self._update_relation(optimizer, rel_table, d_rel_table, hyper, ws)
                stamp(s, "updated")
        join_relation_update()

        out: Dict[str, Any] = {}
        if want_scores:
            out["positive_score"] = pos_out if tdt == torch.float32 else pos_out.to(tdt)
            out["negative_score"] = neg_out if tdt == torch.float32 else neg_out.to(tdt)
        if loss_out is not None:
            out["loss"] = loss_out
        if "ranks" in acc:
            out["ranks"] = torch.cat(acc["ranks"])
        if "metrics" in acc:
            out["metrics"] = torch.cat(acc["metrics"], dim=0)
        return out

    # ---- helpers -------------------------------------------------------------
    def _push_gradients(self, dev, px: _PeerExchange, pl: _Placement, dTN: torch.Tensor,
                        blk: int, use_ce: bool) -> None:
        """block j of this replica's fp32 gradient buffer -> slot [rank] of rank j's gradient
        receive buffer (the reverse of the forward exchange, bess.py:348-350 under autograd)."""
        dst = [q + px.off_grad + pl.rank * blk for q in px.ptrs]
        if use_ce:
            self._ce_blocks(dev, dTN[0].data_ptr(), blk, dst, blk, pl.rank)
        else:
            K.peer_push(dTN[0], blk, dst, blk)

    @staticmethod
    def _cand_rows(ps: _Pass, Hl: torch.Tensor, TNl: torch.Tensor, local: bool) -> L.Rows:
        if ps.cand_from_head or (local and not ps.aug):
            return L.rows(Hl, rmap=ps.cand_map)
        return L.rows(TNl, rmap=ps.cand_map)

    def _apply_masks(self, neg: torch.Tensor, S: int, N: int, p: int, B: int, Nn: int, n: int,
                     nmask: Optional[torch.Tensor], flat: bool, scheme: str) -> None:
        """BAD_NEGATIVE_SCORE on padding negatives and, with augment_negative, on
        each triple's own head/tail column (bess.py:182-245)."""
        if (nmask is not None and not flat
                and bool(self.score_fn.negative_sample_sharing)):
            raise ValueError("negative_mask cannot be combined with non-flat shared negatives: "
                             "the mask is [B, n_shard * Nn] and the scores [B, B * n_shard * Nn] "
                             "(the reference fails on the same shapes, bess.py:182-245)")
        if self.augment_negative:
            half = p // 2 if scheme == "ht" else 0
            n_aug = N - n * Nn
            K.mask_diag(neg, S, N, 1, half, p if scheme == "ht" else 0, BAD_NEGATIVE_SCORE)
            if nmask is not None:
                self._mask_block(neg, S, N, p, n, Nn, nmask, flat, scheme, n_aug)
        elif nmask is not None:
            self._mask_block(neg, S, N, p, n, Nn, nmask, flat, scheme, 0)

    @staticmethod
    def _mask_block(neg: torch.Tensor, S: int, N: int, p: int, n: int, Nn: int,
                    nmask: torch.Tensor, flat: bool, scheme: str, col_off: int) -> None:
        # nmask: [Bm, n*Nn] uint8/bool, True = real negative
        m = nmask.view(torch.uint8) if nmask.dtype == torch.bool else nmask
        width = n * Nn
        if flat and scheme == "ht":
            # row 0 masks the first half of every partition, row 1 the second half
            half = p // 2
            for k in range(2):
                rows_k = neg.view(n, p, N)[:, k * half:(k + 1) * half]
                # strided row blocks: launch per partition (n is small)
                for j in range(n):
                    K.mask_add(rows_k[j], half, width, N, m[k], width, 1, False,
                               BAD_NEGATIVE_SCORE, col_off)
        else:
            K.mask_add(neg, S, width, N, m, width, m.shape[0], False, BAD_NEGATIVE_SCORE, col_off)

    def _update_entity(self, opt, table: torch.Tensor, shard: int, sk, sp, G: int, n_local: int,
                       per: int, dH_l: torch.Tensor, g_dst: int, stride_rows: int,
                       hyper: torch.Tensor, ws: K.Workspace) -> None:
        W = table.shape[1]
        if opt.sparse_exact:
            K.scatter_sgd(table, sk, sp, G, n_local, per, dH_l, g_dst, stride_rows, opt.lr, hyper)
            return
        seg = ws.get("seg_grad", (G, W), torch.float32)
        r2s = ws.get(f"row_to_seg", (table.shape[0],), torch.int32)
        K.fill_i32(r2s, -1)
        K.scatter_collect(W, sk, sp, G, n_local, per, dH_l, g_dst, stride_rows, seg, r2s)
        st = self._opt_state
        key0, key1 = f"ent_s0_{shard}", f"ent_s1_{shard}"
        if key0 not in st:
            st[key0] = torch.zeros(table.shape[0], W, dtype=torch.float32, device=table.device)
        if opt.kind == L.OPT_ADAMW and key1 not in st:
            st[key1] = torch.zeros_like(st[key0])
        b1, b2 = getattr(opt, "betas", (0.0, 0.0))
        K.opt_dense(opt.kind, table, seg, r2s, st[key0], st.get(key1), opt.lr, opt.momentum,
                    opt.dampening, b1, b2, getattr(opt, "eps", 0.0), opt.weight_decay, 1, hyper)

    def _accumulate_entity(self, opt, table: torch.Tensor, shard: int, sk, sp, G: int,
                           n_local: int, per: int, dH_l: torch.Tensor, g_dst: int,
                           stride_rows: int, hyper: torch.Tensor, apply: bool) -> None:
        """Gradient accumulation: add this micro-batch's segment sums to the dense fp32
        accumulator of the shard; on the last micro-batch of the cycle one dense optimizer pass
        consumes (and clears) it.  Rows without gradient see g = 0, so plain SGD stays exact."""
        W = table.shape[1]
        st = self._opt_state
        key = f"ent_acc_{shard}"
        if key not in st:
            st[key] = torch.zeros(table.shape[0], W, dtype=torch.float32, device=table.device)
        K.scatter_accumulate(W, sk, sp, G, n_local, per, dH_l, g_dst, stride_rows, st[key])
        if not apply:
            return
        key0, key1 = f"ent_s0_{shard}", f"ent_s1_{shard}"
        if opt.kind != L.OPT_SGD and key0 not in st:
            st[key0] = torch.zeros_like(st[key])
        if opt.kind == L.OPT_ADAMW and key1 not in st:
            st[key1] = torch.zeros_like(st[key])
        b1, b2 = getattr(opt, "betas", (0.0, 0.0))
        K.opt_dense(opt.kind, table, st[key], None, st.get(key0), st.get(key1), opt.lr,
                    opt.momentum, opt.dampening, b1, b2, getattr(opt, "eps", 0.0),
                    opt.weight_decay, 1, hyper, grad_scale=self._accum_scale, zero_grad=True)

    def _update_relation(self, opt, rel_table: torch.Tensor, d_rel: torch.Tensor,
                         hyper: torch.Tensor, ws: K.Workspace) -> None:
        st = self._opt_state
        if opt.kind != L.OPT_SGD and "rel_s0" not in st:
            st["rel_s0"] = torch.zeros_like(d_rel)
        if opt.kind == L.OPT_ADAMW and "rel_s1" not in st:
            st["rel_s1"] = torch.zeros_like(d_rel)
        b1, b2 = getattr(opt, "betas", (0.0, 0.0))
        K.opt_dense(opt.kind, rel_table, d_rel, None, st.get("rel_s0"), st.get("rel_s1"), opt.lr,
                    opt.momentum, opt.dampening, b1, b2, getattr(opt, "eps", 0.0),
                    opt.weight_decay, 1, hyper)


# ---------------------------------------------------------------------------
# ScoreMoving
# ---------------------------------------------------------------------------
class ScoreMovingBessKGE(BessKGE):
    """Negatives are scored on the shard that stores them; queries are
    replicated (AllGather) and the SCORES travel back (AllToAll) — reference
    bess.py:471-603.  The candidate rows are read in place from the shard
    through the index list (fused gather + score), so the [Q, Nn, D] negative
    tensor of the reference is never materialised.

    Training (under `training_model`) runs the transposed exchange: dL/dscore travels to the
    shards that scored (AllToAll), each computes the gradients of ITS candidate rows — which
    never leave the shard — and of the replicated query vectors; the query gradients are summed
    over the scoring shards in shard order at the shard that owns the query (the transpose of
    the AllGather), then flow through the query prologue into the head / tail rows and the
    relation table.  Eager (not graph-captured)."""

    def _run(self, head, relation, tail, negative, triple_mask, triple_weight, negative_mask,
             optimizer) -> Dict[str, Any]:
        ws, pl = self._setup()
        dev = ws.device
        ent, rel_table = self._tables()
        n = self.sharding.n_shard
        L_rows = head.shape[0]
        bps = L_rows // n
        p = head.shape[-1]
        B, Nn = negative.shape[-2], negative.shape[-1]
        S = n * p
        W, Wr = ent.shape[-1], rel_table.shape[-1]
        cfg = self.score_fn.kernel_cfg()
        dt = L.dtype_code(ent.dtype)
        tdt = ent.dtype
        scheme = self.negative_sampler.corruption_scheme
        flat = self.negative_sampler.flat_negative_format
        triple_based = isinstance(self.negative_sampler, TripleBasedShardedNegativeSampler)
        shared = bool(self.score_fn.negative_sample_sharing)
        R = pl.n_local
        dist = pl.distributed
        train = optimizer is not None
        if train and self.loss_fn is None:
            raise ValueError("training needs a loss_fn")

        def put(name, t, dtype):
            if t is None:
                return None
            if dist:  # this rank only needs its own rows
                t = t[pl.rank::n]
            return self._stage(name, t, dtype, dev)

        h2 = _as_i32(head).reshape(L_rows, S)
        t2 = _as_i32(tail).reshape(L_rows, S)
        gidx = put("gidx", torch.cat([h2, t2], dim=1), torch.int32)
        rel = put("rel", relation.reshape(L_rows, S), torch.int32)
        nidx = put("nidx", negative.reshape(L_rows, n, B, Nn), torch.int32)
        tw = put("tw", None if triple_weight is None else triple_weight.reshape(L_rows, -1),
                 torch.float32)
        nmask = None
        if negative_mask is not None:
            nmask = put("nmask", negative_mask.reshape(L_rows, negative_mask.shape[1], -1),
                        torch.bool)
        tmask = put("tmask", None if triple_mask is None else triple_mask.reshape(L_rows, -1),
                    torch.bool)
        one = ws.get("one", (1,), torch.float32)
        one.fill_(1.0)

        half = p // 2
        # columns every scoring shard contributes to a query: a flat list (replicated for
        # triple-based samplers), the query's own Nn candidates, or — shared, non-flat — the
        # candidates of every query of the group (bess.py:523-534 with negative sample sharing)
        if flat:
            X = Nn if triple_based else n * Nn
        elif shared:
            X = n * (B // 2 if scheme == "ht" else B) * Nn
        else:
            X = Nn
        N = n * X
        # rows of the replicated query arrays are ordered (query shard j, q)
        H = ws.get("H", (n, S, W), tdt)        # head rows of every shard's queries
        T = ws.get("TN", (n, n, p, W), tdt)    # tails received by every replica [replica][src][p]
        rel_all = ws.get("rel_all", (n * S,), torch.int32)
        nvec = K.call("bess_query_nvec", L.C.byref(cfg))
        qv = ws.get("qv", (n * S, nvec, W), torch.float32)
        need_aux = cfg.family == L.BOXE and cfg.norm_p == 2
        need_scale = K.needs_cand_scale(cfg)
        # local mode: the score matrix of all replicas is written in place, column block r*X by
        # scoring shard r.  distributed: this rank scores all n*S queries against ITS candidates
        # ([n*S, X]) and the scores travel back to the shards that own the queries (AllToAll).
        ld_sc = N if not dist else X
        aux = ws.get("aux", (n * S, ld_sc), torch.float32) if need_aux else None
        n_out = bps * R
        pos_out = torch.empty(n_out * S, dtype=torch.float32, device=dev)
        neg_out = torch.empty(n_out * S, N, dtype=torch.float32, device=dev)
        loss_out = (torch.empty(n_out, dtype=torch.float32, device=dev)
                    if self.loss_fn is not None else None)
        # Distributed inference over peer memory (no collective library inside the step): the
        # query rows are stored straight into every rank's replicated query arrays, and the
        # scoring kernels store each owner's block of scores straight into that owner's score
        # matrix over NVLink — compute and exchange are one kernel.  Two flag handshakes per
        # micro-batch.  Training keeps the collective path below.
        peer = bool(dist and USE_PEER_EXCHANGE and not train and S % 4 == 0)
        px: Optional[_PeerExchange] = None
        if peer:
            if self._px is None:
                self._px = _PeerExchange(dev, n, pl.rank)
            px = self._px
            es = ent.element_size()
            h_bytes = _up(n * S * W * es)
            px.ensure(h_bytes + n * S * W * es, S * N * 4, n * S * 4)
            H = px.view(px.off_tn, (n, S, W), tdt)
            T = px.view(px.off_tn + h_bytes, (n, n, p, W), tdt)
            sc_sym = px.view(px.off_grad, (S, N), torch.float32)
            rel_all = px.view(px.off_rel, (n * S,), torch.int32)
            off_H, off_T = px.off_tn, px.off_tn + h_bytes
            rep_h = ws.get("sm_rep_h", (n, S), torch.int32)
            if need_aux:
                aux = ws.get("aux", (S, N), torch.float32)
        elif dist:
            SEND_T = ws.get("SEND", (n, p, W), tdt)
            sc_local = ws.get("sc_local", (n * S, X), torch.float32)
            sc_recv = ws.get("sc_recv", (n, S, X), torch.float32)
        ce = self.loss_fn is not None and self.loss_fn._kind == L.LOSS_SOFTMAX_CE
        if train:
            d_pos = ws.get("d_pos", (R * S,), torch.float32)
            d_neg = ws.get("d_neg", (R * S, N), torch.float32)
            dH = ws.get("sm_dH", (n, S, W), torch.float32)       # local: [replica]; dist: row `me`
            dT = ws.get("sm_dT", (n, n, p, W), torch.float32)    # gradients of T, same layout
            dRq = ws.get("dRq", (R * S, Wr), torch.float32)
            d_qv = ws.get("d_qv", (n * S, nvec, W), torch.float32)
            d_sc = d_neg if not dist else ws.get("sm_dsc", (n * S, X), torch.float32)
            sort_n = max(2 * S, R * S, n * B * Nn, 0 if flat else n * S * Nn)
            sort_ws = ws.get("sort_ws", (K.sort_workspace(sort_n) // 4 + 64,), torch.int32)
            d_rel_table = ws.get("d_rel_table", (_up(rel_table.numel() * 4, 16) // 4,),
                                 torch.float32)[:rel_table.numel()].view(rel_table.shape)
            key_bits = max(1, int(ent.shape[1] - 1).bit_length())
            rel_bits = max(1, int(rel_table.shape[0] - 1).bit_length())
        acc: Dict[str, List] = {}
        stamps = None
        if STAGE_STAMPS and peer:
            stamps = ws.get_zeroed("stage_stamps", (bps, len(SM_STAGE_NAMES)), torch.int64)

        def stamp(s_, name):
            if stamps is not None:
                K.stamp(stamps[s_, SM_STAGE_NAMES.index(name)])

        for s in range(bps):
            stamp(s, "start")
            if peer:
                row0, me = s, pl.rank
                es = ent.element_size()
                # my heads -> H[me] on EVERY rank; my tails for replica j -> T[j][me] on rank j
                # (on every rank when the tails are the query side): remote stores over NVLink
                rep_h.copy_(gidx[row0][:S].unsqueeze(0).expand(n, S))
                K.gather_route(ent[me], rep_h.view(-1), 0, S, None, [q + off_H for q in px.ptrs], me)
                t_idx = gidx[row0][S:]
                blk = n * p * W * es  # one replica's [n(src), p, W] block of T
                if scheme == "t":
                    K.gather_route(ent[me], t_idx, 0, p, None,
                                   [px.ptrs[j] + off_T + j * blk for j in range(n)], me)
                else:
                    for k in range(n):
                        K.gather_route(ent[me], t_idx, 0, p, None,
                                       [px.ptrs[k] + off_T + j * blk for j in range(n)], me)
                K.peer_push(rel[row0], 0, [q + px.off_rel + me * S * 4 for q in px.ptrs], S * 4)
                stamp(s, "query rows stored")
                px.handshake(0)
                stamp(s, "query rows arrived")
                pos = pos_out[s * S:(s + 1) * S]
                K.triple_fwd(cfg, dt, L.rows(H[me]), L.rows(T[me].view(S, W)), rel_table, rel[row0],
                             L.IDENT, S, pos, L.IDENT)
                score_buf, scorers = sc_sym, [(me, row0, me * X)]
            elif dist:
                row0 = s
                me = pl.rank
                # own heads -> H[me]; tails routed to the replica that scores the positive
                K.gather_route(ent[me], gidx[row0], S, p, H[me],
                               [SEND_T[j].data_ptr() for j in range(n)], 0)
                pl.all_to_all(T[me], SEND_T)
                pos = pos_out[s * S:(s + 1) * S]
                K.triple_fwd(cfg, dt, L.rows(H[me]), L.rows(T[me].view(S, W)), rel_table, rel[row0],
                             L.IDENT, S, pos, L.IDENT)
                # replicate the queries (bess.py:519-545)
                torch.distributed.all_gather_into_tensor(rel_all, rel[row0].contiguous())
                if scheme in ("t", "ht"):
                    torch.distributed.all_gather_into_tensor(H.view(-1), H[me].reshape(-1).clone())
                if scheme in ("h", "ht"):
                    torch.distributed.all_gather_into_tensor(T.view(-1), T[me].reshape(-1).clone())
                score_buf, scorers = sc_local, [(me, row0, 0)]
            else:
                base = s * n
                for r in range(n):
                    K.gather_route(ent[r], gidx[base + r], S, p, H[r],
                                   [T[j].data_ptr() for j in range(n)], r)
                    rel_all[r * S:(r + 1) * S].copy_(rel[base + r])
                pos = pos_out[base * S:(base + n) * S]
                K.triple_fwd(cfg, dt, L.rows(H.view(n * S, W)), L.rows(T.view(n * S, W)), rel_table,
                             rel_all, L.IDENT, n * S, pos, L.IDENT)
                score_buf = neg_out[base * S:(base + n) * S]  # [n*S, N]
                scorers = [(r, base + r, r * X) for r in range(n)]
            Hall, Tall = H.view(n * S, W), T.view(n * S, W)
            # (mode, query map over the n*S queries, number of queries, fixed rows, flat-list id)
            if scheme == "t":
                groups = [(L.MODE_TAILS, L.IDENT, n * S, Hall, 0)]
            elif scheme == "h":
                groups = [(L.MODE_HEADS, L.IDENT, n * S, Tall, 0)]
            else:
                groups = [
                    (L.MODE_HEADS, L.rowmap(half, p, 0), n * n * half, Tall, 0),
                    (L.MODE_TAILS, L.rowmap(half, p, half), n * n * half, Hall, 1),
                ]

            def candidates(shard: int, row: int, bsel: int):
                """(index list into shard `shard`, shared?, candidates per query) of one group."""
                idx_r = nidx[row]  # [n(query shard), B, Nn]: candidates stored on `shard`
                if flat:
                    if triple_based:  # one replicated list; "ht": b selects the heads / tails list
                        return idx_r[0, bsel if scheme == "ht" else 0].contiguous(), True, Nn
                    if scheme == "ht":
                        return idx_r[:, bsel].contiguous().view(-1), True, n * Nn
                    return idx_r.reshape(-1), True, n * Nn
                if shared and scheme == "ht":
                    # negatives of the first half of every partition's triples corrupt heads,
                    # of the second half tails (bess.py:556-566): the group's own sub-list
                    blk = idx_r.view(n, n, p, Nn)[:, :, bsel * half:(bsel + 1) * half]
                    sel = blk.contiguous().view(-1)
                    return sel, True, sel.numel()
                # per-query candidates: the query at position pos (among the n*S replicated
                # queries) reads rows [pos * Nn, (pos + 1) * Nn) of the full list
                sel = idx_r.reshape(-1)
                return sel, shared, (sel.numel() if shared else Nn)

            def score_group(gi, shard, row, col0, backward: bool):
                mode, qmap, nq, fixed_buf, bsel = groups[gi]
                table = ent[shard]
                sel, is_shared, n_c = candidates(shard, row, bsel)
                cand = L.rows(table, idx=sel)
                scale = None
                if is_shared and need_scale:
                    scale = K.cand_scales(cfg, dt, cand, n_c, W,
                                          ws.get("cand_scale", (2 * n_c,), torch.float32))
                if not backward and peer:
                    # block j of the (owner-major) queries is scored into rank j's score matrix
                    nqj = nq // n
                    for j in range(n):
                        dst = px.ptrs[j] + px.off_grad
                        qv_j = qv[j * nqj:]
                        if is_shared:
                            K.shared_fwd(cfg, dt, mode, qv_j, nqj, cand, scale, n_c, score_buf, qmap,
                                         N, col0, aux, out_ptr=dst)
                        else:
                            cand_j = L.rows(table, idx=sel[j * S * Nn:])
                            K.pertriple_fwd(cfg, dt, mode, qv_j, nqj, cand_j, Nn, Nn, score_buf,
                                            qmap, N, col0, aux, out_ptr=dst)
                    return None
                if not backward:
                    if is_shared:
                        K.shared_fwd(cfg, dt, mode, qv, nq, cand, scale, n_c, score_buf, qmap,
                                     ld_sc, col0, aux)
                    else:
                        K.pertriple_fwd(cfg, dt, mode, qv, nq, cand, Nn, Nn, score_buf, qmap, ld_sc,
                                        col0, aux)
                    return None
                # ---- backward of this (group, scoring shard): d_qv part + candidate-row grads
                part = d_qv_parts[scorers_pos[shard]]
                if is_shared:
                    d_c = ws.get(f"sm_dc{gi}", (R, n_c, W), torch.float32)[scorers_pos[shard]]
                    cws = ws.get("cand_ws", (max(K.shared_bwd_cand_workspace(cfg, nq, n_c) // 4,
                                                 1),), torch.float32)
                    K.shared_bwd_query(cfg, dt, mode, qv, nq, cand, scale, n_c, score_buf, d_sc,
                                       qmap, ld_sc, col0, aux, part)
                    K.shared_bwd_cand(cfg, dt, mode, qv, nq, cand, scale, n_c, score_buf, d_sc,
                                      qmap, ld_sc, col0, aux, L.rows(d_c), cws, add=False)
                    if need_scale:
                        K.cand_scales_bwd(cfg, dt, cand, n_c, W, scale, L.rows(d_c))
                    return sel, d_c
                # one gradient row per (query position, candidate): both "ht" groups write disjoint
                # rows of the same buffer, which is scattered once
                d_c = ws.get("sm_dc_pt", (R, n * S * Nn, W), torch.float32)[scorers_pos[shard]]
                K.pertriple_bwd(cfg, dt, mode, qv, nq, cand, Nn, Nn, score_buf, d_sc, qmap, ld_sc,
                                col0, aux, part, L.rows(d_c))
                return (sel, d_c) if gi == len(groups) - 1 else None

            scorers_pos = {shard: i for i, (shard, _, _) in enumerate(scorers)}
            for gi, (mode, qmap, nq, fixed_buf, bsel) in enumerate(groups):
                K.prologue_fwd(cfg, dt, mode, L.rows(fixed_buf, rmap=qmap), rel_table, rel_all,
                               qmap, nq, qv)
                for shard, row, col0 in scorers:
                    score_group(gi, shard, row, col0, False)
            stamp(s, "scored")
            if peer:
                px.handshake(1)  # every scoring rank has stored its columns of my score matrix
                stamp(s, "scores arrived")
                neg_out[s * S:(s + 1) * S].copy_(sc_sym)
                finals = [(s, s, pos_out[s * S:(s + 1) * S], neg_out[s * S:(s + 1) * S])]
            elif dist:
                # scores back to the shards that own the queries (bess.py:583-592)
                torch.distributed.all_to_all_single(sc_recv.view(-1), sc_local.view(-1))
                neg_out[s * S:(s + 1) * S].view(S, n, X).copy_(sc_recv.transpose(0, 1))
                finals = [(s, s, pos_out[s * S:(s + 1) * S], neg_out[s * S:(s + 1) * S])]
            else:
                finals = [(base + r, base + r, pos[r * S:(r + 1) * S],
                           score_buf[r * S:(r + 1) * S]) for r in range(n)]
            for li, (o, row, pos_r, neg_r) in enumerate(finals):
                if nmask is not None:
                    EmbeddingMovingBessKGE._mask_block(neg_r, S, N, p, n, X, nmask[row], flat,
                                                       scheme, 0)
                if self.loss_fn is not None:
                    w = tw[row] if tw is not None else one
                    neg_l = neg_r.clone() if (ce and train) else neg_r  # CE shifts scores in place
                    loss, _, _ = self.loss_fn.fwd_bwd(
                        pos_r, neg_l, w, d_pos[li * S:(li + 1) * S] if train else None,
                        d_neg[li * S:(li + 1) * S] if train else None)
                    loss_out[o] = loss
                if self.evaluation is not None:
                    self._finish_metrics({}, pos_r, neg_r,
                                         tmask[row] if tmask is not None else None, acc)
            stamp(s, "masks+metrics")
            if not train:
                continue

            # ======================= backward =======================
            hyper = self._hyper(bps)[s]
            if dist:
                me = pl.rank
                # dL/dscore back to the shards that scored: transpose of the score AllToAll
                send = ws.get("sm_dsend", (n, S, X), torch.float32)
                send.copy_(d_neg.view(S, n, X).transpose(0, 1))
                torch.distributed.all_to_all_single(d_sc.view(-1), send.view(-1))
                # the scores this rank computed (sc_local) are still in place for the kernels
                K.triple_bwd(cfg, dt, L.rows(H[me]), L.rows(T[me].view(S, W)), rel_table, rel[row0],
                             L.IDENT, S, pos_out[s * S:(s + 1) * S], d_pos, L.IDENT, L.rows(dH[me]),
                             L.rows(dT[me].view(S, W)), dRq, False, False, False)
            else:
                K.triple_bwd(cfg, dt, L.rows(Hall), L.rows(Tall), rel_table, rel_all, L.IDENT, n * S,
                             pos, d_pos, L.IDENT, L.rows(dH.view(n * S, W)),
                             L.rows(dT.view(n * S, W)), dRq, False, False, False)
            cand_parts: Dict[int, List[Tuple[torch.Tensor, torch.Tensor]]] = {}
            for gi, (mode, qmap, nq, fixed_buf, bsel) in enumerate(groups):
                K.prologue_fwd(cfg, dt, mode, L.rows(fixed_buf, rmap=qmap), rel_table, rel_all,
                               qmap, nq, qv)
                cnt = nq * nvec * W
                # one slot of query gradients per scoring shard (local mode) / this rank's (dist.)
                d_qv_parts = ws.get("sm_dqv", (len(scorers), cnt), torch.float32)
                for shard, row, col0 in scorers:
                    got = score_group(gi, shard, row, col0, True)
                    if got is not None:
                        cand_parts.setdefault(shard, []).append(got)
                # query gradients: sum over the scoring shards in shard order at the owner
                if dist:
                    # queries are ordered (owner shard j, q): block j of the part goes to rank j
                    recv = ws.get("sm_dqv_recv", (n, cnt // n), torch.float32)
                    torch.distributed.all_to_all_single(recv.view(-1), d_qv_parts.view(-1))
                    K.peer_reduce(recv, n, cnt // n, 1.0, d_qv)
                    own = nq // n  # this rank's queries of the group, at positions qmap(q) of its S
                    d_fix = dT[me].view(S, W) if mode == L.MODE_HEADS else dH[me]
                    fix = T[me].view(S, W) if mode == L.MODE_HEADS else H[me]
                    K.prologue_bwd(cfg, dt, mode, L.rows(fix, rmap=qmap), rel_table, rel[row0],
                                   qmap, own, d_qv, L.rows(d_fix, rmap=qmap), dRq, True, True)
                else:
                    K.peer_reduce(d_qv_parts, n, cnt, 1.0, d_qv)
                    d_fix = dT.view(n * S, W) if mode == L.MODE_HEADS else dH.view(n * S, W)
                    K.prologue_bwd(cfg, dt, mode, L.rows(fixed_buf, rmap=qmap), rel_table, rel_all,
                                   qmap, nq, d_qv, L.rows(d_fix, rmap=qmap), dRq, True, True)
            if cfg.family == L.BOXE:
                K.boxe_rel_finalize(cfg, dt, rel_table, rel[row0] if dist else rel_all,
                                    R * S, dRq)
            # ---- relation table: reduce per-query rows, all-reduce over ranks, update
            rk = ws.get("rel_keys", (R * S,), torch.int32)
            rp = ws.get("rel_perm", (R * S,), torch.int32)
            K.sort_keys((rel[row0] if dist else rel_all).contiguous(), R * S, rel_bits, rk, rp,
                        sort_ws)
            K.relation_grad_reduce(dRq, Wr, rk, rp, R * S, rel_table.shape[0], d_rel_table)
            if dist:
                torch.distributed.all_reduce(d_rel_table)
            if getattr(optimizer, "relation_grad_reduction", "mean") == "mean" and n > 1:
                d_rel_table.mul_(1.0 / n)
            # ---- entity shards: heads (local), tails (rows this shard sent to every replica)
            # and the candidate rows, which never left the shard
            if dist:
                dT_back = ws.get("sm_dT_back", (n, p, W), torch.float32)
                pl.all_to_all(dT_back, dT[me])  # tail gradients back to the shards that own them
            for li, (shard, row, _) in enumerate(scorers):
                if dist:
                    parts = [(gidx[row], 2 * S, S, p, dH[me], dT_back.data_ptr(), p)]
                else:
                    parts = [(gidx[row], 2 * S, S, p, dH[shard],
                              dT.data_ptr() + shard * p * W * 4, n * p)]
                for sel, d_c in cand_parts[shard]:
                    parts.append((sel, sel.numel(), sel.numel(), 1, d_c, 0, 0))
                self._apply_parts(optimizer, ent[shard], shard, parts, key_bits, hyper, sort_ws, ws)
            EmbeddingMovingBessKGE._update_relation(self, optimizer, rel_table, d_rel_table, hyper,
                                                    ws)

        out: Dict[str, Any] = {}
        if self.return_scores:
            out["positive_score"] = pos_out if tdt == torch.float32 else pos_out.to(tdt)
            out["negative_score"] = neg_out if tdt == torch.float32 else neg_out.to(tdt)
        if loss_out is not None:
            out["loss"] = loss_out
        if "ranks" in acc:
            out["ranks"] = torch.cat(acc["ranks"])
        if "metrics" in acc:
            out["metrics"] = torch.cat(acc["metrics"], dim=0)
        return out

    def _apply_parts(self, opt, table: torch.Tensor, shard: int, parts, key_bits: int,
                     hyper: torch.Tensor, sort_ws: torch.Tensor, ws: K.Workspace) -> None:
        """Scatter several (index list, gradient rows) parts into one shard and update it.
        Plain SGD applies each part sparsely (exact); momentum / AdamW first accumulate every
        part into the shard's dense fp32 accumulator, then run one dense-semantics pass."""
        W = table.shape[1]
        st = self._opt_state
        if not opt.sparse_exact:
            key = f"ent_acc_{shard}"
            if key not in st:
                st[key] = torch.zeros(table.shape[0], W, dtype=torch.float32, device=table.device)
        for keys, G, n_local, per, g_local, g_dst, stride_rows in parts:
            sk = ws.get("sm_sort_keys", (G,), torch.int32)
            sp = ws.get("sm_sort_perm", (G,), torch.int32)
            K.sort_keys(keys.contiguous(), G, key_bits, sk, sp, sort_ws)
            if opt.sparse_exact:
                K.scatter_sgd(table, sk, sp, G, n_local, per, g_local, g_dst, stride_rows, opt.lr,
                              hyper)
            else:
                K.scatter_accumulate(W, sk, sp, G, n_local, per, g_local, g_dst, stride_rows,
                                     st[f"ent_acc_{shard}"])
        if opt.sparse_exact:
            return
        key0, key1 = f"ent_s0_{shard}", f"ent_s1_{shard}"
        if opt.kind != L.OPT_SGD and key0 not in st:
            st[key0] = torch.zeros_like(st[f"ent_acc_{shard}"])
        if opt.kind == L.OPT_ADAMW and key1 not in st:
            st[key1] = torch.zeros_like(st[f"ent_acc_{shard}"])
        b1, b2 = getattr(opt, "betas", (0.0, 0.0))
        K.opt_dense(opt.kind, table, st[f"ent_acc_{shard}"], None, st.get(key0), st.get(key1),
                    opt.lr, opt.momentum, opt.dampening, b1, b2, getattr(opt, "eps", 0.0),
                    opt.weight_decay, 1, hyper, grad_scale=1.0, zero_grad=True)


# ---------------------------------------------------------------------------
# training wrapper (the B200 counterpart of poptorch.trainingModel)
# ---------------------------------------------------------------------------
class TrainingModel:
    """Callable returned by `training_model`: one call = `batches_per_step`
    fused forward + backward + optimizer steps; returns the forward dict.

    With `cuda_graph=True` (the default) the whole device sequence of a call —
    gather, exchange, scoring, loss, backward, sort, scatter + update — is
    captured once per input signature into a CUDA graph and replayed: the B200
    counterpart of PopTorch's `deviceIterations` loop (one host launch per call
    instead of ~40 kernel launches).  The first call with a new signature runs
    eagerly (it sizes the persistent workspaces), the second is captured.  The
    returned tensors are then STATIC buffers that the next call overwrites.

    What a captured graph does NOT freeze:
      * optimizer hyper-parameters — `optimizer.lr` (and momentum, weight decay, betas, eps)
        are re-read on every call and handed to the kernels through a small device array, as
        is the step count of AdamW's bias correction, so learning-rate schedules work;
      * buffer addresses — whenever a persistent workspace buffer or the symmetric-memory
        exchange buffer is re-allocated (a larger batch, another dtype, a validation
        `model.forward()` on the same module) every captured graph is dropped and re-captured
        on its next use; replaced buffers stay allocated until then.

    Optimizer state (step count, momentum / Adam moments of the local shards and of the
    relation table) is exposed through `state_dict()` / `load_state_dict()`."""

    def __init__(self, model: BessKGE, optimizer: Union[SGD, AdamW],
                 relation_grad_reduction: str = "mean", cuda_graph: Optional[bool] = None,
                 gradient_accumulation: int = 1, accumulation_reduction: str = "mean") -> None:
        if relation_grad_reduction not in ("mean", "sum"):
            raise ValueError("relation_grad_reduction must be 'mean' or 'sum'")
        if accumulation_reduction not in ("mean", "sum"):
            raise ValueError("accumulation_reduction must be 'mean' or 'sum'")
        if gradient_accumulation < 1:
            raise ValueError("gradient_accumulation must be >= 1")
        model._accum_k = int(gradient_accumulation)
        model._accum_scale = (1.0 / gradient_accumulation if accumulation_reduction == "mean"
                              else 1.0)
        self.model = model
        self.optimizer = optimizer
        self.optimizer.relation_grad_reduction = relation_grad_reduction
        graph_ok = isinstance(model, EmbeddingMovingBessKGE)
        self.cuda_graph = graph_ok if cuda_graph is None else (bool(cuda_graph) and graph_ok)
        self._graphs: Dict[Any, Any] = {}
        self._graph_gen: Optional[Tuple[int, int, int]] = None

    def _use_graph(self) -> bool:
        # NCCL collectives inside a captured step hung on 2 x B200 (NCCL 2.28.9, torch 2.11);
        # with the peer-memory exchange the step contains only this library's kernels
        return self.cuda_graph and (not self.model._setup()[1].distributed or USE_PEER_EXCHANGE)

    def __call__(self, head, relation, tail, negative, triple_mask=None, triple_weight=None,
                 negative_mask=None) -> Dict[str, Any]:
        if not isinstance(self.model, EmbeddingMovingBessKGE):
            # ScoreMovingBessKGE trains eagerly (its exchange uses collectives)
            self.model._setup()
            self.model._begin_steps(self.optimizer, head.shape[0] // self.model.sharding.n_shard)
            return self.model._run(head, relation, tail, negative, triple_mask, triple_weight,
                                   negative_mask, optimizer=self.optimizer)
        staged = self.model.stage(head, relation, tail, negative, triple_mask, triple_weight,
                                  negative_mask)  # H2D into the persistent input buffers
        return self._run(staged)

    def stage(self, **batch) -> _Staged:
        """Copy a batch to the device once; see `run_staged`."""
        return self.model.stage(persistent=True, **batch)

    def run_staged(self, staged: _Staged) -> Dict[str, Any]:
        """Training step(s) on inputs that are already resident in HBM."""
        if not self._use_graph():
            return self._run(staged)
        return self._run(self.model._restage(staged))

    def _eager(self, staged: _Staged) -> Dict[str, Any]:
        return self.model._run(None, None, None, None, None, None, None,
                               optimizer=self.optimizer, staged=staged)

    def _generation(self) -> Tuple[int, int, int]:
        m = self.model
        return (id(m._ws), m._ws.generation, m._px.generation if m._px is not None else 0)

    def _run(self, staged: _Staged) -> Dict[str, Any]:
        bps = staged.dims[0] // staged.dims[1]
        self.model._begin_steps(self.optimizer, bps)
        if not self._use_graph():
            return self._eager(staged)
        o = self.optimizer
        key = (staged.dims, staged.tw is not None, staged.nmask is not None,
               staged.tmask is not None, self.model.score_fn.entity_embedding.dtype,
               o.kind, bool(o.sparse_exact), self.model._accum_phase0)
        if self._graph_gen != self._generation():
            # a buffer some captured graph points into was replaced: drop every graph (their
            # old buffers were kept alive until now) and start over
            self._graphs.clear()
            self.model._ws.release_retired()
            self._graph_gen = self._generation()
        entry = self._graphs.get(key)
        if entry is None:
            self._graphs[key] = "warm"
            return self._eager(staged)
        if entry == "warm":
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph):
                out = self._eager(staged)
            if self._graph_gen != self._generation():
                # capture re-allocated a buffer (sized differently than the warm run): the
                # recorded pointers are already stale; nothing ran, so run this call eagerly
                self._graphs.clear()
                self._graph_gen = None
                del graph
                return self._eager(staged)
            entry = self._graphs[key] = (graph, out)
        graph, out = entry
        graph.replay()
        return out

    # ---- optimizer state -------------------------------------------------------
    def state_dict(self) -> Dict[str, Any]:
        """Optimizer state of THIS process: `step` and fp32 tensors `ent_s0_<shard>` /
        `ent_s1_<shard>` (momentum buffer or Adam m / v of each local shard) and `rel_s0` /
        `rel_s1` (relation table), cloned."""
        return {k: (v.clone() if isinstance(v, torch.Tensor) else v)
                for k, v in self.model._opt_state.items()}

    def load_state_dict(self, state: Dict[str, Any]) -> None:
        st = self.model._opt_state
        dev = self.model._device()
        for k, v in state.items():
            if isinstance(v, torch.Tensor):
                if k in st and st[k].shape == v.shape:
                    st[k].copy_(v)  # in place: captured graphs keep pointing at it
                else:
                    st[k] = v.to(device=dev, dtype=torch.float32).clone()
                    self._graphs.clear()
            else:
                st[k] = v


def training_model(model: BessKGE, optimizer: Union[SGD, AdamW],
                   relation_grad_reduction: str = "mean",
                   cuda_graph: Optional[bool] = None, gradient_accumulation: int = 1,
                   accumulation_reduction: str = "mean") -> TrainingModel:
    """Counterpart of `poptorch.trainingModel(model, options, optimizer)`
    (reference notebooks, e.g. 1_biokg cell 28).  `relation_grad_reduction`:
    how the replicated relation table's gradient is combined over replicas
    ("mean" = PopTorch default, "sum").  `gradient_accumulation = k`
    (`options.Training.gradientAccumulation(k)`, notebook 1 cell 26): gradients of k
    consecutive micro-batches — all computed from the same weights — are accumulated and
    combined by `accumulation_reduction` before one optimizer step."""
    return TrainingModel(model, optimizer, relation_grad_reduction, cuda_graph,
                         gradient_accumulation, accumulation_reduction)


class TopKQueryBessKGE(torch.nn.Module):
    """Top-k completion of (h, r, ?) / (?, r, t) queries against all entities or
    a given candidate set (reference: bess.py:606-921).  Queries of every shard
    are replicated (AllGather), scored on the shard that stores the candidates,
    a running top-(k+1) is kept per (query, scoring shard), the best lists
    travel back (AllToAll) and the owner merges them.

    B200 shape of the loop: the reference's `window_size` sliding window
    (`poptorch.for_loop`, bess.py:771-853) exists to bound IPU SRAM; the result
    does not depend on it (top-k is exact), so here the window is a GEMM N-tile
    of `device_window` candidates: scores of all n*S queries against one window
    come from the tcgen05 GEMM (DistMult / ComplEx) or the CUDA-core tile kernel
    (distance families) and `bess_topk_merge` folds them into the best lists.
    `window_size` is kept for API compatibility.  Inference only."""

    device_window = 4096
    #: Exact ranking: against ALL entities the window scorers keep the
    #: top-(k + 1 + exact_margin) local candidates per (query, shard), which are then re-scored
    #: in one fixed fp32 summation order and re-sorted (csrc/exact.cu) — ids, scores, ranks and
    #: MRR are then bit-identical to the oracle's restatement of that arithmetic, with no
    #: near-tie tolerance.  Applies to TransE / DistMult / ComplEx (reproducible arithmetic)
    #: when k + 1 + exact_margin <= 32; other cases keep the scorers' own fp32 scores.
    exact_rerank = True
    exact_margin = 6

    def __init__(
        self,
        k: int,
        candidate_sampler: Union[TripleBasedShardedNegativeSampler, PlaceholderNegativeSampler],
        score_fn: BaseScoreFunction,
        evaluation: Optional[Evaluation] = None,
        return_scores: bool = False,
        window_size: int = 100,
    ) -> None:
        super().__init__()
        self.sharding = score_fn.sharding
        self.negative_sampler = candidate_sampler
        self.score_fn = score_fn
        self.evaluation = evaluation
        self.return_scores = return_scores
        self.k = k
        self.window_size = window_size
        if self.negative_sampler.flat_negative_format:
            assert (
                score_fn.negative_sample_sharing
            ), "Using flat negative format requires negative sample sharing"
        elif score_fn.negative_sample_sharing:
            raise ValueError(
                "Negative sample sharing cannot be used with non-flat triple-specific negatives"
            )
        if self.negative_sampler.corruption_scheme not in ["h", "t"]:
            raise ValueError("TopKQueryBessKGE only support 'h', 't' corruption scheme")
        if isinstance(self.negative_sampler, TripleBasedShardedNegativeSampler):
            assert self.negative_sampler.mask_on_gather, (
                "TopKQueryBessKGE requires setting mask_on_gather=True in the candidate_sampler"
            )
        self.entity_embedding = self.score_fn.entity_embedding
        self.entity_embedding_size: int = self.entity_embedding.shape[-1]
        self._ws: Optional[K.Workspace] = None
        self._placement: Optional[_Placement] = None
        self._maps: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
        # cached 3xTF32 operands of each local fp32 shard: li -> (key, hi, lo, state)
        self._table_ops: Dict[int, Tuple[Any, torch.Tensor, torch.Tensor, torch.Tensor]] = {}

    def _table_operand(self, li: int, table: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, int]:
        """hi / lo [Es, ld] tensor-core operands of an fp32 shard.  The shard is constant
        during inference, so the split is done once; every call re-checks a device-side
        checksum of the table (one read-only pass) and rebuilds only if the bytes changed."""
        ld = _pad8(table.shape[1])
        key = (table.data_ptr(), tuple(table.shape), table.stride(0), table.device)
        ent = self._table_ops.get(li)
        force = ent is None or ent[0] != key
        if force:
            hi = torch.empty(table.shape[0], ld, dtype=torch.float32, device=table.device)
            lo = torch.empty_like(hi)
            state = torch.zeros(4, dtype=torch.int64, device=table.device)
            ent = (key, hi, lo, state)
            self._table_ops[li] = ent
        K.table_operand_refresh(table, ent[1], ent[2], ld, ent[3], force)
        return ent[1], ent[2], ld

    def _setup(self) -> Tuple[K.Workspace, _Placement]:
        dev = self.score_fn.entity_embedding.device
        if dev.type != "cuda":
            raise L.BessLibraryError(
                "besskge_b200 has no CPU path: move the module to a CUDA device first"
            )
        if self._ws is None or self._ws.device != dev:
            self._ws = K.Workspace(dev)
            self._maps = None
        if self._placement is None:
            self._placement = _Placement(self.sharding.n_shard)
        if self._maps is None:
            self._maps = (
                torch.from_numpy(np.ascontiguousarray(self.sharding.shard_counts)).to(
                    device=dev, dtype=torch.int32),
                torch.from_numpy(np.ascontiguousarray(self.sharding.shard_and_idx_to_entity)).to(
                    device=dev, dtype=torch.int32),
            )
        return self._ws, self._placement

    def forward(
        self,
        relation: torch.Tensor,
        head: Optional[torch.Tensor] = None,
        tail: Optional[torch.Tensor] = None,
        negative: Optional[torch.Tensor] = None,
        triple_mask: Optional[torch.Tensor] = None,
        negative_mask: Optional[torch.Tensor] = None,
    ) -> Dict[str, Any]:
        """Shapes as the reference (bess.py:691-740) with a leading `bps * n_shard`
        axis: relation / head / tail [L, S]; negative / negative_mask
        [L, n_shard, B, Nn] (candidates local to the shard of row L) or None to
        score against every entity; triple_mask [L, S]."""
        ws, pl = self._setup()
        dev = ws.device
        counts_dev, s2e_dev = self._maps
        ent = self.score_fn.entity_embedding.data
        rel_table = self.score_fn.relation_embedding.data
        n = self.sharding.n_shard
        Es = ent.shape[1]
        W = ent.shape[-1]
        cfg = self.score_fn.kernel_cfg()
        tdt = ent.dtype
        dt = L.dtype_code(tdt)
        scheme = self.negative_sampler.corruption_scheme
        mode = L.MODE_TAILS if scheme == "t" else L.MODE_HEADS
        fixed_h, truth_h = (head, tail) if scheme == "t" else (tail, head)
        if fixed_h is None:
            raise ValueError("queries need the known entity: head for 't', tail for 'h'")
        L_rows, S = relation.shape[0], relation.shape[-1]
        if L_rows % n != 0:
            raise ValueError(f"leading axis {L_rows} is not a multiple of n_shard={n}")
        bps = L_rows // n
        kb = self.k + 1
        nS = n * S
        R = pl.n_local
        flat = self.negative_sampler.flat_negative_format

        def to_dev(t, dtype):
            if t is None:
                return None
            if pl.distributed:
                t = t[pl.rank::n]
            return t.to(device=dev, dtype=dtype, non_blocking=True).contiguous()

        rel = to_dev(relation.reshape(L_rows, S), torch.int32)
        fixed = to_dev(fixed_h.reshape(L_rows, S), torch.int32)
        truth = to_dev(None if truth_h is None else truth_h.reshape(L_rows, S), torch.int32)
        tmask = to_dev(None if triple_mask is None else triple_mask.reshape(L_rows, S), torch.bool)
        cand_idx = cand_mask = None
        if negative is not None:
            assert negative_mask is not None
            Nn = negative.shape[-1]
            cand_idx = to_dev(negative.reshape(L_rows, -1, Nn), torch.int32)
            cand_mask = to_dev(negative_mask.reshape(L_rows, -1, Nn), torch.uint8)
            n_cand_total = Nn
        else:
            n_cand_total = Es

        win = min(self.device_window, _pad8(n_cand_total))
        nvec = K.call("bess_query_nvec", L.C.byref(cfg))
        use_tc = USE_TENSOR_CORES and cfg.family in (L.DISTMULT, L.COMPLEX) and (
            negative is None or flat)
        need_aux = cfg.family == L.BOXE and cfg.norm_p == 2
        need_scale = K.needs_cand_scale(cfg)
        Q = ws.get("tk_Q", (nS, W), tdt)
        rel_all = ws.get("tk_rel", (nS,), torch.int32)
        qv = ws.get("tk_qv", (nS, nvec, W), torch.float32)
        scores = ws.get("tk_scores", (nS, win), torch.float32)
        aux = ws.get("tk_aux", (nS, win), torch.float32) if need_aux else None
        scale = ws.get("tk_scale", (2 * win,), torch.float32) if need_scale else None
        # running best lists: kb entries, or kb + margin when they are re-ranked exactly
        exact = bool(self.exact_rerank and negative is None and kb + self.exact_margin <= 32
                     and K.topk_exact_supported(cfg.family))
        kbm = kb + self.exact_margin if exact else kb
        best_s = ws.get("tk_best_s", (R, nS, kbm), torch.float32)
        best_i = ws.get("tk_best_i", (R, nS, kbm), torch.int32)
        if exact:
            send_s = ws.get("tk_send_s", (R, nS, kb), torch.float32)
            send_i = ws.get("tk_send_i", (R, nS, kb), torch.int32)
        else:
            send_s, send_i = best_s, best_i
        recv_s = ws.get("tk_recv_s", (R, n, S, kb), torch.float32)
        recv_i = ws.get("tk_recv_i", (R, n, S, kb), torch.int32)
        gemm_ws = None
        if use_tc:
            gemm_ws = ws.get("gemm_ws", (max(K.dot_gemm_workspace(nS, win, W) // 4, 1),),
                             torch.float32)
        n_out = bps * R
        ids_out = torch.empty(n_out * S, self.k, dtype=torch.int32, device=dev)
        sc_out = torch.empty(n_out * S, self.k, dtype=torch.float32, device=dev)
        acc: Dict[str, List] = {}

        for s in range(bps):
            rows = [s] if pl.distributed else [s * n + r for r in pl.shards]
            # ---- queries of every shard, replicated (bess.py:763-769)
            if pl.distributed:
                mine = ws.get("tk_Qmine", (S, W), tdt)
                K.gather_rows(ent[pl.rank], fixed[rows[0]], mine)
                torch.distributed.all_gather_into_tensor(Q.view(-1), mine.view(-1))
                torch.distributed.all_gather_into_tensor(rel_all, rel[rows[0]].contiguous())
            else:
                for li, (row, shard) in enumerate(zip(rows, pl.shards)):
                    K.gather_rows(ent[shard], fixed[row], Q[li * S:(li + 1) * S])
                    rel_all[li * S:(li + 1) * S].copy_(rel[row])
            K.prologue_fwd(cfg, dt, mode, L.rows(Q), rel_table, rel_all, L.IDENT, nS, qv)
            q_op = None
            if use_tc:
                q_op = _TcOperand(ws, "tkq", nS, W, tdt, False)
                q_op.fill(L.F32, L.rows(qv.view(nS, W)), dt, None, dev)
            K.fill_f32(best_s, BAD_NEGATIVE_SCORE)
            K.fill_i32(best_i, Es)
            # ---- score where the candidates live, keep a running top-(k+1)
            for li, (row, shard) in enumerate(zip(rows, pl.shards)):
                table = ent[shard]
                tab_hi = tab_lo = None
                if use_tc and negative is None:
                    # whole-shard operands: the table itself for halves, cached hi / lo for fp32
                    if tdt == torch.float32:
                        tab_hi, tab_lo, tab_ld = self._table_operand(li, table)
                    elif table.stride(0) == _pad8(W):
                        tab_hi, tab_ld = table, table.stride(0)
                for c0 in range(0, n_cand_total, win):
                    nc = min(win, n_cand_total - c0)
                    if negative is None:
                        cand = L.rows(table, offset_elems=c0 * table.stride(0))
                        ids, ld_ids, id0 = None, 0, c0
                        per_query = False
                    elif flat:
                        sel = cand_idx[row, 0, c0:c0 + nc].contiguous()
                        cand = L.rows(table, idx=sel)
                        ids, ld_ids, id0 = sel, 0, 0
                        per_query = False
                    else:
                        sel = cand_idx[row, :, c0:c0 + nc].contiguous()  # [n*S, nc]
                        cand = L.rows(table, idx=sel.view(-1))
                        ids, ld_ids, id0 = sel, nc, 0
                        per_query = True
                    if per_query:
                        K.pertriple_fwd(cfg, dt, mode, qv, nS, cand, nc, nc, scores, L.IDENT, win,
                                        0, aux)
                    elif use_tc and tab_hi is not None:
                        K.dot_gemm(dt, q_op.hi, q_op.lo, q_op.ld, tab_hi[c0:c0 + nc],
                                   None if tab_lo is None else tab_lo[c0:c0 + nc], tab_ld, nS, nc,
                                   W, scores, L.IDENT, win, 0, False, gemm_ws)
                    elif use_tc:
                        c_op = _TcOperand(ws, "tkc", nc, W, tdt, False)
                        c_op.fill(dt, cand, dt, None, dev)
                        K.dot_gemm(dt, q_op.hi, q_op.lo, q_op.ld, c_op.hi, c_op.lo, c_op.ld, nS, nc,
                                   W, scores, L.IDENT, win, 0, False, gemm_ws)
                    else:
                        sc_ = None
                        if need_scale:
                            sc_ = K.cand_scales(cfg, dt, cand, nc, W, scale)
                        K.shared_fwd(cfg, dt, mode, qv, nS, cand, sc_, nc, scores, L.IDENT, win, 0,
                                     aux)
                    if cand_mask is not None:
                        m = (cand_mask[row, 0:1, c0:c0 + nc] if flat
                             else cand_mask[row, :, c0:c0 + nc]).contiguous()  # [1 or n*S, nc]
                        K.mask_add(scores, nS, nc, win, m, nc, m.shape[0], False,
                                   BAD_NEGATIVE_SCORE)
                    K.topk_merge(scores, win, nS, nc, ids, ld_ids, id0, best_s[li], best_i[li], kbm)
                if exact:
                    K.topk_exact_rescore(cfg, dt, mode, Q, rel_table, rel_all, table, best_i[li],
                                         best_s[li], nS, kbm, kb, send_s[li], send_i[li])
            # ---- best lists back to the shard that owns the queries (bess.py:856-863)
            if pl.distributed:
                torch.distributed.all_to_all_single(recv_s.view(-1), send_s.view(-1))
                torch.distributed.all_to_all_single(recv_i.view(-1), send_i.view(-1))
            else:
                recv_s.copy_(send_s.view(R, n, S, kb).transpose(0, 1))
                recv_i.copy_(send_i.view(R, n, S, kb).transpose(0, 1))
            for li in range(R):
                o = s * R + li
                K.topk_finalize(recv_s[li], recv_i[li], n, S, kb, counts_dev, s2e_dev, Es, self.k,
                                BAD_NEGATIVE_SCORE, sc_out[o * S:(o + 1) * S],
                                ids_out[o * S:(o + 1) * S])
                if self.evaluation is not None:
                    assert truth is not None, "Evaluation requires providing ground truth entities"
                    rank = self.evaluation.ranks_from_indices(truth[rows[li]],
                                                              ids_out[o * S:(o + 1) * S])
                    if self.evaluation.return_ranks:
                        acc.setdefault("ranks", []).append(rank)
                    acc.setdefault("metrics", []).append(
                        self.evaluation.stacked_metrics_from_ranks(
                            rank, None if tmask is None else tmask[rows[li]]))

        out: Dict[str, Any] = dict(topk_global_id=ids_out)
        if self.return_scores:
            out["topk_scores"] = sc_out if tdt == torch.float32 else sc_out.to(tdt)
        if "ranks" in acc:
            out["ranks"] = torch.cat(acc["ranks"])
        if "metrics" in acc:
            out["metrics"] = torch.cat(acc["metrics"], dim=0)
        return out


class AllScoresBESS(torch.nn.Module):
    """Scores of (h, r, ?) / (?, r, t) queries against the entities of every
    shard, returned in blocks (reference: bess.py:924-1062).  Queries are
    replicated (AllGather), scored on the shard that stores the candidates and
    the scores travel back to the shard that owns the query (AllToAll).

    `forward(step, relation, head|tail)` is the reference call: block `step` of
    `window_size` local entities per shard, result `[L * S, n_shard * window_size]`
    (column `j * window_size + l` = entity `min(step * window_size + l, Es - 1)`
    of shard j; the clamp repeats the last entity in the final block exactly as
    bess.py:1030-1036 does).  `score_all` is the B200 form the pipeline uses:
    every local entity of every shard in one pass through the tcgen05 GEMM
    (DistMult / ComplEx) or the tile scorers, `[L * S, n_shard * Es]`.
    Inference only."""

    device_window = 4096

    def __init__(self, candidate_sampler: PlaceholderNegativeSampler,
                 score_fn: BaseScoreFunction, window_size: int = 1000) -> None:
        super().__init__()
        self.sharding = score_fn.sharding
        self.score_fn = score_fn
        self.negative_sampler = candidate_sampler
        self.window_size = window_size
        if not score_fn.negative_sample_sharing:
            raise ValueError("AllScoresBESS requires using negative sample sharing")
        if self.negative_sampler.corruption_scheme not in ["h", "t"]:
            raise ValueError("AllScoresBESS only support 'h', 't' corruption scheme")
        if not isinstance(self.negative_sampler, PlaceholderNegativeSampler):
            raise ValueError(
                "AllScoresBESS requires a `PlaceholderNegativeSampler` candidate_sampler"
            )
        self.entity_embedding = self.score_fn.entity_embedding
        self.entity_embedding_size: int = self.entity_embedding.shape[-1]
        self.candidate = torch.arange(self.window_size, dtype=torch.int32)
        self.n_step = int(np.ceil(self.sharding.max_entity_per_shard / self.window_size))
        # the windowed scorer (query replication, prologue, GEMM / tile kernels, cached
        # table operands) is the one TopKQueryBessKGE runs; only the reduction differs
        self._engine = TopKQueryBessKGE(1, candidate_sampler, score_fn, None, False, window_size)

    def _queries(self, relation: torch.Tensor, head: Optional[torch.Tensor],
                 tail: Optional[torch.Tensor]):
        scheme = self.negative_sampler.corruption_scheme
        fixed = head if scheme == "t" else tail
        if fixed is None:
            raise ValueError("queries need the known entity: head for 't', tail for 'h'")
        L_rows, S = relation.shape[0], relation.shape[-1]
        n = self.sharding.n_shard
        if L_rows % n != 0:
            raise ValueError(f"leading axis {L_rows} is not a multiple of n_shard={n}")
        return fixed.reshape(L_rows, S), relation.reshape(L_rows, S), L_rows, S

    def _score(self, relation: torch.Tensor, fixed_h: torch.Tensor, L_rows: int, S: int,
               windows: List[Tuple[int, int]], clamp_to: int) -> torch.Tensor:
        """Scores of all queries against local entities [c0, c0 + width) of every shard, for
        each (c0, width) in `windows`; rows beyond the shard (>= Es) repeat row Es - 1 when
        `clamp_to` > 0.  Returns [bps * R * S, n * sum(width)] on the device (R = local shards)."""
        eng = self._engine
        ws, pl = eng._setup()
        dev = ws.device
        ent = self.score_fn.entity_embedding.data
        rel_table = self.score_fn.relation_embedding.data
        n = self.sharding.n_shard
        Es, W = ent.shape[1], ent.shape[-1]
        cfg = self.score_fn.kernel_cfg()
        tdt = ent.dtype
        dt = L.dtype_code(tdt)
        mode = L.MODE_TAILS if self.negative_sampler.corruption_scheme == "t" else L.MODE_HEADS
        bps = L_rows // n
        nS = n * S
        R = pl.n_local
        total_w = sum(w for _, w in windows)

        def to_dev(t):
            if pl.distributed:
                t = t[pl.rank::n]
            return t.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous()

        rel = to_dev(relation)
        fixed = to_dev(fixed_h)
        nvec = K.call("bess_query_nvec", L.C.byref(cfg))
        use_tc = USE_TENSOR_CORES and cfg.family in (L.DISTMULT, L.COMPLEX)
        need_aux = cfg.family == L.BOXE and cfg.norm_p == 2
        need_scale = K.needs_cand_scale(cfg)
        max_w = max(min(self.device_window, w) for _, w in windows)
        Q = ws.get("as_Q", (nS, W), tdt)
        rel_all = ws.get("as_rel", (nS,), torch.int32)
        qv = ws.get("as_qv", (nS, nvec, W), torch.float32)
        aux = ws.get("as_aux", (nS, _pad8(max_w)), torch.float32) if need_aux else None
        scale = ws.get("as_scale", (2 * max_w,), torch.float32) if need_scale else None
        gemm_ws = None
        if use_tc:
            gemm_ws = ws.get("gemm_ws", (max(K.dot_gemm_workspace(nS, _pad8(max_w), W) // 4, 1),),
                             torch.float32)
        # local mode: scoring shard r writes column block r of every query row directly;
        # distributed: [n*S, total_w] local scores, AllToAll, transposed copy to [S, n, total_w]
        ld_out = n * total_w
        out = torch.empty(bps * R * S, ld_out, dtype=torch.float32, device=dev)
        if pl.distributed:
            sc_local = ws.get("as_local", (nS, total_w), torch.float32)
            sc_recv = ws.get("as_recv", (n, S, total_w), torch.float32)

        for s in range(bps):
            rows = [s] if pl.distributed else [s * n + r for r in pl.shards]
            if pl.distributed:
                mine = ws.get("as_Qmine", (S, W), tdt)
                K.gather_rows(ent[pl.rank], fixed[rows[0]], mine)
                torch.distributed.all_gather_into_tensor(Q.view(-1), mine.view(-1))
                torch.distributed.all_gather_into_tensor(rel_all, rel[rows[0]].contiguous())
            else:
                for li, (row, shard) in enumerate(zip(rows, pl.shards)):
                    K.gather_rows(ent[shard], fixed[row], Q[li * S:(li + 1) * S])
                    rel_all[li * S:(li + 1) * S].copy_(rel[row])
            K.prologue_fwd(cfg, dt, mode, L.rows(Q), rel_table, rel_all, L.IDENT, nS, qv)
            q_op = None
            if use_tc:
                q_op = _TcOperand(ws, "asq", nS, W, tdt, False)
                q_op.fill(L.F32, L.rows(qv.view(nS, W)), dt, None, dev)
            for li, shard in enumerate(pl.shards):
                table = ent[shard]
                if pl.distributed:
                    dst, ld, base_col = sc_local, total_w, 0
                else:
                    dst, ld, base_col = out[s * nS:(s + 1) * nS], ld_out, shard * total_w
                tab_hi = tab_lo = None
                tab_ld = 0
                if use_tc:
                    if tdt == torch.float32:
                        tab_hi, tab_lo, tab_ld = eng._table_operand(li, table)
                    elif table.stride(0) == _pad8(W):
                        tab_hi, tab_ld = table, table.stride(0)
                col = base_col
                for c_start, width in windows:
                    for c0 in range(c_start, c_start + width, self.device_window):
                        nc = min(self.device_window, c_start + width - c0)
                        real = max(0, min(nc, Es - c0))  # columns that are real local rows
                        if real > 0:
                            cand = L.rows(table, offset_elems=c0 * table.stride(0))
                            if use_tc and tab_hi is not None:
                                K.dot_gemm(dt, q_op.hi, q_op.lo, q_op.ld, tab_hi[c0:c0 + real],
                                           None if tab_lo is None else tab_lo[c0:c0 + real], tab_ld,
                                           nS, real, W, dst, L.IDENT, ld, col, False, gemm_ws)
                            elif use_tc:
                                c_op = _TcOperand(ws, "asc", real, W, tdt, False)
                                c_op.fill(dt, cand, dt, None, dev)
                                K.dot_gemm(dt, q_op.hi, q_op.lo, q_op.ld, c_op.hi, c_op.lo, c_op.ld,
                                           nS, real, W, dst, L.IDENT, ld, col, False, gemm_ws)
                            else:
                                sc_ = None
                                if need_scale:
                                    sc_ = K.cand_scales(cfg, dt, cand, real, W, scale)
                                K.shared_fwd(cfg, dt, mode, qv, nS, cand, sc_, real, dst, L.IDENT, ld,
                                             col, aux)
                        if real < nc:  # clamped tail of the last block: entity Es - 1 again
                            assert clamp_to > 0
                            last = dst[:, col + real - 1:col + real] if real > 0 else None
                            if last is None:
                                last = self._last_entity_scores(table, Es, cfg, dt, mode, qv, nS, ws,
                                                                need_scale, aux)
                            dst[:, col + real:col + nc] = last
                        col += nc
            if pl.distributed:
                torch.distributed.all_to_all_single(sc_recv.view(-1), sc_local.view(-1))
                out[s * S:(s + 1) * S].view(S, n, total_w).copy_(sc_recv.transpose(0, 1))
        return out

    def _last_entity_scores(self, table, Es, cfg, dt, mode, qv, nS, ws, need_scale, aux):
        """[nS, 1] scores against local entity Es - 1 (a block that lies entirely beyond the
        shard; only reachable when window_size does not divide into the shard evenly)."""
        cand = L.rows(table, offset_elems=(Es - 1) * table.stride(0))
        one = ws.get("as_last", (nS, 8), torch.float32)
        sc_ = None
        if need_scale:
            sc_ = K.cand_scales(cfg, dt, cand, 1, table.shape[-1],
                                ws.get("as_scale1", (2,), torch.float32))
        K.shared_fwd(cfg, dt, mode, qv, nS, cand, sc_, 1, one, L.IDENT, 8, 0, aux)
        return one[:, 0:1]

    def forward(self, step: torch.Tensor, relation: torch.Tensor,
                head: Optional[torch.Tensor] = None,
                tail: Optional[torch.Tensor] = None) -> torch.Tensor:
        """relation / head / tail [L, S] (L = bps * n_shard); step: the block index (the same
        for every row, bess.py:984-1062).  Returns [L * S, n_shard * window_size] (this rank's
        rows in distributed mode)."""
        fixed, rel, L_rows, S = self._queries(relation, head, tail)
        i = int(torch.as_tensor(step).reshape(-1)[0])
        return self._score(rel, fixed, L_rows, S, [(i * self.window_size, self.window_size)],
                           clamp_to=self.sharding.max_entity_per_shard)

    def score_all(self, relation: torch.Tensor, head: Optional[torch.Tensor] = None,
                  tail: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Every block at once: [L * S, n_shard * Es], column j * Es + l = local entity l of
        shard j (padding rows of a shard included; the caller maps columns to global ids)."""
        fixed, rel, L_rows, S = self._queries(relation, head, tail)
        Es = self.sharding.max_entity_per_shard
        return self._score(rel, fixed, L_rows, S, [(0, Es)], clamp_to=0)


# ---------------------------------------------------------------------------
# Names used by BASELINE.json's north_star.  The reference has no classes called
# `ScatterAllToAllBessKGE` / `AllScatterAllGatherBessKGE` (its distribution schemes are
# EmbeddingMovingBessKGE, bess.py:308, and ScoreMovingBessKGE, bess.py:471 — SURVEY.md §0);
# the names describe those two schemes by their collectives and are exported as plain aliases,
# with no semantics of their own:
#   * ScatterAllToAll       — gathered tail / negative rows are scattered to the shard that
#                             scores them with one balanced AllToAll  = EmbeddingMovingBessKGE;
#   * AllScatterAllGather   — queries are AllGathered, negatives are scored where they are
#                             stored and the scores travel back       = ScoreMovingBessKGE.
# ---------------------------------------------------------------------------
ScatterAllToAllBessKGE = EmbeddingMovingBessKGE
AllScatterAllGatherBessKGE = ScoreMovingBessKGE
