"""Pythonic launchers over the C-ABI (`_lib`).  Every function enqueues CUDA
work on torch's current stream of the tensors' device and returns immediately.
Inputs must be CUDA tensors; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib as L
from ._lib import IDENT, RowMap, Rows, ScoreCfg, call, dtype_code, ptr, require_cuda, rowmap, rows

BAD_NEGATIVE_SCORE = -50000.0


def _st(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


class Workspace:
    """Persistent device buffers keyed by name (grown on demand, never shrunk),
    so the steady-state step performs no allocation."""

    def __init__(self, device: torch.device) -> None:
        self.device = device
        self._buf: Dict[str, torch.Tensor] = {}
        # `generation` moves whenever an existing buffer is REPLACED (grown or re-typed): raw
        # pointers baked into a captured CUDA graph then refer to the old buffer.  Replaced
        # buffers are parked in `_retired` (still allocated, so a stale replay cannot touch
        # freed memory) until the owner of the graphs has dropped them (`release_retired`).
        self.generation = 0
        self._retired: List[torch.Tensor] = []

    def get(self, name: str, shape: Sequence[int], dtype: torch.dtype) -> torch.Tensor:
        n = int(math.prod(shape)) if len(shape) else 1
        cur = self._buf.get(name)
        if cur is None or cur.dtype != dtype or cur.numel() < n:
            if cur is not None:
                self._retired.append(cur)
                self.generation += 1
            cur = torch.empty(max(n, 1), dtype=dtype, device=self.device)
            self._buf[name] = cur
        return cur[:n].view(*shape)

    def get_zeroed(self, name: str, shape: Sequence[int], dtype: torch.dtype) -> torch.Tensor:
        """Like `get`, but the buffer is zero-filled when it is first created (kernel-owned
        state words that the kernels themselves leave zero afterwards)."""
        fresh = name not in self._buf
        t = self.get(name, shape, dtype)
        if fresh or self._buf[name].data_ptr() != getattr(self, "_zeroed", {}).get(name):
            self._buf[name].zero_()
            self.__dict__.setdefault("_zeroed", {})[name] = self._buf[name].data_ptr()
        return t

    def release_retired(self) -> None:
        """Free replaced buffers; call only when no captured graph refers to them."""
        self._retired.clear()

    def bytes(self) -> int:
        return sum(b.numel() * b.element_size() for b in self._buf.values())


# ------------------------------------------------------------------ gather --
def gather_route(table: torch.Tensor, idx: torch.Tensor, n_local: int, per_dst: int,
                 local_out: Optional[torch.Tensor], dst_ptrs: Sequence[int], slot: int) -> None:
    """table [Es, W]; idx int32 [n_local + len(dst_ptrs) * per_dst]."""
    require_cuda(table, idx)
    n_dst = len(dst_ptrs)
    arr = (C.c_void_p * max(n_dst, 1))(*[C.c_void_p(p) for p in dst_ptrs])
    call("bess_gather_route", table.data_ptr(), table.stride(0), dtype_code(table.dtype),
         table.shape[1], idx.data_ptr(), n_local, n_dst, per_dst, ptr(local_out), arr, slot,
         _st(table))


def gather_rows(table: torch.Tensor, idx: torch.Tensor, out: torch.Tensor) -> None:
    require_cuda(table, idx, out)
    call("bess_gather_rows", table.data_ptr(), table.stride(0), dtype_code(table.dtype),
         table.shape[1], idx.data_ptr(), idx.numel(), out.data_ptr(), _st(table))


# ------------------------------------------------------------- score ops ----
def triple_fwd(cfg: ScoreCfg, dt: int, head: Rows, tail: Rows, rel_table: torch.Tensor,
               rel_id: torch.Tensor, rel_map: RowMap, n: int, score: torch.Tensor,
               score_map: RowMap) -> None:
    call("bess_score_triple_fwd", C.byref(cfg), dt, head, tail, rel_table.data_ptr(),
         rel_id.data_ptr(), rel_map, n, score.data_ptr(), score_map, _st(score))


def triple_bwd(cfg: ScoreCfg, dt: int, head: Rows, tail: Rows, rel_table: torch.Tensor,
               rel_id: torch.Tensor, rel_map: RowMap, n: int, score: torch.Tensor,
               d_score: torch.Tensor, score_map: RowMap, d_head: Rows, d_tail: Rows,
               d_rel: torch.Tensor, add_head: bool, add_tail: bool, add_rel: bool) -> None:
    call("bess_score_triple_bwd", C.byref(cfg), dt, head, tail, rel_table.data_ptr(),
         rel_id.data_ptr(), rel_map, n, score.data_ptr(), d_score.data_ptr(), score_map, d_head,
         d_tail, d_rel.data_ptr(), int(add_head), int(add_tail), int(add_rel), _st(score))


def prologue_fwd(cfg: ScoreCfg, dt: int, mode: int, fixed: Rows, rel_table: torch.Tensor,
                 rel_id: torch.Tensor, rel_map: RowMap, n: int, qv: torch.Tensor) -> None:
    call("bess_query_prologue_fwd", C.byref(cfg), dt, mode, fixed, rel_table.data_ptr(),
         rel_id.data_ptr(), rel_map, n, qv.data_ptr(), _st(qv))


def prologue_bwd(cfg: ScoreCfg, dt: int, mode: int, fixed: Rows, rel_table: torch.Tensor,
                 rel_id: torch.Tensor, rel_map: RowMap, n: int, d_qv: torch.Tensor, d_fixed: Rows,
                 d_rel: torch.Tensor, add_fixed: bool, add_rel: bool) -> None:
    call("bess_query_prologue_bwd", C.byref(cfg), dt, mode, fixed, rel_table.data_ptr(),
         rel_id.data_ptr(), rel_map, n, d_qv.data_ptr(), d_fixed, d_rel.data_ptr(), int(add_fixed),
         int(add_rel), _st(d_qv))


def boxe_rel_finalize(cfg: ScoreCfg, dt: int, rel_table: torch.Tensor, rel_id: torch.Tensor,
                      n: int, d_rel: torch.Tensor) -> None:
    call("bess_boxe_rel_finalize", C.byref(cfg), dt, rel_table.data_ptr(), rel_id.data_ptr(), n,
         d_rel.data_ptr(), _st(d_rel))


def cand_inv_norm(dt: int, cand: Rows, n: int, width: int, out: torch.Tensor) -> None:
    call("bess_cand_inv_norm", dt, cand, n, width, out.data_ptr(), _st(out))


def cand_norm_bwd(dt: int, cand: Rows, n: int, width: int, inv_norm: torch.Tensor,
                  d_cand: Rows) -> None:
    call("bess_cand_norm_bwd", dt, cand, n, width, inv_norm.data_ptr(), d_cand, _st(inv_norm))


# families whose candidates are L2-normalised before scoring: PairRE / TripleRE normalise the
# whole row (one inverse norm per candidate), InterHT / TranS the two halves [main | aux]
# separately (two inverse norms per candidate, stored [main norms | aux norms])
def needs_cand_scale(cfg: ScoreCfg) -> bool:
    return bool(cfg.normalize) and cfg.family in (L.PAIRRE, L.TRIPLERE, L.INTERHT, L.TRANS)


def cand_scale_len(cfg: ScoreCfg, n: int) -> int:
    return 2 * n if cfg.family in (L.INTERHT, L.TRANS) else n


def _shift_rows(r: Rows, elems: int, elem_bytes: int) -> Rows:
    return Rows(r.base + elems * elem_bytes, r.idx, r.map, r.pitch)


def cand_scales(cfg: ScoreCfg, dt: int, cand: Rows, n: int, width: int,
                out: torch.Tensor) -> torch.Tensor:
    """inverse norms of the n candidate rows -> out[:cand_scale_len(cfg, n)] (returned)."""
    out = out[:cand_scale_len(cfg, n)]
    if cfg.family in (L.INTERHT, L.TRANS):
        d, es = width // 2, (4 if dt == L.F32 else 2)
        cand_inv_norm(dt, cand, n, d, out[:n])
        cand_inv_norm(dt, _shift_rows(cand, d, es), n, d, out[n:])
    else:
        cand_inv_norm(dt, cand, n, width, out)
    return out


def cand_scales_bwd(cfg: ScoreCfg, dt: int, cand: Rows, n: int, width: int, scale: torch.Tensor,
                    d_cand: Rows) -> None:
    """chain rule of the candidate normalisation, in place on the fp32 gradient rows."""
    if cfg.family in (L.INTERHT, L.TRANS):
        d, es = width // 2, (4 if dt == L.F32 else 2)
        cand_norm_bwd(dt, cand, n, d, scale[:n], d_cand)
        cand_norm_bwd(dt, _shift_rows(cand, d, es), n, d, scale[n:], _shift_rows(d_cand, d, 4))
    else:
        cand_norm_bwd(dt, cand, n, width, scale, d_cand)


def shared_fwd(cfg: ScoreCfg, dt: int, mode: int, qv: torch.Tensor, n_query: int, cand: Rows,
               cand_scale: Optional[torch.Tensor], n_cand: int, out: torch.Tensor,
               score_map: RowMap, ld: int, col0: int, aux: Optional[torch.Tensor],
               out_ptr: Optional[int] = None) -> None:
    """`out_ptr`: raw device address of the score matrix instead of `out` (a peer GPU's buffer)."""
    call("bess_score_shared_fwd", C.byref(cfg), dt, mode, qv.data_ptr(), n_query, cand,
         ptr(cand_scale), n_cand, out.data_ptr() if out_ptr is None else out_ptr, score_map, ld,
         col0, ptr(aux), _st(qv))


def shared_bwd_query(cfg: ScoreCfg, dt: int, mode: int, qv: torch.Tensor, n_query: int, cand: Rows,
                     cand_scale: Optional[torch.Tensor], n_cand: int, score: torch.Tensor,
                     d_score: torch.Tensor, score_map: RowMap, ld: int, col0: int,
                     aux: Optional[torch.Tensor], d_qv: torch.Tensor) -> None:
    call("bess_score_shared_bwd_query", C.byref(cfg), dt, mode, qv.data_ptr(), n_query, cand,
         ptr(cand_scale), n_cand, score.data_ptr(), d_score.data_ptr(), score_map, ld, col0,
         ptr(aux), d_qv.data_ptr(), _st(d_qv))


def shared_bwd_cand_workspace(cfg: ScoreCfg, n_query: int, n_cand: int) -> int:
    return int(call("bess_shared_bwd_cand_workspace", C.byref(cfg), n_query, n_cand))


def shared_bwd_cand(cfg: ScoreCfg, dt: int, mode: int, qv: torch.Tensor, n_query: int, cand: Rows,
                    cand_scale: Optional[torch.Tensor], n_cand: int, score: torch.Tensor,
                    d_score: torch.Tensor, score_map: RowMap, ld: int, col0: int,
                    aux: Optional[torch.Tensor], d_cand: Rows, workspace: torch.Tensor,
                    add: bool = False) -> None:
    call("bess_score_shared_bwd_cand", C.byref(cfg), dt, mode, qv.data_ptr(), n_query, cand,
         ptr(cand_scale), n_cand, score.data_ptr(), d_score.data_ptr(), score_map, ld, col0,
         ptr(aux), d_cand, int(add), workspace.data_ptr(), _st(score))


# ------------------------------------------------ tensor-core DOT path ------
def dot_gemm_workspace(m: int, n: int, k: int) -> int:
    return int(call("bess_dot_gemm_workspace", m, n, k))


def split_operand(src_dt: int, src: Rows, n_rows: int, width: int,
                  row_scale: Optional[torch.Tensor], out_dt: int,
                  hi: Optional[torch.Tensor], lo: Optional[torch.Tensor], ld: int,
                  hi_t: Optional[torch.Tensor], lo_t: Optional[torch.Tensor], ld_t: int,
                  device: torch.device, op_scale: Optional[torch.Tensor] = None) -> None:
    """rows -> dense K-major GEMM operand(s); see bess_split_operand."""
    call("bess_split_operand", src_dt, src, n_rows, width, ptr(row_scale), out_dt, ptr(hi), ptr(lo),
         ld, ptr(hi_t), ptr(lo_t), ld_t, ptr(op_scale),
         torch.cuda.current_stream(device).cuda_stream)


def operand_scale(src_dt: int, src: Rows, n_rows: int, width: int,
                  row_scale: Optional[torch.Tensor], factor: float, scale: torch.Tensor,
                  state: torch.Tensor) -> None:
    """{s, 1 / s} of a 3xFP16 operand from factor * max |x|; see bess_operand_scale."""
    call("bess_operand_scale", src_dt, src, n_rows, width, ptr(row_scale), float(factor),
         scale.data_ptr(), state.data_ptr(), _st(scale))


# ------------------------------------------ norm-expanded L2 (tensor cores) --
def row_sqnorm(dt: int, rows_: Rows, n: int, width: int, out: torch.Tensor) -> None:
    call("bess_row_sqnorm", dt, rows_, n, width, out.data_ptr(), _st(out))


def l2_from_dot(score: torch.Tensor, score_map: RowMap, ld: int, col0: int, n_query: int,
                n_cand: int, qn: torch.Tensor, cn: torch.Tensor) -> None:
    """dots -> -sqrt(max(qn + cn - 2 dot, 0)) in place; see bess_l2_from_dot."""
    call("bess_l2_from_dot", score.data_ptr(), score_map, ld, col0, n_query, n_cand, qn.data_ptr(),
         cn.data_ptr(), _st(score))


def l2_coef_workspace(n_query: int, n_cand: int) -> int:
    return int(call("bess_l2_coef_workspace", n_query, n_cand))


def l2_coef(d_score: torch.Tensor, score: torch.Tensor, score_map: RowMap, ld: int, col0: int,
            n_query: int, n_cand: int, coef: torch.Tensor, ld_coef: int, row_sum: torch.Tensor,
            col_sum: torch.Tensor, workspace: torch.Tensor) -> None:
    call("bess_l2_coef", d_score.data_ptr(), score.data_ptr(), score_map, ld, col0, n_query, n_cand,
         coef.data_ptr(), ld_coef, row_sum.data_ptr(), col_sum.data_ptr(), workspace.data_ptr(),
         _st(score))


def rows_axpy(dt: int, alpha: torch.Tensor, scale: float, src: Rows, out: Rows, n: int,
              width: int) -> None:
    """out_i += scale * alpha[i] * src_i; see bess_rows_axpy."""
    call("bess_rows_axpy", dt, alpha.data_ptr(), float(scale), src, out, n, width, _st(alpha))


def table_operand_refresh(table: torch.Tensor, hi: torch.Tensor, lo: torch.Tensor, ld: int,
                          state: torch.Tensor, force: bool) -> None:
    """fp32 table [rows, W] -> cached hi / lo operand arrays, rebuilt only when the table's
    bytes changed (device-side checksum compare); see bess_table_operand_refresh."""
    require_cuda(table, hi, lo, state)
    call("bess_table_operand_refresh", table.data_ptr(), table.shape[0], table.shape[1],
         table.stride(0), hi.data_ptr(), lo.data_ptr(), ld, state.data_ptr(), int(force), _st(table))


def dot_gemm(dt: int, a_hi: torch.Tensor, a_lo: Optional[torch.Tensor], lda: int,
             b_hi: torch.Tensor, b_lo: Optional[torch.Tensor], ldb: int, m: int, n: int, k: int,
             out: torch.Tensor, out_map: RowMap, ld_out: int, col0: int, accumulate: bool,
             workspace: Optional[torch.Tensor], out_ptr: Optional[int] = None,
             a_mn_major: bool = False, a_offset_elems: int = 0,
             a_scale: Optional[torch.Tensor] = None, b_scale: Optional[torch.Tensor] = None) -> None:
    """out[out_map(i) * ld_out + col0 + j] (+)= sum_k A[i, k] * B[j, k] on the tcgen05 path.
    a_mn_major: A is stored transposed [K, M] (M contiguous, leading dimension lda)."""
    ws_bytes = 0 if workspace is None else workspace.numel() * workspace.element_size()
    off = a_offset_elems * a_hi.element_size()
    call("bess_dot_gemm", dt, a_hi.data_ptr() + off, None if a_lo is None else a_lo.data_ptr() + off,
         lda, int(a_mn_major), b_hi.data_ptr(), ptr(b_lo), ldb, m, n, k,
         out.data_ptr() if out_ptr is None else out_ptr, out_map, ld_out, col0, int(accumulate),
         ptr(workspace), ws_bytes, ptr(a_scale), ptr(b_scale), _st(a_hi))


def pertriple_fwd(cfg: ScoreCfg, dt: int, mode: int, qv: torch.Tensor, n_query: int, cand: Rows,
                  q_stride: int, n_per: int, out: torch.Tensor, score_map: RowMap, ld: int,
                  col0: int, aux: Optional[torch.Tensor], out_ptr: Optional[int] = None) -> None:
    call("bess_score_pertriple_fwd", C.byref(cfg), dt, mode, qv.data_ptr(), n_query, cand,
         q_stride, n_per, out.data_ptr() if out_ptr is None else out_ptr, score_map, ld, col0,
         ptr(aux), _st(qv))


def pertriple_bwd(cfg: ScoreCfg, dt: int, mode: int, qv: torch.Tensor, n_query: int, cand: Rows,
                  q_stride: int, n_per: int, score: torch.Tensor, d_score: torch.Tensor,
                  score_map: RowMap, ld: int, col0: int, aux: Optional[torch.Tensor],
                  d_qv: torch.Tensor, d_cand: Rows) -> None:
    call("bess_score_pertriple_bwd", C.byref(cfg), dt, mode, qv.data_ptr(), n_query, cand,
         q_stride, n_per, score.data_ptr(), d_score.data_ptr(), score_map, ld, col0, ptr(aux),
         d_qv.data_ptr(), d_cand, _st(d_qv))


# ---------------------------------------------------------- masks / loss ----
def mask_add(score: torch.Tensor, n_row: int, n_col: int, ld: int, mask: torch.Tensor,
             ld_mask: int, mask_rows: int, flag: bool, value: float, col_offset: int = 0) -> None:
    """score[r, col_offset + c] += value where (mask[r or 0, c] != 0) == flag."""
    call("bess_mask_add", score.data_ptr() + 4 * col_offset, n_row, n_col, ld, mask.data_ptr(),
         ld_mask, mask_rows, int(flag), float(value), _st(score))


def mask_diag(score: torch.Tensor, n_row: int, ld: int, step: int, half_group: int, group: int,
              value: float) -> None:
    call("bess_mask_diag", score.data_ptr(), n_row, ld, step, half_group, group, float(value),
         _st(score))


def loss_fwd_bwd(kind: int, margin: float, adversarial: bool, adv_scale: float, loss_scale: float,
                 n_entity: int, pos: torch.Tensor, neg: torch.Tensor, n: int, n_neg: int, ld: int,
                 weight: torch.Tensor, row_loss: torch.Tensor, d_pos: torch.Tensor,
                 d_neg: torch.Tensor) -> None:
    call("bess_loss_fwd_bwd", kind, float(margin), int(adversarial), float(adv_scale),
         float(loss_scale), int(n_entity), pos.data_ptr(), neg.data_ptr(), n, n_neg, ld,
         weight.data_ptr(), weight.numel(), row_loss.data_ptr(), d_pos.data_ptr(),
         d_neg.data_ptr(), _st(pos))


def loss_fwd_bwd_operand(kind: int, margin: float, adversarial: bool, adv_scale: float,
                         loss_scale: float, n_entity: int, pos: torch.Tensor, neg: torch.Tensor,
                         n: int, n_neg: int, ld: int, weight: torch.Tensor, row_loss: torch.Tensor,
                         d_pos: torch.Tensor, grad_dt: int, d_neg_hi: torch.Tensor,
                         d_neg_lo: Optional[torch.Tensor], ld_grad: int,
                         grad_scale: Optional[torch.Tensor] = None) -> None:
    """loss + dL/dscore with the [n, n_neg] gradient written as GEMM operand arrays."""
    call("bess_loss_fwd_bwd_operand", kind, float(margin), int(adversarial), float(adv_scale),
         float(loss_scale), int(n_entity), pos.data_ptr(), neg.data_ptr(), n, n_neg, ld,
         weight.data_ptr(), weight.numel(), row_loss.data_ptr(), d_pos.data_ptr(), grad_dt,
         d_neg_hi.data_ptr(), ptr(d_neg_lo), ld_grad, ptr(grad_scale), _st(pos))


def sum_f32(x: torch.Tensor, n: int, out: torch.Tensor) -> None:
    call("bess_sum_f32", x.data_ptr(), n, out.data_ptr(), _st(x))


def rank_from_scores(pos: torch.Tensor, neg: torch.Tensor, n: int, n_neg: int, ld: int, mode: int,
                     worst_inf: bool, rank: torch.Tensor) -> None:
    call("bess_rank_from_scores", pos.data_ptr(), neg.data_ptr(), n, n_neg, ld, mode,
         int(worst_inf), rank.data_ptr(), _st(pos))


# ------------------------------------------------------- sort / scatter -----
def sort_workspace(n: int) -> int:
    return int(call("bess_sort_workspace", n))


def sort_keys(keys: torch.Tensor, n: int, key_bits: int, keys_out: torch.Tensor,
              perm_out: torch.Tensor, workspace: torch.Tensor) -> None:
    call("bess_sort_keys", keys.data_ptr(), n, key_bits, keys_out.data_ptr(), perm_out.data_ptr(),
         workspace.data_ptr(), _st(keys))


def scatter_sgd(table: torch.Tensor, sorted_keys: torch.Tensor, perm: torch.Tensor, n: int,
                n_local: int, per_dst: int, grad_local: torch.Tensor, grad_dst_ptr: int,
                dst_stride_rows: int, lr: float, hyper: Optional[torch.Tensor] = None) -> None:
    call("bess_scatter_sgd", table.data_ptr(), table.stride(0), dtype_code(table.dtype),
         table.shape[1], sorted_keys.data_ptr(), perm.data_ptr(), n, n_local, per_dst,
         grad_local.data_ptr(), grad_dst_ptr, dst_stride_rows, float(lr), ptr(hyper), _st(table))


def scatter_collect(row_elems: int, sorted_keys: torch.Tensor, perm: torch.Tensor, n: int,
                    n_local: int, per_dst: int, grad_local: torch.Tensor, grad_dst_ptr: int,
                    dst_stride_rows: int, seg_grad: torch.Tensor, row_to_seg: torch.Tensor) -> None:
    call("bess_scatter_collect", row_elems, sorted_keys.data_ptr(), perm.data_ptr(), n, n_local,
         per_dst, grad_local.data_ptr(), grad_dst_ptr, dst_stride_rows, seg_grad.data_ptr(),
         row_to_seg.data_ptr(), _st(seg_grad))


def opt_dense(kind: int, table: torch.Tensor, seg_grad: torch.Tensor,
              row_to_seg: Optional[torch.Tensor], state0: Optional[torch.Tensor],
              state1: Optional[torch.Tensor], lr: float, momentum: float, dampening: float,
              beta1: float, beta2: float, eps: float, weight_decay: float, step: int,
              hyper: Optional[torch.Tensor] = None, grad_scale: float = 1.0,
              zero_grad: bool = False) -> None:
    call("bess_opt_dense", kind, table.data_ptr(), table.stride(0), dtype_code(table.dtype),
         table.shape[0], table.shape[1], seg_grad.data_ptr(), ptr(row_to_seg), ptr(state0),
         ptr(state1), float(lr), float(momentum), float(dampening), float(beta1), float(beta2),
         float(eps), float(weight_decay), int(step), ptr(hyper), float(grad_scale),
         int(zero_grad), _st(table))


def scatter_accumulate(row_elems: int, sorted_keys: torch.Tensor, perm: torch.Tensor, n: int,
                       n_local: int, per_dst: int, grad_local: torch.Tensor, grad_dst_ptr: int,
                       dst_stride_rows: int, acc: torch.Tensor) -> None:
    call("bess_scatter_accumulate", row_elems, sorted_keys.data_ptr(), perm.data_ptr(), n, n_local,
         per_dst, grad_local.data_ptr(), grad_dst_ptr, dst_stride_rows, acc.data_ptr(), _st(acc))


def set_hyper(hyper: torch.Tensor, lr: float, momentum: float, dampening: float, beta1: float,
              beta2: float, eps: float, weight_decay: float, step: int) -> None:
    """Write one optimizer step's hyper-parameters (and Adam bias corrections of `step`) into
    the device array the update kernels read; stream-ordered, safe to call every step."""
    call("bess_set_hyper", hyper.data_ptr(), float(lr), float(momentum), float(dampening),
         float(beta1), float(beta2), float(eps), float(weight_decay), int(step), _st(hyper))


def relation_grad_reduce(rows_: torch.Tensor, width: int, sorted_rel: torch.Tensor,
                         perm: torch.Tensor, n: int, n_rel: int, d_table: torch.Tensor) -> None:
    call("bess_relation_grad_reduce", rows_.data_ptr(), width, sorted_rel.data_ptr(),
         perm.data_ptr(), n, n_rel, d_table.data_ptr(), _st(d_table))


def topk_merge(win_score: torch.Tensor, ld: int, n_query: int, n_win: int,
               win_ids: Optional[torch.Tensor], ld_ids: int, win_id0: int,
               best_score: torch.Tensor, best_id: torch.Tensor, k: int) -> None:
    call("bess_topk_merge", win_score.data_ptr(), ld, n_query, n_win, ptr(win_ids), ld_ids,
         win_id0, best_score.data_ptr(), best_id.data_ptr(), k, _st(win_score))


def topk_exact_supported(family: int) -> bool:
    return bool(call("bess_topk_exact_supported", int(family)))


def topk_exact_rescore(cfg: ScoreCfg, dt: int, mode: int, fixed: torch.Tensor,
                       rel_table: torch.Tensor, rel_id: torch.Tensor, table: torch.Tensor,
                       ids_in: torch.Tensor, score_in: torch.Tensor, n_query: int, k_in: int,
                       k_out: int, score_out: torch.Tensor, ids_out: torch.Tensor) -> None:
    """Fixed-order fp32 re-score + re-sort of the best lists; see bess_topk_exact_rescore."""
    require_cuda(fixed, rel_table, rel_id, table, ids_in, score_in, score_out, ids_out)
    call("bess_topk_exact_rescore", C.byref(cfg), dt, mode, fixed.data_ptr(), fixed.stride(0),
         rel_table.data_ptr(), rel_table.stride(0), rel_id.data_ptr(), table.data_ptr(),
         table.stride(0), table.shape[0], table.shape[1], ids_in.data_ptr(), score_in.data_ptr(),
         n_query, k_in, k_out, score_out.data_ptr(), ids_out.data_ptr(), _st(table))


def topk_finalize(score: torch.Tensor, idx: torch.Tensor, n_shard: int, n_query: int, kb: int,
                  shard_counts: torch.Tensor, shard_idx_to_entity: torch.Tensor, es: int, k: int,
                  bad: float, out_score: torch.Tensor, out_id: torch.Tensor) -> None:
    call("bess_topk_finalize", score.data_ptr(), idx.data_ptr(), n_shard, n_query, kb,
         shard_counts.data_ptr(), shard_idx_to_entity.data_ptr(), es, k, float(bad),
         out_score.data_ptr(), out_id.data_ptr(), _st(score))


def select_scores(src: torch.Tensor, row_idx: Optional[torch.Tensor], n_rows: int,
                  col_idx: torch.Tensor, fill: float, out: torch.Tensor) -> None:
    """out[i, e] = src[row_idx[i], col_idx[e]] (fill where col_idx[e] < 0); see bess_select_scores."""
    require_cuda(src, col_idx, out)
    for r0 in range(0, n_rows, 65535):  # grid.y limit
        nr = min(65535, n_rows - r0)
        if row_idx is not None:
            src_ptr, ridx = src.data_ptr(), row_idx.data_ptr() + 4 * r0
        else:
            src_ptr, ridx = src.data_ptr() + 4 * r0 * src.stride(0), None
        call("bess_select_scores", src_ptr, src.stride(0), ridx, nr, col_idx.data_ptr(),
             col_idx.numel(), float(fill), out.data_ptr() + 4 * r0 * out.stride(0), out.stride(0),
             _st(src))


def pairs_get(mat: torch.Tensor, rows_: Optional[torch.Tensor], cols: torch.Tensor,
              out: torch.Tensor) -> None:
    require_cuda(mat, cols, out)
    call("bess_pairs_get", mat.data_ptr(), mat.stride(0), ptr(rows_), cols.data_ptr(), cols.numel(),
         out.data_ptr(), _st(mat))


def pairs_set(mat: torch.Tensor, rows_: Optional[torch.Tensor], cols: torch.Tensor,
              values: Optional[torch.Tensor], value: float = 0.0) -> None:
    require_cuda(mat, cols)
    call("bess_pairs_set", mat.data_ptr(), mat.stride(0), ptr(rows_), cols.data_ptr(), cols.numel(),
         ptr(values), float(value), _st(mat))


# ------------------------------------------------------ peer exchange -------
def _ptr_array(ptrs: Sequence[int]):
    return (C.c_void_p * max(len(ptrs), 1))(*[C.c_void_p(p) for p in ptrs])


def peer_signal(counter: torch.Tensor, peer_flag_ptrs: Sequence[int], my_rank: int) -> None:
    call("bess_peer_signal", counter.data_ptr(), _ptr_array(peer_flag_ptrs), my_rank,
         len(peer_flag_ptrs), _st(counter))


def peer_timeout_ms() -> int:
    """Wall-clock bound of a peer wait: BESS_PEER_TIMEOUT_S seconds (default 600; 0 = none).
    Long enough for a checkpoint save or a slow batch on one rank, short enough that a dead
    peer ends in a CUDA error rather than a hung GPU."""
    import os
    return int(float(os.environ.get("BESS_PEER_TIMEOUT_S", "600")) * 1000)


def peer_wait(counter: torch.Tensor, my_flags: torch.Tensor, n: int,
              timeout_ms: Optional[int] = None) -> None:
    call("bess_peer_wait", counter.data_ptr(), my_flags.data_ptr(), n,
         peer_timeout_ms() if timeout_ms is None else int(timeout_ms), _st(counter))


def peer_push(src: torch.Tensor, src_stride_bytes: int, dst_ptrs: Sequence[int],
              bytes_each: int) -> None:
    call("bess_peer_push", src.data_ptr(), src_stride_bytes, _ptr_array(dst_ptrs), len(dst_ptrs),
         bytes_each, _st(src))


def peer_copy(src_ptr: int, dst_ptr: int, nbytes: int, stream: "torch.cuda.Stream") -> None:
    """copy-engine transfer of one contiguous block (local or peer-mapped destination)."""
    call("bess_peer_copy", src_ptr, dst_ptr, nbytes, stream.cuda_stream)


def peer_reduce(slots: torch.Tensor, n: int, count: int, scale: float, out: torch.Tensor) -> None:
    call("bess_peer_reduce", slots.data_ptr(), n, count, float(scale), out.data_ptr(), _st(out))


class PeerBuffer:
    """A zero-filled device buffer from bess_peer_alloc, exportable to other processes
    through a CUDA IPC handle; `tensor` is a zero-copy uint8 view of it."""

    def __init__(self, nbytes: int, device: torch.device) -> None:
        p = C.c_void_p()
        with torch.cuda.device(device):
            call("bess_peer_alloc", int(nbytes), C.byref(p))
        self.ptr, self.nbytes, self.device = int(p.value), int(nbytes), device
        self.__cuda_array_interface__ = dict(shape=(self.nbytes,), typestr="|u1",
                                             data=(self.ptr, False), version=3, strides=None)
        self.tensor = torch.as_tensor(self, device=device)
        assert self.tensor.data_ptr() == self.ptr

    def export(self) -> bytes:
        h = (C.c_ubyte * 64)()
        call("bess_peer_export", self.ptr, h)
        return bytes(h)


def peer_import(handle: bytes, device: torch.device) -> int:
    """Map another process's PeerBuffer (by its exported handle); returns the device pointer."""
    buf = (C.c_ubyte * 64).from_buffer_copy(handle)
    p = C.c_void_p()
    with torch.cuda.device(device):
        call("bess_peer_import", buf, C.byref(p))
    return int(p.value)


def stamp(slot: torch.Tensor) -> None:
    """slot (int64, 1 element) <- %globaltimer when the current stream gets here."""
    call("bess_stamp", slot.data_ptr(), _st(slot))


def fill_f32(t: torch.Tensor, v: float) -> None:
    call("bess_fill_f32", t.data_ptr(), t.numel(), float(v), _st(t))


def fill_i32(t: torch.Tensor, v: int) -> None:
    call("bess_fill_i32", t.data_ptr(), t.numel(), int(v), _st(t))


def cast_from_f32(src: torch.Tensor, dst: torch.Tensor) -> None:
    call("bess_cast_from_f32", src.data_ptr(), dst.data_ptr(), dtype_code(dst.dtype), src.numel(),
         _st(src))
