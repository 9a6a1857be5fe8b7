"""ctypes binding of `include/besskge_b200.h` (the C-ABI of the CUDA library).

There is NO CPU fallback: if the library cannot be loaded every device entry
point raises.  Host-only modules (sharding, samplers) do not import this file.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

import torch

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libbesskge_b200.so"

F32, F16, BF16 = 0, 1, 2
F16X3 = 3  # operand format of the tensor-core path: scaled fp16 (hi, lo) pairs of fp32 values
TRANSE, ROTATE, DISTMULT, COMPLEX, PAIRRE, BOXE, TRIPLERE, INTERHT, TRANS = range(9)
MODE_TAILS, MODE_HEADS = 0, 1
LOSS_LOGSIGMOID, LOSS_MARGIN_RANKING, LOSS_SOFTMAX_CE = 0, 1, 2
OPT_SGD, OPT_SGDM, OPT_ADAMW = 0, 1, 2
MAX_SHARD = 16
(HYPER_LR, HYPER_MOMENTUM, HYPER_DAMPENING, HYPER_BETA1, HYPER_BETA2, HYPER_EPS, HYPER_WEIGHT_DECAY,
 HYPER_BC1, HYPER_BC2, HYPER_FIRST_STEP) = range(10)
HYPER_COUNT = 16


class RowMap(C.Structure):
    _fields_ = [("group", C.c_int32), ("stride", C.c_int32), ("offset", C.c_int32),
                ("group1", C.c_int32), ("stride1", C.c_int32)]


class Rows(C.Structure):
    _fields_ = [
        ("base", C.c_void_p),
        ("idx", C.c_void_p),
        ("map", RowMap),
        ("pitch", C.c_int64),
    ]


class ScoreCfg(C.Structure):
    _fields_ = [
        ("family", C.c_int32),
        ("norm_p", C.c_int32),
        ("d", C.c_int32),
        ("normalize", C.c_int32),
        ("apply_tanh", C.c_int32),
        ("per_dim", C.c_int32),
        ("eps", C.c_float),
        ("rel_u", C.c_float),
    ]


IDENT = RowMap(0, 0, 0, 0, 0)

_P = C.c_void_p
_I = C.c_int
_L = C.c_int64
_F = C.c_float
_CFG = C.POINTER(ScoreCfg)

# name -> argtypes; every function returns int (0 = ok) unless listed in _RESTYPE
SIGNATURES = {
    "bess_version": [],
    "bess_launch_count": [],
    "bess_entity_width": [_CFG],
    "bess_relation_width": [_CFG],
    "bess_query_nvec": [_CFG],
    "bess_gather_route": [_P, _L, _I, _I, _P, _I, _I, _I, _P, C.POINTER(_P), _I, _P],
    "bess_gather_rows": [_P, _L, _I, _I, _P, _I, _P, _P],
    "bess_score_triple_fwd": [_CFG, _I, Rows, Rows, _P, _P, RowMap, _I, _P, RowMap, _P],
    "bess_score_triple_bwd": [_CFG, _I, Rows, Rows, _P, _P, RowMap, _I, _P, _P, RowMap, Rows, Rows,
                              _P, _I, _I, _I, _P],
    "bess_query_prologue_fwd": [_CFG, _I, _I, Rows, _P, _P, RowMap, _I, _P, _P],
    "bess_query_prologue_bwd": [_CFG, _I, _I, Rows, _P, _P, RowMap, _I, _P, Rows, _P, _I, _I, _P],
    "bess_boxe_rel_finalize": [_CFG, _I, _P, _P, _I, _P, _P],
    "bess_cand_inv_norm": [_I, Rows, _I, _I, _P, _P],
    "bess_cand_norm_bwd": [_I, Rows, _I, _I, _P, Rows, _P],
    "bess_score_shared_fwd": [_CFG, _I, _I, _P, _I, Rows, _P, _I, _P, RowMap, _L, _I, _P, _P],
    "bess_score_shared_bwd_query": [_CFG, _I, _I, _P, _I, Rows, _P, _I, _P, _P, RowMap, _L, _I, _P,
                                    _P, _P],
    "bess_shared_bwd_cand_workspace": [_CFG, _I, _I],
    "bess_score_shared_bwd_cand": [_CFG, _I, _I, _P, _I, Rows, _P, _I, _P, _P, RowMap, _L, _I, _P,
                                   Rows, _I, _P, _P],
    "bess_dot_gemm_workspace": [_I, _I, _I],
    "bess_dot_gemm": [_I, _P, _P, _L, _I, _P, _P, _L, _I, _I, _I, _P, RowMap, _L, _I, _I, _P, _L, _P, _P,
                      _P],
    "bess_split_operand": [_I, Rows, _I, _I, _P, _I, _P, _P, _L, _P, _P, _L, _P, _P],
    "bess_operand_scale": [_I, Rows, _I, _I, _P, _F, _P, _P, _P],
    "bess_row_sqnorm": [_I, Rows, _I, _I, _P, _P],
    "bess_l2_from_dot": [_P, RowMap, _L, _I, _I, _I, _P, _P, _P],
    "bess_l2_coef_workspace": [_I, _I],
    "bess_l2_coef": [_P, _P, RowMap, _L, _I, _I, _I, _P, _L, _P, _P, _P, _P],
    "bess_rows_axpy": [_I, _P, _F, Rows, Rows, _I, _I, _P],
    "bess_table_operand_refresh": [_P, _L, _I, _L, _P, _P, _L, _P, _I, _P],
    "bess_score_pertriple_fwd": [_CFG, _I, _I, _P, _I, Rows, _L, _I, _P, RowMap, _L, _I, _P, _P],
    "bess_score_pertriple_bwd": [_CFG, _I, _I, _P, _I, Rows, _L, _I, _P, _P, RowMap, _L, _I, _P, _P,
                                 Rows, _P],
    "bess_mask_add": [_P, _I, _I, _L, _P, _L, _I, _I, _F, _P],
    "bess_mask_diag": [_P, _I, _L, _I, _I, _I, _F, _P],
    "bess_loss_fwd_bwd": [_I, _F, _I, _F, _F, _L, _P, _P, _I, _I, _L, _P, _I, _P, _P, _P, _P],
    "bess_loss_fwd_bwd_operand": [_I, _F, _I, _F, _F, _L, _P, _P, _I, _I, _L, _P, _I, _P, _P, _I, _P, _P,
                                  _L, _P, _P],
    "bess_sum_f32": [_P, _I, _P, _P],
    "bess_rank_from_scores": [_P, _P, _I, _I, _L, _I, _I, _P, _P],
    "bess_sort_workspace": [_I],
    "bess_sort_keys": [_P, _I, _I, _P, _P, _P, _P],
    "bess_scatter_sgd": [_P, _L, _I, _I, _P, _P, _I, _I, _I, _P, _P, _L, _F, _P, _P],
    "bess_scatter_collect": [_I, _P, _P, _I, _I, _I, _P, _P, _L, _P, _P, _P],
    "bess_opt_dense": [_I, _P, _L, _I, _I, _I, _P, _P, _P, _P, _F, _F, _F, _F, _F, _F, _F, _I, _P, _F, _I,
                       _P],
    "bess_scatter_accumulate": [_I, _P, _P, _I, _I, _I, _P, _P, _L, _P, _P],
    "bess_set_hyper": [_P, _F, _F, _F, _F, _F, _F, _F, _I, _P],
    "bess_relation_grad_reduce": [_P, _I, _P, _P, _I, _I, _P, _P],
    "bess_topk_merge": [_P, _L, _I, _I, _P, _L, _I, _P, _P, _I, _P],
    "bess_topk_exact_supported": [_I],
    "bess_topk_exact_rescore": [_CFG, _I, _I, _P, _L, _P, _L, _P, _P, _L, _I, _I, _P, _P, _I, _I, _I, _P,
                                _P, _P],
    "bess_topk_finalize": [_P, _P, _I, _I, _I, _P, _P, _I, _I, _F, _P, _P, _P],
    "bess_select_scores": [_P, _L, _P, _I, _P, _I, _F, _P, _L, _P],
    "bess_pairs_get": [_P, _L, _P, _P, _I, _P, _P],
    "bess_pairs_set": [_P, _L, _P, _P, _I, _P, _F, _P],
    "bess_peer_signal": [_P, C.POINTER(_P), _I, _I, _P],
    "bess_peer_wait": [_P, _P, _I, _L, _P],
    "bess_peer_push": [_P, _L, C.POINTER(_P), _I, _L, _P],
    "bess_peer_copy": [_P, _P, _L, _P],
    "bess_peer_reduce": [_P, _I, _L, _F, _P, _P],
    "bess_peer_alloc": [_L, C.POINTER(_P)],
    "bess_peer_free": [_P],
    "bess_peer_export": [_P, _P],
    "bess_peer_import": [_P, C.POINTER(_P)],
    "bess_peer_unmap": [_P],
    "bess_take_along_rows": [_P, _I, _L, _I, _P, _I, _I, _P, _P],
    "bess_complex_mul": [_I, _P, _P, _I, _I, _I, _P, _P],
    "bess_stamp": [_P, _P],
    "bess_fill_f32": [_P, _L, _F, _P],
    "bess_fill_i32": [_P, _L, C.c_int32, _P],
    "bess_cast_from_f32": [_P, _P, _I, _L, _P],
}
_RESTYPE = {
    "bess_last_error": C.c_char_p,
    "bess_shared_bwd_cand_workspace": C.c_int64,
    "bess_sort_workspace": C.c_int64,
    "bess_dot_gemm_workspace": C.c_int64,
    "bess_l2_coef_workspace": C.c_int64,
    "bess_launch_count": C.c_int64,
}
_NO_STATUS = {"bess_version", "bess_topk_exact_supported", "bess_launch_count", "bess_entity_width", "bess_relation_width", "bess_query_nvec",
              "bess_shared_bwd_cand_workspace", "bess_sort_workspace", "bess_dot_gemm_workspace",
              "bess_l2_coef_workspace"}

_lib: Optional[C.CDLL] = None


class BessLibraryError(RuntimeError):
    pass


def library_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """dlopen the CUDA library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise BessLibraryError(
            f"{_LIB_PATH} not found: build it with `python -m besskge_b200._build` "
            "(besskge_b200 has no CPU fallback)"
        )
    lib = C.CDLL(str(_LIB_PATH))
    lib.bess_last_error.restype = C.c_char_p
    lib.bess_last_error.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, C.c_int)
    _lib = lib
    return lib


def call(name: str, *args):
    """Invoke a status-returning entry point; raise on a non-zero status."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if name in _NO_STATUS:
        return rc
    if rc != 0:
        msg = lib.bess_last_error()
        raise BessLibraryError(f"{name} failed ({rc}): {msg.decode() if msg else ''}")
    return rc


# ---------------------------------------------------------------- helpers ---
def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.float16:
        return F16
    if dt == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported table dtype {dt}")


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise BessLibraryError(
                "besskge_b200 runs on CUDA tensors only (no CPU fallback); got a "
                f"{t.device} tensor"
            )


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def rowmap(group: int = 0, stride: int = 0, offset: int = 0, group1: int = 0,
           stride1: int = 0) -> RowMap:
    return RowMap(int(group), int(stride), int(offset), int(group1), int(stride1))


def rows(t: torch.Tensor, idx: Optional[torch.Tensor] = None, rmap: Optional[RowMap] = None,
         pitch: Optional[int] = None, offset_elems: int = 0) -> Rows:
    """Row set over the last dim of `t` (any leading shape, contiguous rows)."""
    if pitch is None:
        pitch = t.stride(-2) if t.dim() >= 2 else t.shape[-1]
    base = t.data_ptr() + offset_elems * t.element_size()
    return Rows(base, None if idx is None else idx.data_ptr(), rmap if rmap is not None else IDENT,
                int(pitch))
