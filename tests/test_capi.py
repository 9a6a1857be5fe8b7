"""The C-ABI shared library loads (no GPU needed) and exports every symbol that
include/besskge_b200.h declares; the ctypes binding covers them all; the product
fails loudly (no CPU fallback) when handed CPU tensors."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

import besskge_b200
from besskge_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "besskge_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bess_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_all_declared_symbols():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.bess_version() >= 100


def test_binding_covers_header():
    bound = set(_lib.SIGNATURES) | {"bess_last_error"}
    assert set(declared_symbols()) == bound


def test_struct_layout_matches_header():
    # bess_rowmap_t: 5 x int32; bess_rows_t: 2 pointers + rowmap + int64 pitch
    assert ctypes.sizeof(_lib.RowMap) == 20
    assert ctypes.sizeof(_lib.Rows) == 48
    assert ctypes.sizeof(_lib.ScoreCfg) == 32  # 6 x int32 + eps + rel_u


def test_widths_from_library():
    lib = _lib.load()
    cfg = _lib.ScoreCfg(family=_lib.BOXE, norm_p=2, d=8, normalize=0, apply_tanh=1, per_dim=1,
                        eps=1e-6)
    assert lib.bess_entity_width(ctypes.byref(cfg)) == 16
    assert lib.bess_relation_width(ctypes.byref(cfg)) == 34
    assert lib.bess_query_nvec(ctypes.byref(cfg)) == 3
    tri = _lib.ScoreCfg(family=_lib.TRIPLERE, norm_p=1, d=8, normalize=1, rel_u=0.5)
    assert lib.bess_entity_width(ctypes.byref(tri)) == 8
    assert lib.bess_relation_width(ctypes.byref(tri)) == 24
    assert lib.bess_query_nvec(ctypes.byref(tri)) == 2


def test_no_cpu_fallback():
    from besskge_b200.scoring import TransE
    from besskge_b200.sharding import Sharding

    sh = Sharding.create(40, 2, seed=1)
    sf = TransE(True, 1, sh, 3, 8)
    h = torch.randn(4, 8)
    with pytest.raises(besskge_b200.BessLibraryError):
        sf.score_triple(h, torch.zeros(4, dtype=torch.int32), h)
    from besskge_b200.bess import EmbeddingMovingBessKGE
    from besskge_b200.negative_sampler import RandomShardedNegativeSampler

    ns = RandomShardedNegativeSampler(4, sh, 0, "t", False, True)
    model = EmbeddingMovingBessKGE(ns, sf, return_scores=True)
    z = torch.zeros(2, 2, 2, dtype=torch.int32)
    with pytest.raises(besskge_b200.BessLibraryError):
        model(z, z, z, torch.zeros(2, 2, 1, 4, dtype=torch.int32))


def test_constructor_validation_matches_reference():
    """bess.py:79-102."""
    from besskge_b200.bess import EmbeddingMovingBessKGE, ScoreMovingBessKGE
    from besskge_b200.loss import LogSigmoidLoss
    from besskge_b200.negative_sampler import RandomShardedNegativeSampler
    from besskge_b200.scoring import TransE
    from besskge_b200.sharding import Sharding

    sh = Sharding.create(40, 2, seed=1)
    flat_ns = RandomShardedNegativeSampler(4, sh, 0, "t", False, True)
    with pytest.raises(ValueError):
        EmbeddingMovingBessKGE(flat_ns, TransE(True, 1, sh, 3, 8))
    with pytest.raises(AssertionError):
        EmbeddingMovingBessKGE(flat_ns, TransE(False, 1, sh, 3, 8), return_scores=True)
    with pytest.raises(AssertionError):
        ScoreMovingBessKGE(flat_ns, TransE(True, 1, sh, 3, 8), LogSigmoidLoss(1.0, False),
                           augment_negative=True)
