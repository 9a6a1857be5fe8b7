"""EmbeddingMovingBessKGE._plan on the CPU: the passes' row maps must address exactly the rows
the reference's reshape / transpose / split chain selects from the exchanged buffer
(/root/reference/besskge/bess.py:341-466), in the same column order.  No device work."""
import types

import numpy as np
import pytest
import torch

from besskge_b200 import _lib as L
from besskge_b200.bess import EmbeddingMovingBessKGE


def map_row(m, x):
    """csrc/common.cuh map_row."""
    if m.group <= 0:
        return x + m.offset
    r = m.offset
    if m.group1 > 0:
        a = x // m.group1
        r += a * m.stride1
        x -= a * m.group1
    g = x // m.group
    return r + g * m.stride + (x - g * m.group)


def plan(scheme, flat, shared, n, p, B, Nn):
    fake = types.SimpleNamespace(
        negative_sampler=types.SimpleNamespace(corruption_scheme=scheme, flat_negative_format=flat,
                                               local_sampling=False),
        score_fn=types.SimpleNamespace(negative_sample_sharing=shared),
        augment_negative=False)
    return EmbeddingMovingBessKGE._plan(fake, n, p, B, Nn)


def reference_columns(scheme, flat, shared, n, p, B, Nn):
    """For every query position s in [0, S): the storage rows (in the [n, per] receive buffer of
    one replica) of its negative columns, derived with the reference's own tensor ops."""
    per = p + B * Nn
    S = n * p
    # received buffer: block j = [p tails | B*Nn negatives] from shard j; value = storage row
    buf = torch.arange(n * per).view(n, per)
    neg = buf[:, p:]                                          # bess.py:351-355 (split)
    neg = neg.reshape(n, B, Nn).transpose(0, 1).flatten(1, 2)  # [B, n*Nn]   (bess.py:356-360)

    def scored(cand, n_query):
        """column rows per query of score_heads / score_tails (scoring.py:176-200, 231-255)."""
        if shared or cand.shape[0] == 1:
            flat_list = cand.reshape(-1)
            return flat_list.unsqueeze(0).expand(n_query, -1)
        return cand

    if scheme in ("h", "t"):
        return scored(neg, S)
    cut = p // 2
    if flat:
        nh, nt = neg[0:1], neg[1:2]                           # bess.py:419-422
    else:
        ne = neg.reshape(n, p, -1)                            # bess.py:423-429
        nh, nt = ne[:, :cut].flatten(end_dim=1), ne[:, cut:].flatten(end_dim=1)
    s1 = scored(nh, n * cut).reshape(n, cut, -1)
    s2 = scored(nt, n * (p - cut)).reshape(n, p - cut, -1)
    return torch.cat([s1, s2], dim=1).flatten(end_dim=1)      # bess.py:457-465


@pytest.mark.parametrize("scheme", ["h", "t", "ht"])
@pytest.mark.parametrize("flat", [True, False])
@pytest.mark.parametrize("shared", [True, False])
@pytest.mark.parametrize("n", [1, 2, 4])
def test_plan_addresses_the_reference_columns(scheme, flat, shared, n):
    if flat and not shared:
        pytest.skip("flat negatives are one list for every query: scored shared or not alike")
    p, Nn = 6, 3
    S = n * p
    B = ((2 if scheme == "ht" else 1) if flat else S)
    passes, n_col = plan(scheme, flat, shared, n, p, B, Nn)
    want = reference_columns(scheme, flat, shared, n, p, B, Nn)
    assert want.shape == (S, n_col)
    got = np.full((S, n_col), -1, dtype=np.int64)
    for ps in passes:
        assert not ps.cand_from_head and not ps.aug
        for q in range(ps.n_query):
            s = map_row(ps.qmap, q)  # position of the query in the micro-batch
            for c in range(ps.n_cand):
                # shared: candidate c of the pass; per-query (csrc/pair.cu pertriple kernels):
                # cand.map(c) + position of the query in the micro-batch * q_stride
                row = map_row(ps.cand_map, c) + (0 if ps.shared else s * ps.q_stride)
                got[s, ps.col0 + c] = row
    assert (got >= 0).all(), "every column of every query is written by exactly one pass"
    np.testing.assert_array_equal(got, want.numpy())


def test_plan_query_side_rows():
    """the fixed (query-side) rows of an 'ht' plan: tails of the first half of every partition
    for the heads pass, heads of the second half for the tails pass (bess.py:404-418)."""
    n, p, Nn = 2, 6, 3
    per = p + 2 * Nn
    passes, _ = plan("ht", True, True, n, p, 2, Nn)
    heads_pass, tails_pass = passes
    assert heads_pass.mode == L.MODE_HEADS and not heads_pass.fixed_from_head
    assert tails_pass.mode == L.MODE_TAILS and tails_pass.fixed_from_head
    half = p // 2
    for q in range(n * half):
        i, r = divmod(q, half)
        assert map_row(heads_pass.fixed_map, q) == i * per + r          # tail row in TN[i]
        assert map_row(tails_pass.fixed_map, q) == i * p + half + r      # head row in H
        assert map_row(heads_pass.qmap, q) == i * p + r
        assert map_row(tails_pass.qmap, q) == i * p + half + r
