"""CPU: besskge_b200.utils.get_entity_filter / EntityFilterIndex (sort-merge join) returns the
same rows IN THE SAME ORDER as the reference's dense comparison + nonzero (utils.py:36-69)."""
import numpy as np
import pytest
import torch

from besskge_b200.utils import EntityFilterIndex, get_entity_filter


def dense_reference(triples, filter_triples, mode):
    """utils.py:56-69 restated: dense [x, y] match matrix, nonzero in row-major order."""
    ent_col = 0 if mode == "t" else 2
    rel_f = filter_triples[:, 1] == triples[:, 1].view(-1, 1)
    ent_f = filter_triples[:, ent_col] == triples[:, ent_col].view(-1, 1)
    f = (ent_f & rel_f).nonzero(as_tuple=False)
    f[:, 1] = filter_triples[:, 2 - ent_col].view(1, -1)[:, f[:, 1]]
    return f


@pytest.mark.parametrize("mode", ["h", "t"])
@pytest.mark.parametrize("x,y,E,R", [(50, 400, 20, 3), (1000, 3000, 300, 7), (5, 5, 2, 1),
                                     (64, 2000, 5, 2), (300, 10, 1000, 50)])
def test_entity_filter_matches_dense_reference(mode, x, y, E, R):
    rng = np.random.default_rng(x * 31 + y)
    mk = lambda n: torch.from_numpy(np.stack(
        [rng.integers(E, size=n), rng.integers(R, size=n), rng.integers(E, size=n)], axis=1))
    tr, ft = mk(x), mk(y)
    want = dense_reference(tr, ft, mode)
    assert torch.equal(get_entity_filter(tr, ft, mode), want)
    # the prebuilt index answers batches without re-sorting, numpy or torch inputs alike
    index = EntityFilterIndex(ft.numpy(), mode)
    parts = [index.query(tr[i:i + 17]) for i in range(0, x, 17)]
    got = torch.cat([torch.stack([p[:, 0] + i * 17, p[:, 1]], dim=1) for i, p in enumerate(parts)])
    assert torch.equal(got, want)


def test_entity_filter_edge_cases():
    empty = torch.zeros((0, 3), dtype=torch.int64)
    some = torch.tensor([[1, 0, 2], [3, 1, 4]])
    assert get_entity_filter(empty, some, "t").shape == (0, 2)
    assert get_entity_filter(some, empty, "h").shape == (0, 2)
    # a relation id beyond the filter set's range matches nothing (no key aliasing)
    q = torch.tensor([[1, 7, 2]])
    assert EntityFilterIndex(some, "t").query(q).shape == (0, 2)
    with pytest.raises(ValueError):
        get_entity_filter(some, some, "ht")
