"""Shared loader for the AllScoresPipeline goldens (tests/golden/pipeline_*.npz)."""
import numpy as np
import torch

from .conftest import load_golden

EW = {"TransE": 1, "ComplEx": 2}
RW = {"TransE": lambda d: d, "ComplEx": lambda d: 2 * d}


def load_case(name):
    """cfg, arrays and the tables (regenerated from the seed; pinned by the stored probes)."""
    cfg, g = load_golden(name)
    gen = torch.Generator().manual_seed(cfg["seed"])
    d = cfg["d"]
    ent = torch.randn(cfg["n_entity"], EW[cfg["family"]] * d, generator=gen)
    rel = torch.randn(cfg["n_rel"], RW[cfg["family"]](d), generator=gen)
    assert torch.equal(ent[::97, ::5], torch.from_numpy(g["ent_probe"])), "torch.randn stream changed"
    assert torch.equal(rel[:, ::5], torch.from_numpy(g["rel_probe"]))
    return cfg, g, ent, rel


def filter_list(cfg, triples):
    if not cfg["filter"]:
        return None
    gt_col = 0 if cfg["scheme"] == "h" else 2
    extra = np.copy(triples)
    extra[:, gt_col] += 1
    return [extra] if cfg["extra_only"] else [triples, extra]
