"""Shared helpers for the -m gpu parity tests (product CUDA path vs oracle / goldens)."""
import numpy as np
import torch

import besskge_b200 as B
from besskge_b200 import bess, loss, negative_sampler, scoring
from besskge_b200.sharding import Sharding

DEV = "cuda"


def T(a):
    return torch.from_numpy(np.asarray(a))


def score_cfg(fam, d, p=2, **kw):
    return dict(family=fam, d=d, norm_p=p or 2, normalize=kw.get("normalize", True),
                apply_tanh=kw.get("apply_tanh", True), per_dim=kw.get("per_dim", True), eps=1e-6,
                rel_u=kw.get("u", kw.get("offset", 1.0 if fam in ("InterHT", "TranS") else 0.0)))


def make_score_fn(fam, sharing, p, sh, n_rel, d, ent, rel, dtype=torch.float32, **kw):
    cls = getattr(scoring, fam)
    if fam in ("DistMult", "ComplEx"):
        sf = cls(sharing, sh, n_rel, d, entity_initializer=ent, relation_initializer=rel, **kw)
    else:
        sf = cls(sharing, p, sh, n_rel, d, entity_initializer=ent, relation_initializer=rel, **kw)
    return sf.to(device=DEV, dtype=dtype)


def fake_sampler(scheme, flat, triple_based=True, local=False):
    """Sampler object carrying only the flags the device modules read."""
    cls = (negative_sampler.TripleBasedShardedNegativeSampler if triple_based
           else negative_sampler.RandomShardedNegativeSampler)
    ns = cls.__new__(cls)
    ns.corruption_scheme = scheme
    ns.flat_negative_format = flat
    ns.local_sampling = local
    ns.mask_on_gather = False
    return ns


def make_loss(cfg):
    kind = cfg["kind"]
    if kind == "logsigmoid":
        return loss.LogSigmoidLoss(cfg["margin"], cfg.get("negative_adversarial_sampling", False),
                                   cfg.get("negative_adversarial_scale", 1.0),
                                   cfg.get("loss_scale", 1.0))
    if kind == "margin_ranking":
        return loss.MarginRankingLoss(cfg["margin"], cfg.get("negative_adversarial_sampling", False),
                                      cfg.get("negative_adversarial_scale", 1.0),
                                      cfg.get("loss_scale", 1.0))
    return loss.SampledSoftmaxCrossEntropyLoss(cfg["n_entity"], cfg.get("loss_scale", 1.0))


def oracle_loss_cfg(case):
    return dict(kind=case["kind"], margin=case.get("margin", 0.0),
                adversarial=case.get("negative_adversarial_sampling", False),
                adv_scale=case.get("negative_adversarial_scale", 1.0),
                loss_scale=case.get("loss_scale", 1.0), n_entity=case.get("n_entity", 2))
