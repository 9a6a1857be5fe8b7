"""TEST INFRASTRUCTURE — CPU oracle for the BESS-KGE hot path.

A plain numpy / torch-CPU restatement of the reference algorithm
(graphcore-research/bess-kge, files under /root/reference/besskge/), written
functionally and without any collective: the multi-shard step is evaluated
with the explicit routing rule

    replica r scores blocks (r, j), j = 0..n-1, with
      heads      shard[r][head[r, j, :]]
      tails      shard[j][tail[j, r, :]]
      relations  relation[r, j, :]
      negatives  shard[j][negative[j, r, b, :]]   concatenated over j.

Pinned: `tests/test_oracle_golden.py` checks every function here against
`tests/golden/*.npz`, which were produced by importing the UNMODIFIED
reference in the build container (`tests/golden/make_golden.py`).
Only tests/, `__graft_entry__.smoke()` and bench.py's CPU-baseline leg may
import this module; the product never does.

Third-party arithmetic restated (not under /root/reference):
  pea.distance_matrix(a, b, p)  (poptorch-experimental-addons @ 899aec4a,
  call site scoring.py:195-197)  ==  pairwise p-distance matrix;
  backward + optimizer == torch.autograd + dense torch.optim (SURVEY.md §8c:
  "parity unpinned" by any reference test; defined by this oracle).
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

BAD_NEGATIVE_SCORE = -50000.0


# =============================================================================
# host side: sharding / partitioning / sampling  (numpy, bit-exact)
# =============================================================================
def sharding_create(n_entity: int, n_shard: int, seed: int, type_offsets=None) -> Dict[str, Any]:
    """sharding.py:67-137."""
    rng = np.random.default_rng(seed)
    rows = int(np.ceil(n_entity / n_shard))
    table = np.sort(rng.permutation(n_shard * rows).reshape(n_shard, rows), axis=1)
    e2s, e2i = np.divmod(np.argsort(table.flatten())[:n_entity], rows)
    deduct = np.sum(table[:, -n_shard:] >= n_entity, axis=-1)
    out = dict(n_shard=n_shard, entity_to_shard=e2s, entity_to_idx=e2i,
               shard_and_idx_to_entity=table, shard_counts=rows - deduct,
               entity_type_counts=None, entity_type_offsets=None)
    if type_offsets is not None:
        tid = (np.digitize(table, bins=type_offsets)
               + len(type_offsets) * np.arange(n_shard)[:, None] - 1)
        counts = np.bincount(tid.flatten(), minlength=len(type_offsets) * n_shard).reshape(
            n_shard, -1)
        offs = np.c_[[0] * n_shard, np.cumsum(counts, axis=1)[:, :-1]]
        counts[:, -1] -= deduct
        out.update(entity_type_counts=counts, entity_type_offsets=offs)
    return out


def partition_triples(triples: np.ndarray, sh: Dict[str, Any], mode: str):
    """sharding.py:226-265."""
    n = sh["n_shard"]
    if mode == "ht_shardpair":
        a, b = sh["entity_to_shard"][triples[:, [0, 2]].T]
        pid = a * n + b
        counts = np.bincount(pid, minlength=n * n).reshape(n, n)
        offsets = np.concatenate([np.array([0]), np.cumsum(counts)[:-1]]).reshape(n, n)
    else:
        col = 0 if mode == "h_shard" else -1
        pid = sh["entity_to_shard"][triples[:, col]]
        counts = np.bincount(pid, minlength=n)
        offsets = np.concatenate([np.array([0]), np.cumsum(counts)[:-1]])
    order = np.argsort(pid)
    st = triples[order]
    if mode in ("h_shard", "ht_shardpair"):
        st[:, 0] = sh["entity_to_idx"][st[:, 0]]
    if mode in ("t_shard", "ht_shardpair"):
        st[:, -1] = sh["entity_to_idx"][st[:, -1]]
    return st, counts, offsets, order


def random_negatives(rng: np.random.Generator, shard_counts: np.ndarray, bps: int, n: int, B: int,
                     n_negative: int) -> np.ndarray:
    """negative_sampler.py:119-131."""
    return (rng.integers(1 << 31, size=(bps, n, n, B, n_negative)).astype(np.int32)
            % shard_counts[None, :, None, None, None])


def random_sample_idx(rng: np.random.Generator, offsets: np.ndarray, counts: np.ndarray,
                      size: Tuple[int, ...]) -> np.ndarray:
    """batch_sampler.py:391-398."""
    return np.expand_dims(offsets, axis=(0, -1)) + rng.integers(1 << 63, size=size) % np.expand_dims(
        counts, axis=(0, -1))


# =============================================================================
# score functions (torch CPU, any float dtype)
# =============================================================================
def _norm(v: torch.Tensor, p: int) -> torch.Tensor:
    return torch.norm(v, p=p, dim=-1)


def _cmul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """utils.py:72-89."""
    c = a.shape[-1] // 2
    ar, ai, br, bi = a[..., :c], a[..., c:], b[..., :c], b[..., c:]
    return torch.cat([ar * br - ai * bi, ar * bi + ai * br], dim=-1)


def _crot(v: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
    """utils.py:92-112 (non-IPU branch: full-precision cos/sin, no pi factor)."""
    return _cmul(v, torch.cat([torch.cos(r), torch.sin(r)], dim=-1))


def _pdist(q: torch.Tensor, c: torch.Tensor, p: int, shared: bool) -> torch.Tensor:
    """scoring.py:176-200: q [S, W]; c [B, N, W]."""
    if shared:
        flat = c.reshape(-1, q.shape[-1])
        return _norm(q.unsqueeze(1) - flat.unsqueeze(0), p)
    return _norm(q.unsqueeze(1) - c, p)


def _pdot(q: torch.Tensor, c: torch.Tensor, shared: bool) -> torch.Tensor:
    """scoring.py:231-255."""
    if shared:
        return q @ c.reshape(-1, q.shape[-1]).T
    return torch.sum(q.unsqueeze(1) * c, dim=-1)


def _boxe_score(cfg: Dict[str, Any], bumped, center, width, size) -> torch.Tensor:
    """scoring.py:1250-1339."""
    eps, p = cfg["eps"], cfg["norm_p"]
    width = torch.abs(width)
    width = width / torch.clamp(
        torch.exp(torch.mean(torch.log(torch.clamp(width, min=eps)), dim=-1, keepdim=True)),
        min=eps)
    width = width * (1.0 + torch.nn.functional.elu(size.unsqueeze(-1).float()).to(width.dtype))
    if cfg["apply_tanh"]:
        low = torch.tanh(center - 0.5 * width)
        up = torch.tanh(low + width)
        center = 0.5 * (low + up)
        width = up - low
        cd = torch.abs(torch.tanh(bumped) - center)
    else:
        cd = torch.abs(bumped - center)
    wp1 = 1.0 + width
    k = 0.5 * width * (wp1 - torch.reciprocal(wp1))
    inside = torch.le(cd, 0.5 * width)
    if not cfg["per_dim"]:
        inside = torch.all(inside, dim=-1, keepdim=True)
    dist = torch.where(inside, cd / wp1, cd * wp1 - k)
    return -_norm(dist, p).sum(-1)


def _normalize(x: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.normalize(x, p=2, dim=-1)


def score_triple(cfg: Dict[str, Any], h: torch.Tensor, rel_table: torch.Tensor, r: torch.Tensor,
                 t: torch.Tensor) -> torch.Tensor:
    fam, d, p = cfg["family"], cfg["d"], cfg.get("norm_p", 2)
    re = rel_table[r.long()]
    if fam == "TransE":  # scoring.py:321-330
        return -_norm(h + re - t, p)
    if fam == "RotatE":  # scoring.py:423-434
        return -_norm(_crot(h, re) - t, p)
    if fam == "DistMult":  # scoring.py:804-813
        return torch.sum(h * re * t, dim=-1)
    if fam == "ComplEx":  # scoring.py:905-916
        return torch.sum(_cmul(h, re) * t, dim=-1)
    if fam == "PairRE":  # scoring.py:540-553
        rh, rt = re[..., :d], re[..., d:]
        if cfg.get("normalize", True):
            h, t = _normalize(h), _normalize(t)
        return -_norm(h * rh - t * rt, p)
    if fam == "TripleRE":  # scoring.py:683-697
        rh, rm, rt = re[..., :d], re[..., d:2 * d], re[..., 2 * d:]
        u = cfg.get("rel_u", 0.0)
        if u > 0.0:
            rh, rt = rh + u, rt + u
        if cfg.get("normalize", True):
            h, t = _normalize(h), _normalize(t)
        return -_norm(h * rh - t * rt + rm, p)
    if fam in ("InterHT", "TranS"):  # scoring.py:1495-1527, 1665-1697
        hm, ha, tm, ta = h[..., :d], h[..., d:], t[..., :d], t[..., d:]
        if cfg.get("normalize", True):
            hm, ha, tm, ta = _normalize(hm), _normalize(ha), _normalize(tm), _normalize(ta)
        o = cfg.get("rel_u", 1.0)  # the constructors' `offset`
        if fam == "InterHT":
            return -_norm(hm * (ta + o) + re - tm * (ha + o), p)
        rr, rb, rh = re[..., :d], re[..., d:2 * d], re[..., 2 * d:]
        return -_norm(hm * (ta + o + rb) - tm * (ha + o - rh) + rr, p)
    if fam == "BoxE":  # scoring.py:1342-1363
        center, width, size = torch.split(re, 2 * d, dim=-1)
        bumped = h.view(-1, 2, d) + t.view(-1, 2, d)[:, [1, 0]]
        return _boxe_score(cfg, bumped, center.view(-1, 2, d), width.view(-1, 2, d),
                           size.view(-1, 2))
    raise ValueError(fam)


def score_candidates(cfg: Dict[str, Any], mode: str, fixed: torch.Tensor, rel_table: torch.Tensor,
                     r: torch.Tensor, cand: torch.Tensor, shared: bool) -> torch.Tensor:
    """score_tails (mode 't': fixed = heads) / score_heads (mode 'h': fixed =
    tails); cand [B, N, W]."""
    fam, d, p = cfg["family"], cfg["d"], cfg.get("norm_p", 2)
    re = rel_table[r.long()]
    if fam == "TransE":  # scoring.py:333-354
        q = fixed + re if mode == "t" else fixed - re
        return -_pdist(q, cand, p, shared)
    if fam == "RotatE":  # scoring.py:437-462
        q = _crot(fixed, re) if mode == "t" else _crot(fixed, -re)
        return -_pdist(q, cand, p, shared)
    if fam == "DistMult":  # scoring.py:816-837
        return _pdot(fixed * re, cand, shared)
    if fam == "ComplEx":  # scoring.py:919-946
        if mode == "t":
            return _pdot(_cmul(fixed, re), cand, shared)
        conj = torch.cat([re[..., :d], -re[..., d:]], dim=-1)
        return _pdot(_cmul(conj, fixed), cand, shared)
    if fam == "PairRE":  # scoring.py:556-593
        rh, rt = re[..., :d], re[..., d:]
        if cfg.get("normalize", True):
            fixed, cand = _normalize(fixed), _normalize(cand)
        if shared:
            cand = cand.reshape(1, -1, d)
        if mode == "t":
            return -_norm(cand * rt.unsqueeze(1) - (fixed * rh).unsqueeze(1), p)
        return -_norm(cand * rh.unsqueeze(1) - (fixed * rt).unsqueeze(1), p)
    if fam == "TripleRE":  # scoring.py:699-743
        rh, rm, rt = re[..., :d], re[..., d:2 * d], re[..., 2 * d:]
        u = cfg.get("rel_u", 0.0)
        if u > 0.0:
            rh, rt = rh + u, rt + u
        if cfg.get("normalize", True):
            fixed, cand = _normalize(fixed), _normalize(cand)
        if shared:
            cand = cand.reshape(1, -1, d)
        if mode == "t":
            return -_norm(cand * rt.unsqueeze(1) - (fixed * rh + rm).unsqueeze(1), p)
        return -_norm(cand * rh.unsqueeze(1) - (fixed * rt - rm).unsqueeze(1), p)
    if fam in ("InterHT", "TranS"):  # scoring.py:1530-1572, 1700-1750
        fm, fa, cm, ca = fixed[..., :d], fixed[..., d:], cand[..., :d], cand[..., d:]
        if cfg.get("normalize", True):
            fm, fa, cm, ca = _normalize(fm), _normalize(fa), _normalize(cm), _normalize(ca)
        if shared:
            cm, ca = cm.reshape(1, -1, d), ca.reshape(1, -1, d)
        o = cfg.get("rel_u", 1.0)
        if fam == "InterHT":
            rr, rb, rh = re, 0.0, 0.0
        else:
            rr, rb, rh = re[..., :d], re[..., d:2 * d].unsqueeze(1), re[..., 2 * d:].unsqueeze(1)
        fm, fa, rr = fm.unsqueeze(1), fa.unsqueeze(1), rr.unsqueeze(1)
        if mode == "t":  # fixed = head, candidates = tails
            return -_norm(fm * (ca + o + rb) - cm * (fa + o - rh) + rr, p)
        return -_norm(cm * (fa + o + rb) - fm * (ca + o - rh) + rr, p)
    if fam == "BoxE":  # scoring.py:1366-1415
        center, width, size = torch.split(re, 2 * d, dim=-1)
        if shared:
            cand = cand.reshape(1, -1, 2 * d)
        if mode == "t":
            bumped = fixed.view(-1, 1, 2, d) + cand.view(cand.shape[0], -1, 2, d)[:, :, [1, 0]]
        else:
            bumped = cand.view(cand.shape[0], -1, 2, d) + fixed.view(-1, 1, 2, d)[:, :, [1, 0]]
        return _boxe_score(cfg, bumped, center.view(-1, 1, 2, d), width.view(-1, 1, 2, d),
                           size.view(-1, 1, 2))
    raise ValueError(fam)


# =============================================================================
# losses (loss.py) — fp32, sum over the micro-batch
# =============================================================================
def _neg_weights(neg: torch.Tensor, adversarial: bool, scale: float) -> torch.Tensor:
    """loss.py:28-51."""
    if adversarial:
        return torch.softmax(scale * neg, dim=-1).detach()
    return torch.tensor(1.0 / neg.shape[-1])


def loss_value(cfg: Dict[str, Any], pos: torch.Tensor, neg: torch.Tensor,
               w: torch.Tensor) -> torch.Tensor:
    kind = cfg["kind"]
    ls = cfg.get("loss_scale", 1.0)
    if kind == "logsigmoid":  # loss.py:115-134
        nw = _neg_weights(neg, cfg["adversarial"], cfg.get("adv_scale", 1.0))
        m = cfg["margin"]
        pl = torch.nn.functional.logsigmoid(pos + m)
        nl = torch.nn.functional.logsigmoid(-neg - m)
        return ls * (-0.5) * torch.sum(w * (pl + torch.sum(nw * nl, dim=-1)))
    if kind == "margin_ranking":  # loss.py:179-195
        nw = _neg_weights(neg, cfg["adversarial"], cfg.get("adv_scale", 1.0))
        comb = torch.relu(neg - pos.unsqueeze(1) + cfg["margin"])
        return ls * torch.sum(w * torch.sum(nw * comb, dim=-1))
    if kind == "softmax_ce":  # loss.py:226-251 (the reference shifts neg in place)
        adj = neg + (math.log(cfg["n_entity"] - 1) - math.log(neg.shape[1]))
        logits = torch.cat([pos.unsqueeze(1), adj], dim=-1)
        ce = torch.logsumexp(logits, dim=-1) - pos
        return ls * torch.sum(w * ce)
    raise ValueError(kind)


# =============================================================================
# metrics (metric.py)
# =============================================================================
def ranks_from_scores(pos: torch.Tensor, cand: torch.Tensor, mode: str,
                      worst_rank_infty: bool) -> torch.Tensor:
    """metric.py:129-183."""
    n_neg = cand.shape[1]
    # metric.py:151 — IN PLACE on (a view of) the caller's tensor, and with torch's defaults for
    # the infinities: nan -> -inf, but +inf / -inf -> +FLT_MAX / -FLT_MAX.  A positive whose score
    # was masked to -inf therefore still outranks -inf candidates, and the pipeline (which
    # restores the true scores afterwards, pipeline.py:296-301) returns -FLT_MAX for it.
    pos = pos.reshape(-1, 1)
    pos.nan_to_num_(-torch.inf)
    gt = torch.sum(cand > pos, dim=-1).float()
    ge = torch.sum(cand >= pos, dim=-1).float()
    if mode == "optimistic":
        better, worst = gt, gt == n_neg
    elif mode == "pessimistic":
        better, worst = ge, ge == n_neg
    else:
        better, worst = 0.5 * (gt + ge), (gt == n_neg) | (ge == n_neg)
    rank = 1.0 + better
    if worst_rank_infty:
        rank[worst] = torch.inf
    return rank


def ranks_from_indices(truth: torch.Tensor, ids: torch.Tensor,
                       worst_rank_infty: bool) -> torch.Tensor:
    """metric.py:185-220."""
    n = ids.shape[1]
    worst = torch.inf if worst_rank_infty else float(n + 1)
    pos = torch.arange(1, n + 1, dtype=torch.float32, device=ids.device)
    return torch.where(truth.reshape(-1, 1) == ids, pos, worst).min(dim=-1)[0]


# =============================================================================
# the sharded step, evaluated with the explicit routing rule
# =============================================================================
def embedding_moving_forward(
    cfg: Dict[str, Any],
    ent: torch.Tensor,  # [n, Es, W]
    rel_table: torch.Tensor,
    head: torch.Tensor,  # [n, n, p]      (one step; host layout without the bps axis)
    relation: torch.Tensor,  # [n, n, p]
    tail: torch.Tensor,  # [n(shard_t), n(shard_h), p]
    negative: torch.Tensor,  # [n(src), n(dst), B, Nn]
    scheme: str,
    flat: bool,
    shared: bool,
    negative_mask: Optional[torch.Tensor] = None,  # [n, B, n, Nn] (per scoring replica)
    augment: bool = False,
    local_sampling: bool = False,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """bess.py:322-468 + the mask handling of bess.py:182-245.  Returns
    (positive [n, S], negative [n, S, N]) for the n replicas."""
    n, _, p = head.shape
    B, Nn = negative.shape[-2:]
    W = ent.shape[-1]
    pos_all, neg_all = [], []
    for r in range(n):
        h_emb = ent[r][head[r].long()]  # [n, p, W]
        t_emb = torch.stack([ent[j][tail[j, r].long()] for j in range(n)])  # [n, p, W]
        if local_sampling:
            neg_emb = ent[r][negative[r].long()]  # [n, B, Nn, W]
        else:
            neg_emb = torch.stack([ent[j][negative[j, r].long()] for j in range(n)])
        neg_emb = neg_emb.transpose(0, 1).flatten(start_dim=1, end_dim=2)  # [B, n*Nn, W]
        rel = relation[r]
        hf, tf, rf = h_emb.flatten(end_dim=1), t_emb.flatten(end_dim=1), rel.flatten()
        pos = score_triple(cfg, hf, rel_table, rf, tf)
        if scheme == "h":
            cand = neg_emb
            if augment:
                cand = torch.cat([h_emb.reshape(cand.shape[0], -1, W), cand], dim=1)
            neg = score_candidates(cfg, "h", tf, rel_table, rf, cand, shared)
        elif scheme == "t":
            cand = neg_emb
            if augment:
                cand = torch.cat([t_emb.reshape(cand.shape[0], -1, W), cand], dim=1)
            neg = score_candidates(cfg, "t", hf, rel_table, rf, cand, shared)
        else:
            cut = p // 2
            if flat:
                nh, nt = neg_emb[0:1], neg_emb[1:2]
            else:
                ne = neg_emb.reshape(n, p, -1, W)
                nh, nt = ne[:, :cut].flatten(end_dim=1), ne[:, cut:].flatten(end_dim=1)
            if augment:
                nh = torch.cat([h_emb[:, :cut].reshape(nh.shape[0], -1, W), nh], dim=1)
                nt = torch.cat([t_emb[:, cut:].reshape(nt.shape[0], -1, W), nt], dim=1)
            s1 = score_candidates(cfg, "h", t_emb[:, :cut].flatten(end_dim=1), rel_table,
                                  rel[:, :cut].flatten(), nh, shared)
            s2 = score_candidates(cfg, "t", h_emb[:, cut:].flatten(end_dim=1), rel_table,
                                  rel[:, cut:].flatten(), nt, shared)
            neg = torch.cat([s1.reshape(n, cut, -1), s2.reshape(n, p - cut, -1)], dim=1).flatten(
                end_dim=1)
        neg = apply_masks(neg, relation[r], negative.shape, scheme, flat, augment,
                          None if negative_mask is None else negative_mask[r])
        pos_all.append(pos)
        neg_all.append(neg)
    return torch.stack(pos_all), torch.stack(neg_all)


def apply_masks(neg: torch.Tensor, relation_r: torch.Tensor, neg_shape, scheme: str, flat: bool,
                augment: bool, nmask: Optional[torch.Tensor]) -> torch.Tensor:
    """bess.py:182-245 for one replica; nmask [B, n, Nn] bool (True = real)."""
    n, p = relation_r.shape
    if nmask is not None:
        nmask = nmask.flatten(start_dim=-2)
        if flat and scheme == "ht":
            cut = p // 2
            nmask = torch.cat([nmask[0:1].expand(n, cut, -1),
                               nmask[1:2].expand(n, p - cut, -1)], dim=1).flatten(end_dim=1)
    if augment:
        step = 1 if flat else 1 + neg_shape[0] * neg_shape[-1]
        aug = (torch.arange(neg.shape[1], device=neg.device)[None, :]
               == step * torch.arange(neg.shape[0], device=neg.device)[:, None])
        if scheme == "ht":
            aug = aug[: aug.shape[0] // 2].reshape(n, p // 2, -1).repeat(1, 2, 1).flatten(end_dim=1)
        if nmask is not None:
            aug[:, -nmask.shape[1]:] = ~nmask
        return neg + BAD_NEGATIVE_SCORE * aug
    if nmask is not None:
        return neg + BAD_NEGATIVE_SCORE * (~nmask)
    return neg


def score_moving_forward(
    cfg: Dict[str, Any], ent: torch.Tensor, rel_table: torch.Tensor, head: torch.Tensor,
    relation: torch.Tensor, tail: torch.Tensor, negative: torch.Tensor, scheme: str, flat: bool,
    shared: bool, triple_based: bool, negative_mask: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """bess.py:490-603: queries are replicated, scored against the LOCAL
    negatives of every shard, scores routed back.  negative [n(src), n(dst), B, Nn]."""
    n, _, p = head.shape
    S = n * p
    # per replica: gathered heads / tails-for-others
    h_all = torch.stack([ent[r][head[r].long()] for r in range(n)])  # [n, n, p, W]
    t_loc = torch.stack([ent[r][tail[r].long()] for r in range(n)])  # [n(shard_t), n(shard_h), p, W]
    t_recv = t_loc.transpose(0, 1)  # [n(replica), n(src), p, W] == all_to_all result
    cut = p // 2
    per_src: List[torch.Tensor] = []
    for r in range(n):  # scoring shard
        ne = ent[r][negative[r].long()]  # [n(dst), B, Nn, W]
        if triple_based and flat:
            ne = ne[0].unsqueeze(0)
        rel_all = relation  # [n(replica), n, p]
        if scheme == "h":
            t_q = t_loc.transpose(0, 1)  # all_gather(tail).transpose(0,1): [j, r', p, W]
            sc = score_candidates(cfg, "h", t_q.flatten(end_dim=2), rel_table, rel_all.flatten(),
                                  ne.flatten(end_dim=1), shared)
        elif scheme == "t":
            sc = score_candidates(cfg, "t", h_all.flatten(end_dim=2), rel_table, rel_all.flatten(),
                                  ne.flatten(end_dim=1), shared)
        else:
            t_q = t_loc[:, :, :cut].transpose(0, 1)
            h_q = h_all[:, :, cut:]
            if flat:
                nh, nt = ne[:, 0:1].flatten(end_dim=1), ne[:, 1:2].flatten(end_dim=1)
            else:
                ne5 = ne.reshape(n, n, p, -1, ne.shape[-1])
                nh, nt = ne5[:, :, :cut].flatten(end_dim=2), ne5[:, :, cut:].flatten(end_dim=2)
            s1 = score_candidates(cfg, "h", t_q.flatten(end_dim=2), rel_table,
                                  rel_all[:, :, :cut].flatten(), nh, shared)
            s2 = score_candidates(cfg, "t", h_q.flatten(end_dim=2), rel_table,
                                  rel_all[:, :, cut:].flatten(), nt, shared)
            sc = torch.cat([s1.reshape(n, n, cut, -1), s2.reshape(n, n, p - cut, -1)],
                           dim=2).flatten(end_dim=2)
        per_src.append(sc.reshape(n, S, -1))  # [replica j, S, X]
    pos_all, neg_all = [], []
    for j in range(n):
        neg = torch.stack([per_src[r][j] for r in range(n)]).transpose(0, 1).flatten(start_dim=1)
        pos = score_triple(cfg, h_all[j].flatten(end_dim=1), rel_table, relation[j].flatten(),
                           t_recv[j].flatten(end_dim=1))
        if negative_mask is not None:
            neg = apply_masks(neg, relation[j], negative.shape, scheme, flat, False,
                              negative_mask[j])
        pos_all.append(pos)
        neg_all.append(neg)
    return torch.stack(pos_all), torch.stack(neg_all)


def topk_forward(cfg: Dict[str, Any], ent: torch.Tensor, rel_table: torch.Tensor,
                 sh: Dict[str, Any], relation: torch.Tensor, fixed_idx: torch.Tensor, scheme: str,
                 k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """bess.py:691-921 against ALL entities, un-windowed: queries of replica r
    (relation [n, S], local ids fixed_idx [n, S]) -> top-k global ids/scores.
    Padding rows of each shard are excluded exactly as bess.py:866-873 does."""
    n, S = relation.shape
    Es = ent.shape[1]
    ids_out, sc_out = [], []
    for r in range(n):
        q = ent[r][fixed_idx[r].long()]
        scores, gids = [], []
        for j in range(n):
            cnt = int(sh["shard_counts"][j])
            cand = ent[j][:cnt].unsqueeze(0)
            mode = "t" if scheme == "t" else "h"
            scores.append(score_candidates(cfg, mode, q, rel_table, relation[r], cand, True))
            gids.append(torch.from_numpy(sh["shard_and_idx_to_entity"][j][:cnt].astype(np.int64)))
        sc = torch.cat(scores, dim=1)
        gi = torch.cat(gids)
        top = torch.topk(sc, k=k, dim=1)
        sc_out.append(top.values)
        ids_out.append(gi[top.indices])
    return torch.stack(ids_out), torch.stack(sc_out)


def exact_scores(cfg: Dict[str, Any], mode: str, fixed: torch.Tensor, rel_rows: torch.Tensor,
                 cand: torch.Tensor) -> torch.Tensor:
    """Scores of queries (fixed [Q, W] entity rows + rel_rows [Q, Wr]) against cand [C, W] in
    the ONE fixed fp32 arithmetic that defines exact ranking (csrc/exact.cu, SURVEY.md 7.2
    item 3): the score_tails / score_heads expressions of scoring.py:335-356 (TransE),
    :815-840 (DistMult), :918-946 (ComplEx) evaluated coordinate by coordinate, k = 0..W-1,
    every product and every sum a separate fp32 rounding (each torch op below rounds once;
    nothing is fused), IEEE sqrt.  Pure torch: runs on whatever device the inputs live on."""
    fam, p = cfg["family"], cfg.get("norm_p", 2)
    tails = mode == "t"
    f, r, c = fixed.float(), rel_rows.float(), cand.float()
    W = c.shape[1]
    if fam == "DistMult":
        q = f * r
    elif fam == "ComplEx":
        e = W // 2
        f_re, f_im, r_re, r_im = f[:, :e], f[:, e:], r[:, :e], r[:, e:]
        if tails:
            q = torch.cat([f_re * r_re - f_im * r_im, f_re * r_im + f_im * r_re], dim=1)
        else:
            n_im = -r_im
            q = torch.cat([r_re * f_re - n_im * f_im, r_re * f_im + n_im * f_re], dim=1)
    elif fam == "TransE":
        q = f + r if tails else f - r
    else:
        raise ValueError(f"no bit-reproducible arithmetic for {fam}")
    acc = torch.zeros(q.shape[0], c.shape[0], dtype=torch.float32, device=c.device)
    ct = c.t().contiguous()
    for k in range(W):
        if fam == "TransE":
            d = q[:, k:k + 1] - ct[k][None, :]
            term = d.abs() if p == 1 else d * d
        else:
            term = q[:, k:k + 1] * ct[k][None, :]
        acc = acc + term
    if fam == "TransE":
        return -(acc if p == 1 else torch.sqrt(acc))
    return acc


def topk_exact(cfg: Dict[str, Any], ent: torch.Tensor, rel_table: torch.Tensor,
               sh: Dict[str, Any], relation: torch.Tensor, fixed_idx: torch.Tensor, scheme: str,
               k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """`topk_forward` under the exact-ranking arithmetic: `exact_scores` against every real
    entity, ordered by (score descending, shard ascending, local id ascending) — the order
    the product's per-shard lists and final merge produce.  Returns (global ids, scores)
    [n, S, k]; the product must reproduce both BIT FOR BIT."""
    n, S = relation.shape
    dev = ent.device
    ids_out, sc_out = [], []
    for r in range(n):
        fixed = ent[r][fixed_idx[r].long()]
        rel_rows = rel_table[relation[r].long()]
        scores, gids = [], []
        for j in range(n):
            cnt = int(sh["shard_counts"][j])
            scores.append(exact_scores(cfg, scheme, fixed, rel_rows, ent[j][:cnt]))
            gids.append(torch.from_numpy(sh["shard_and_idx_to_entity"][j][:cnt].astype(np.int64)))
        sc = torch.cat(scores, dim=1)
        order = torch.sort(sc, dim=1, descending=True, stable=True)
        sc_out.append(order.values[:, :k])
        ids_out.append(torch.cat(gids).to(dev)[order.indices[:, :k]])
    return torch.stack(ids_out), torch.stack(sc_out)


def all_scores_pipeline(cfg: Dict[str, Any], ent_unsharded: torch.Tensor, rel_table: torch.Tensor,
                        triples: torch.Tensor, scheme: str, filter_triples: Optional[torch.Tensor],
                        candidate_ents: Optional[np.ndarray], k: int, mode: str = "average"
                        ) -> Dict[str, torch.Tensor]:
    """pipeline.py:192-320 restated densely for triples given with GLOBAL ids [x, 3]:
    scores of every completion against the un-sharded table, -inf on non-candidates
    (:262-265), true scores read (:266-271), -inf on filtered completions (:272-279,
    utils.py:36-69), ranks with the true completion masked (:285-292), true scores
    restored (:296-301), top-k (:304-309).  Row order = order of `triples`."""
    x = triples.shape[0]
    gt_col = 0 if scheme == "h" else 2
    fixed = ent_unsharded[triples[:, 2 - gt_col].long()]
    sc = score_candidates(cfg, "t" if scheme == "t" else "h", fixed, rel_table, triples[:, 1],
                          ent_unsharded.unsqueeze(0), True).clone()
    ar = torch.arange(x)
    truth = triples[:, gt_col].long()
    if candidate_ents is not None:
        drop = np.setdiff1d(np.arange(ent_unsharded.shape[0]), candidate_ents)
        sc[:, torch.from_numpy(drop)] = -torch.inf
    true = sc[ar, truth].clone()
    if filter_triples is not None:
        ent_col = 0 if scheme == "t" else 2
        hit = ((filter_triples[:, 1] == triples[:, 1].view(-1, 1))
               & (filter_triples[:, ent_col] == triples[:, ent_col].view(-1, 1))).nonzero()
        sc[hit[:, 0], filter_triples[hit[:, 1], 2 - ent_col].long()] = -torch.inf
    sc[ar, truth] = -torch.inf
    ranks = ranks_from_scores(true, sc, mode, False)
    sc[ar, truth] = true
    top = torch.topk(sc, k=k, dim=1).indices
    return dict(scores=sc, ranks=ranks, topk_global_id=top, true_scores=true)


# =============================================================================
# training: forward -> torch autograd -> dense torch.optim
# =============================================================================
def training_steps(
    cfg: Dict[str, Any], loss_cfg: Dict[str, Any], opt_cfg: Dict[str, Any], ent: torch.Tensor,
    rel_table: torch.Tensor, batches: List[Dict[str, torch.Tensor]], scheme: str, flat: bool,
    shared: bool, relation_grad_reduction: str = "mean", augment: bool = False,
    weights: Optional[List[Optional[torch.Tensor]]] = None,
    lr_schedule: Optional[List[float]] = None, accumulate: int = 1,
    accumulation_reduction: str = "mean", model: str = "embedding_moving",
    triple_based: bool = False,
) -> Dict[str, Any]:
    """Runs len(batches) micro-batch steps (each: n replicas, summed losses, one
    optimizer step).  Entity-table gradients are the plain sum over replicas
    (no all-reduce on the sharded table, custom_ops/remove_all_reduce_pattern.cpp);
    the replicated relation table's gradient is reduced by `relation_grad_reduction`.
    `lr_schedule[i]`: learning rate of optimizer step i.  `accumulate = k`: gradients of k
    consecutive micro-batches are accumulated (all from the same weights) and reduced by
    `accumulation_reduction` before ONE optimizer step — PopTorch's
    `options.Training.gradientAccumulation(k)` (notebook 1 cell 26: 6, notebook 2 cell 14: 2)."""
    ent = ent.clone().requires_grad_(True)
    rel_table = rel_table.clone().requires_grad_(True)
    n = ent.shape[0]
    if opt_cfg["kind"] == "sgd":
        mk = lambda ps: torch.optim.SGD(ps, lr=opt_cfg["lr"], momentum=opt_cfg.get("momentum", 0.0),
                                        dampening=opt_cfg.get("dampening", 0.0),
                                        weight_decay=opt_cfg.get("weight_decay", 0.0))
    else:
        mk = lambda ps: torch.optim.AdamW(ps, lr=opt_cfg["lr"], betas=opt_cfg.get("betas", (0.9, 0.999)),
                                          eps=opt_cfg.get("eps", 1e-8),
                                          weight_decay=opt_cfg.get("weight_decay", 1e-2))
    opt_e, opt_r = mk([ent]), mk([rel_table])
    losses, grads_e, grads_r = [], [], []
    for bi, b in enumerate(batches):
        if bi % accumulate == 0:
            opt_e.zero_grad(set_to_none=True)
            opt_r.zero_grad(set_to_none=True)
            rel_acc = None
        if model == "score_moving":  # bess.py:471-603 (ScoreMovingBessKGE is a full BessKGE)
            pos, neg = score_moving_forward(cfg, ent, rel_table, b["head"], b["relation"],
                                            b["tail"], b["negative"], scheme, flat, shared,
                                            triple_based, b.get("negative_mask"))
        else:
            pos, neg = embedding_moving_forward(cfg, ent, rel_table, b["head"], b["relation"],
                                                b["tail"], b["negative"], scheme, flat, shared,
                                                b.get("negative_mask"), augment)
        step_losses = []
        for r in range(n):
            w = torch.tensor([1.0]) if weights is None or weights[bi] is None else weights[bi][r]
            step_losses.append(loss_value(loss_cfg, pos[r].float(), neg[r].float(), w))
        before = None if rel_table.grad is None else rel_table.grad.detach().clone()
        torch.stack(step_losses).sum().backward()
        if relation_grad_reduction == "mean":  # only this micro-batch's contribution
            cur = rel_table.grad if before is None else rel_table.grad - before
            rel_table.grad = (cur / n) if before is None else before + cur / n
        losses.append(torch.stack(step_losses).detach())
        if (bi + 1) % accumulate != 0:
            continue
        if accumulate > 1 and accumulation_reduction == "mean":
            ent.grad.div_(accumulate)
            rel_table.grad.div_(accumulate)
        grads_e.append(ent.grad.detach().clone())
        grads_r.append(rel_table.grad.detach().clone())
        if lr_schedule is not None:
            for o in (opt_e, opt_r):
                for gp in o.param_groups:
                    gp["lr"] = lr_schedule[len(grads_e) - 1]
        opt_e.step()
        opt_r.step()
    return dict(loss=torch.stack(losses), ent=ent.detach(), rel=rel_table.detach(),
                grad_ent=grads_e, grad_rel=grads_r)
