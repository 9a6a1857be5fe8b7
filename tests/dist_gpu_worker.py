"""Multi-rank parity worker (launched with torchrun by tests/test_gpu_distributed.py, or by
hand on an N-GPU box): distributed EmbeddingMoving training steps — the NVLink peer-memory
exchange: routed remote gather, flag handshakes, gradient push, relation all-reduce, all
inside one captured CUDA graph per rank — vs the CPU oracle, then the distributed inference
modules vs their local-mode results.

    one rank per GPU (NCCL):   torchrun --nproc-per-node N tests/dist_gpu_worker.py
    BESS_TEST_SAME_DEVICE=1:   every rank on cuda:0 (a 1-GPU box), rendezvous over gloo; the
                               symmetric-memory mapping and the kernels are the same, peers are
                               other processes time-sliced on the same GPU.  NCCL cannot put two
                               ranks on one device, so the inference modules (which use NCCL
                               collectives) are skipped in this mode."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from besskge_b200.bess import EmbeddingMovingBessKGE, training_model  # noqa: E402
from besskge_b200.optim import SGD  # noqa: E402
from besskge_b200.sharding import Sharding  # noqa: E402
from oracle import besskge_oracle as O  # noqa: E402
from tests import gpu_helpers as H  # noqa: E402


def main() -> None:
    same_device = os.environ.get("BESS_TEST_SAME_DEVICE", "0") == "1"
    local = 0 if same_device else int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    if same_device:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, n = dist.get_rank(), dist.get_world_size()
    for fam, p, scheme, flat, shared, lkind in [
        ("TransE", 1, "t", True, True, "logsigmoid"),
        ("DistMult", 2, "t", True, True, "logsigmoid"),  # tensor-core path with the early gradient push
        ("DistMult", 2, "ht", True, True, "margin_ranking"),
        ("RotatE", 2, "h", False, False, "logsigmoid"),
    ]:
        d, n_rel, n_ent, p_part, Nn = 32, 5, 50 * n, 8, 6
        sh = Sharding.create(n_ent, n, seed=3)
        gen = torch.Generator().manual_seed(11)
        ew = 2 if fam in ("RotatE", "ComplEx") else 1
        ent = torch.randn(n, sh.max_entity_per_shard, ew * d, generator=gen) * 0.5
        rel = torch.randn(n_rel, d, generator=gen) * 0.5
        S = n * p_part
        Bn = (2 if scheme == "ht" else 1) if flat else S
        lo = int(sh.shard_counts.min())
        batches = []
        for _ in range(3):
            batches.append(dict(
                head=torch.randint(lo, (n, n, p_part), generator=gen, dtype=torch.int32),
                tail=torch.randint(lo, (n, n, p_part), generator=gen, dtype=torch.int32),
                relation=torch.randint(n_rel, (n, n, p_part), generator=gen, dtype=torch.int32),
                negative=torch.randint(lo, (n, n, Bn, Nn), generator=gen, dtype=torch.int32)))
        lcfg = dict(kind=lkind, margin=2.0, negative_adversarial_sampling=True)
        want = O.training_steps(H.score_cfg(fam, d, p), H.oracle_loss_cfg(lcfg),
                                dict(kind="sgd", lr=0.1), ent, rel, batches, scheme, flat, shared,
                                "mean")
        sf = H.make_score_fn(fam, shared, p, sh, n_rel, d, ent, rel)
        model = EmbeddingMovingBessKGE(H.fake_sampler(scheme, flat, triple_based=False), sf,
                                       loss_fn=H.make_loss(lcfg))
        step = training_model(model, SGD(lr=0.1))
        for s, b in enumerate(batches):
            res = step(**b)
            torch.testing.assert_close(res["loss"].cpu(), want["loss"][s][rank:rank + 1],
                                       rtol=1e-5, atol=1e-4)
        torch.cuda.synchronize()
        torch.testing.assert_close(sf.entity_embedding.detach()[rank].cpu(), want["ent"][rank],
                                   rtol=1e-5, atol=2e-6)
        torch.testing.assert_close(sf.relation_embedding.detach().cpu(), want["rel"], rtol=1e-5,
                                   atol=2e-6)
        assert step.cuda_graph and len([g for g in step._graphs.values() if g != "warm"]) == 1, \
            "the distributed step was not captured / replayed"
        if rank == 0:
            print(f"distributed parity ok: {fam} {scheme} flat={flat} n={n}", flush=True)
    if not same_device:
        score_moving_training(rank, n)
        inference_parity(rank, n)
    dist.barrier()
    dist.destroy_process_group()


def score_moving_training(rank: int, n: int) -> None:
    """Distributed ScoreMovingBessKGE training steps (score gradients back to the scoring ranks,
    query gradients summed at the owner, candidate rows updated in place) vs the oracle."""
    from besskge_b200.bess import ScoreMovingBessKGE
    from besskge_b200.optim import AdamW
    for fam, p, scheme, flat, shared, opt_kind in [("TransE", 1, "t", True, True, "sgd"),
                                                   ("RotatE", 1, "h", False, False, "sgd"),
                                                   ("DistMult", 2, "ht", True, True, "adamw")]:
        d, n_rel, n_ent, p_part, Nn = 16, 5, 50 * n, 6, 5
        sh = Sharding.create(n_ent, n, seed=3)
        gen = torch.Generator().manual_seed(13)
        ew = 2 if fam in ("RotatE", "ComplEx") else 1
        ent = torch.randn(n, sh.max_entity_per_shard, ew * d, generator=gen) * 0.5
        rel = torch.randn(n_rel, d, generator=gen) * 0.5
        S = n * p_part
        Bn = (2 if scheme == "ht" else 1) if flat else S
        lo = int(sh.shard_counts.min())
        batches = [dict(
            head=torch.randint(lo, (n, n, p_part), generator=gen, dtype=torch.int32),
            tail=torch.randint(lo, (n, n, p_part), generator=gen, dtype=torch.int32),
            relation=torch.randint(n_rel, (n, n, p_part), generator=gen, dtype=torch.int32),
            negative=torch.randint(lo, (n, n, Bn, Nn), generator=gen, dtype=torch.int32))
            for _ in range(2)]
        lcfg = dict(kind="logsigmoid", margin=2.0, negative_adversarial_sampling=True)
        ocfg = dict(kind="sgd", lr=0.1) if opt_kind == "sgd" else dict(kind="adamw", lr=0.01, eps=1e-4)
        want = O.training_steps(H.score_cfg(fam, d, p), H.oracle_loss_cfg(lcfg), ocfg, ent, rel,
                                batches, scheme, flat, shared, "mean", model="score_moving")
        sf = H.make_score_fn(fam, shared, p, sh, n_rel, d, ent, rel)
        model = ScoreMovingBessKGE(H.fake_sampler(scheme, flat, triple_based=False), sf,
                                   loss_fn=H.make_loss(lcfg))
        step = training_model(model, SGD(lr=0.1) if opt_kind == "sgd" else AdamW(lr=0.01, eps=1e-4))
        for s, b in enumerate(batches):
            res = step(**b)
            torch.testing.assert_close(res["loss"].cpu(), want["loss"][s][rank:rank + 1],
                                       rtol=2e-5, atol=1e-4)
        torch.cuda.synchronize()
        tol = dict(rtol=1e-5, atol=2e-6) if opt_kind == "sgd" else dict(rtol=1e-4, atol=2e-5)
        torch.testing.assert_close(sf.entity_embedding.detach()[rank].cpu(), want["ent"][rank], **tol)
        torch.testing.assert_close(sf.relation_embedding.detach().cpu(), want["rel"], **tol)
        if rank == 0:
            print(f"distributed ScoreMoving training ok: {fam} {scheme} flat={flat} n={n}", flush=True)


def inference_parity(rank: int, n: int) -> None:
    """ScoreMovingBessKGE and TopKQueryBessKGE: the distributed result of this rank equals
    rows [rank] of the same module run in local mode (all shards on this GPU)."""
    import besskge_b200.bess as bess_mod
    from besskge_b200.bess import ScoreMovingBessKGE, TopKQueryBessKGE
    from besskge_b200.metric import Evaluation
    from besskge_b200.negative_sampler import PlaceholderNegativeSampler

    d, n_rel, n_ent, p_part, Nn, k = 32, 5, 60 * n, 6, 9, 4
    S = n * p_part
    sh = Sharding.create(n_ent, n, seed=3)
    lo = int(sh.shard_counts.min())
    for fam, p, scheme, flat in [("TransE", 1, "t", True), ("DistMult", 2, "ht", True),
                                 ("RotatE", 1, "h", False), ("PairRE", 1, "t", False)]:
        gen = torch.Generator().manual_seed(23)
        ew = 2 if fam in ("RotatE", "ComplEx") else 1
        rw = 2 * d if fam in ("ComplEx", "PairRE") else d
        ent = torch.randn(n, sh.max_entity_per_shard, ew * d, generator=gen) * 0.5
        rel = torch.randn(n_rel, rw, generator=gen) * 0.5
        Bn = (2 if scheme == "ht" else 1) if flat else S
        batch = dict(
            head=torch.randint(lo, (2 * n, n, p_part), generator=gen, dtype=torch.int32),
            tail=torch.randint(lo, (2 * n, n, p_part), generator=gen, dtype=torch.int32),
            relation=torch.randint(n_rel, (2 * n, n, p_part), generator=gen, dtype=torch.int32),
            negative=torch.randint(lo, (2 * n, n, Bn, Nn), generator=gen, dtype=torch.int32))
        results = []
        for force_local in (True, False):
            bess_mod.FORCE_LOCAL = force_local
            sf = H.make_score_fn(fam, flat, p, sh, n_rel, d, ent, rel)
            ev = Evaluation(["mrr", "hits@3"], mode="average", reduction="sum", return_ranks=True)
            model = ScoreMovingBessKGE(H.fake_sampler(scheme, flat, triple_based=False), sf,
                                       evaluation=ev, return_scores=True)
            res = model(**batch)
            torch.cuda.synchronize()
            results.append(res)
        bess_mod.FORCE_LOCAL = False
        loc, dis = results
        rows = torch.cat([torch.arange(S) + (st * n + rank) * S for st in range(2)])
        torch.testing.assert_close(dis["positive_score"].cpu(), loc["positive_score"].cpu()[rows],
                                   rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(dis["negative_score"].cpu(), loc["negative_score"].cpu()[rows],
                                   rtol=1e-6, atol=1e-6)
        assert dis["metrics"].shape[0] == 2
        if rank == 0:
            print(f"distributed ScoreMoving == local: {fam} {scheme} flat={flat} n={n}")

    for fam, p, scheme in [("DistMult", 2, "t"), ("TransE", 1, "h"), ("ComplEx", 2, "t")]:
        gen = torch.Generator().manual_seed(29)
        ew = 2 if fam in ("RotatE", "ComplEx") else 1
        rw = 2 * d if fam in ("ComplEx", "PairRE") else d
        ent = torch.randn(n, sh.max_entity_per_shard, ew * d, generator=gen)
        rel = torch.randn(n_rel, rw, generator=gen)
        q = dict(relation=torch.randint(n_rel, (2 * n, S), generator=gen, dtype=torch.int32),
                 head=torch.randint(lo, (2 * n, S), generator=gen, dtype=torch.int32),
                 tail=torch.randint(lo, (2 * n, S), generator=gen, dtype=torch.int32))
        results = []
        for force_local in (True, False):
            bess_mod.FORCE_LOCAL = force_local
            sf = H.make_score_fn(fam, True, p, sh, n_rel, d, ent, rel)
            model = TopKQueryBessKGE(k=k, candidate_sampler=PlaceholderNegativeSampler(scheme),
                                     score_fn=sf, return_scores=True)
            model.device_window = 16
            res = model(**q)
            torch.cuda.synchronize()
            results.append(res)
        bess_mod.FORCE_LOCAL = False
        loc, dis = results
        rows = torch.cat([torch.arange(S) + (st * n + rank) * S for st in range(2)])
        torch.testing.assert_close(dis["topk_scores"].cpu(), loc["topk_scores"].cpu()[rows],
                                   rtol=1e-6, atol=1e-6)
        assert torch.equal(dis["topk_global_id"].cpu(), loc["topk_global_id"].cpu()[rows])
        if rank == 0:
            print(f"distributed TopKQuery == local: {fam} {scheme} n={n}")

    from besskge_b200.bess import AllScoresBESS
    for fam, p, scheme in [("ComplEx", 2, "t"), ("TransE", 1, "h")]:
        gen = torch.Generator().manual_seed(31)
        ew = 2 if fam in ("RotatE", "ComplEx") else 1
        rw = 2 * d if fam in ("ComplEx", "PairRE") else d
        ent = torch.randn(n, sh.max_entity_per_shard, ew * d, generator=gen)
        rel = torch.randn(n_rel, rw, generator=gen)
        q = dict(relation=torch.randint(n_rel, (2 * n, S), generator=gen, dtype=torch.int32))
        q["head" if scheme == "t" else "tail"] = torch.randint(lo, (2 * n, S), generator=gen,
                                                                dtype=torch.int32)
        results = []
        for force_local in (True, False):
            bess_mod.FORCE_LOCAL = force_local
            sf = H.make_score_fn(fam, True, p, sh, n_rel, d, ent, rel)
            mod = AllScoresBESS(PlaceholderNegativeSampler(scheme), sf, window_size=25)
            mod.device_window = 16
            full = mod.score_all(**q)
            blk = mod(step=torch.full((2 * n, 1), mod.n_step - 1, dtype=torch.int32), **q)
            torch.cuda.synchronize()
            results.append((full.cpu(), blk.cpu()))
        bess_mod.FORCE_LOCAL = False
        (loc_full, loc_blk), (dis_full, dis_blk) = results
        rows = torch.cat([torch.arange(S) + (st * n + rank) * S for st in range(2)])
        torch.testing.assert_close(dis_full, loc_full[rows], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(dis_blk, loc_blk[rows], rtol=1e-6, atol=1e-6)
        if rank == 0:
            print(f"distributed AllScoresBESS == local: {fam} {scheme} n={n}")


if __name__ == "__main__":
    main()
