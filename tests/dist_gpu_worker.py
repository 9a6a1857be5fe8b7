"""Multi-GPU parity worker (launch with torchrun, one rank per GPU, NCCL):
distributed EmbeddingMoving training steps vs the CPU oracle.  Run by
scripts/gpu_multi.sh on an N-GPU box; not collected by pytest."""
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
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, n = dist.get_rank(), dist.get_world_size()
    for fam, p, scheme, flat, shared, lkind in [
        ("TransE", 1, "t", True, True, "logsigmoid"),
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
        if rank == 0:
            print(f"distributed parity ok: {fam} {scheme} flat={flat} n={n}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
