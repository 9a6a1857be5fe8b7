"""-m gpu: parity beyond the committed small fixtures.

* BASELINE.json configs[1] at its full per-GPU size (biokg shard, DistMult d=256
  fp32, shard_bs 16384, 2048 shared tail negatives): one training step through
  the product path against the oracle's arithmetic evaluated in fp64 (the
  oracle is pure torch, so the test runs it on the GPU in double precision to
  finish in seconds), plus bit-determinism of the step.
* bf16 / fp16 tables (configs[3]: TransE-L1 bf16/fp16; DistMult / ComplEx through
  the kind::f16 tensor-core path): product vs the oracle in fp32 started from
  the same rounded tables, tolerance 1e-2 (north_star).
"""
import pytest
import torch
from torch.testing import assert_close

from oracle import besskge_oracle as O

pytestmark = pytest.mark.gpu


def _imports():
    import besskge_b200 as B
    from . import gpu_helpers as H
    return B, H


def _cfg2_problem(S, N, E=93773, R=51, d=256, seed=0):
    g = torch.Generator().manual_seed(seed)
    ent = torch.randn(1, E, d, generator=g) * 0.3
    rel = torch.randn(R, d, generator=g) * 0.3
    batch = dict(
        head=torch.randint(E, (1, 1, S), generator=g, dtype=torch.int32),
        tail=torch.randint(E, (1, 1, S), generator=g, dtype=torch.int32),
        relation=torch.randint(R, (1, 1, S), generator=g, dtype=torch.int32),
        negative=torch.randint(E, (1, 1, 1, N), generator=g, dtype=torch.int32),
    )
    return ent, rel, batch


def test_full_size_cfg2_training_step_vs_fp64_and_determinism():
    B, H = _imports()
    from besskge_b200.bess import EmbeddingMovingBessKGE, training_model
    from besskge_b200.optim import SGD
    from besskge_b200.sharding import Sharding
    S, N, E, R, d, lr = 16384, 2048, 93773, 51, 256, 0.5
    sh = Sharding.create(E, 1, seed=1234)
    ent, rel, batch = _cfg2_problem(S, N, E, R, d)
    lcfg = dict(kind="logsigmoid", margin=2.0, negative_adversarial_sampling=True)

    tables = []
    for _ in range(2):
        sf = H.make_score_fn("DistMult", True, 2, sh, R, d, ent, rel)
        model = EmbeddingMovingBessKGE(H.fake_sampler("t", True, triple_based=False), sf,
                                       loss_fn=H.make_loss(lcfg), return_scores=True)
        step = training_model(model, SGD(lr=lr), cuda_graph=False)
        res = step(**batch)
        torch.cuda.synchronize()
        tables.append((sf.entity_embedding.detach().clone(), sf.relation_embedding.detach().clone(),
                       res["loss"].clone(), res["positive_score"].clone(),
                       res["negative_score"].clone()))
    # same inputs, same state -> same bits (stable sort + fixed-order segment sums, fixed
    # split-K reduction order)
    for a, b in zip(tables[0], tables[1]):
        assert torch.equal(a, b)

    # the oracle's arithmetic in fp64 on the GPU
    ent64 = ent.double().cuda().requires_grad_(True)
    rel64 = rel.double().cuda().requires_grad_(True)
    dev_batch = {k: v.cuda() for k, v in batch.items()}
    pos, neg = O.embedding_moving_forward(H.score_cfg("DistMult", d, 2), ent64, rel64,
                                          dev_batch["head"], dev_batch["relation"],
                                          dev_batch["tail"], dev_batch["negative"], "t", True, True)
    loss = O.loss_value(H.oracle_loss_cfg(lcfg), pos[0], neg[0], torch.ones(1, dtype=torch.float64,
                                                                           device="cuda"))
    loss.backward()
    new_ent = (ent64 - lr * ent64.grad).detach().float()
    new_rel = (rel64 - lr * rel64.grad).detach().float()
    t_ent, t_rel, t_loss, t_pos, t_neg = tables[0]
    # scores: 3xTF32 products, fp32 accumulate -> bounded by 4e-6 * sum |a_k b_k|
    assert_close(t_pos, pos[0].float(), rtol=1e-5, atol=2e-5)
    assert_close(t_neg, neg[0].float(), rtol=1e-5, atol=2e-5)
    assert_close(t_loss[0].double(), loss.detach(), rtol=1e-5, atol=0)
    assert_close(t_ent, new_ent, rtol=1e-5, atol=2e-6)
    # each relation row sums ~S/R = 321 per-triple gradients of either sign (sum of magnitudes
    # ~ 15 x lr): fp32 accumulation, here and in the reference's own fp32 autograd, is good to
    # ~1e-6 of that sum of magnitudes, not of the (cancelled) result -> absolute floor 2e-5
    assert_close(t_rel, new_rel, rtol=1e-5, atol=2e-5)
    # rows that received no gradient are bit-untouched
    touched = torch.zeros(E, dtype=torch.bool)
    for k in ("head", "tail", "negative"):
        touched[batch[k].flatten().long()] = True
    assert torch.equal(t_ent[0][~touched.cuda()], ent[0].cuda()[~touched.cuda()])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("fam,p,scheme", [("DistMult", 2, "t"), ("ComplEx", 2, "h"),
                                          ("TransE", 1, "t"), ("RotatE", 1, "ht")])
def test_training_half_tables_vs_fp32_oracle(dtype, fam, p, scheme):
    B, H = _imports()
    from besskge_b200.bess import EmbeddingMovingBessKGE, training_model
    from besskge_b200.optim import SGD
    from besskge_b200.sharding import Sharding
    n, p_part, Nn, d, n_rel, n_ent = 2, 8, 24, 32, 5, 120
    sh = Sharding.create(n_ent, n, seed=3)
    gen = torch.Generator().manual_seed(5)
    ew = 2 if fam in ("RotatE", "ComplEx") else 1
    rw = 2 * d if fam == "ComplEx" else d
    # tables exactly representable in the table dtype: both sides start from the same numbers
    ent = (torch.randn(n, sh.max_entity_per_shard, ew * d, generator=gen) * 0.5).to(dtype).float()
    rel = (torch.randn(n_rel, rw, generator=gen) * 0.5).to(dtype).float()
    Bn = 2 if scheme == "ht" else 1
    lo = int(sh.shard_counts.min())
    batches = [dict(
        head=torch.randint(lo, (n, n, p_part), generator=gen, dtype=torch.int32),
        tail=torch.randint(lo, (n, n, p_part), generator=gen, dtype=torch.int32),
        relation=torch.randint(n_rel, (n, n, p_part), generator=gen, dtype=torch.int32),
        negative=torch.randint(lo, (n, n, Bn, Nn), generator=gen, dtype=torch.int32))]
    lcfg = dict(kind="logsigmoid", margin=3.0, negative_adversarial_sampling=True)
    want = O.training_steps(H.score_cfg(fam, d, p), H.oracle_loss_cfg(lcfg), dict(kind="sgd", lr=0.05),
                            ent, rel, batches, scheme, True, True, "mean")
    sf = H.make_score_fn(fam, True, p, sh, n_rel, d, ent, rel, dtype=dtype)
    model = EmbeddingMovingBessKGE(H.fake_sampler(scheme, True, triple_based=False), sf,
                                   loss_fn=H.make_loss(lcfg), return_scores=True)
    step = training_model(model, SGD(lr=0.05))
    res = step(**batches[0])
    torch.cuda.synchronize()
    assert res["positive_score"].dtype == dtype
    assert_close(res["loss"].cpu(), want["loss"][0], rtol=1e-2, atol=1e-2)
    assert_close(sf.entity_embedding.detach().float().cpu(), want["ent"], rtol=1e-2, atol=2e-3)
    assert_close(sf.relation_embedding.detach().float().cpu(), want["rel"], rtol=1e-2, atol=2e-3)
