"""-m gpu: behaviour of the training wrapper around the captured step — the things a CUDA
graph could silently freeze or dangle, each checked against the oracle (dense torch.optim on
the autograd gradient of the reference forward):

  * learning-rate schedules and AdamW's step count under graph replay;
  * workspace / input-buffer re-allocation between replays (a bigger batch, a validation
    forward on the same module) — captured graphs must be dropped, not replayed stale;
  * optimizer state_dict round trip;
  * gradient accumulation over micro-batches (options.Training.gradientAccumulation);
  * the Python-surface helpers of utils.py.
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


def _problem(fam="DistMult", p=2, n=2, p_part=8, Nn=6, d=32, n_rel=5, n_ent=120, n_batch=6, seed=5):
    from besskge_b200.sharding import Sharding
    sh = Sharding.create(n_ent, n, seed=3)
    gen = torch.Generator().manual_seed(seed)
    ew = 2 if fam in ("RotatE", "ComplEx") else 1
    rw = 2 * d if fam == "ComplEx" else d
    ent = torch.randn(n, sh.max_entity_per_shard, ew * d, generator=gen) * 0.5
    rel = torch.randn(n_rel, rw, generator=gen) * 0.5
    lo = int(sh.shard_counts.min())

    def batch(pp, nn):
        return dict(
            head=torch.randint(lo, (n, n, pp), generator=gen, dtype=torch.int32),
            tail=torch.randint(lo, (n, n, pp), generator=gen, dtype=torch.int32),
            relation=torch.randint(n_rel, (n, n, pp), generator=gen, dtype=torch.int32),
            negative=torch.randint(lo, (n, n, 1, nn), generator=gen, dtype=torch.int32))
    return sh, ent, rel, [batch(p_part, Nn) for _ in range(n_batch)], batch


LCFG = dict(kind="logsigmoid", margin=2.0, negative_adversarial_sampling=True)
# Adam with its default eps = 1e-8 turns a gradient of ONE fp32 ulp (the residual of two
# sigmoid-saturated +-0.5 terms that cancel: 1.5e-8 under one rounding, 0 or 3e-8 under another)
# into a step of ~lr/2 — in torch.optim as much as here.  The wrapper tests therefore run Adam
# with eps = 1e-4, where the update is a well-conditioned function of the gradient, and keep
# tight tolerances; the reference-fixture tests (tests/test_gpu_bess.py) cover default eps.
ADAM_EPS = 1e-4


def _model(H, fam, p, sh, n_rel, d, ent, rel):
    from besskge_b200.bess import EmbeddingMovingBessKGE
    sf = H.make_score_fn(fam, True, p, sh, n_rel, d, ent, rel)
    model = EmbeddingMovingBessKGE(H.fake_sampler("t", True, triple_based=False), sf,
                                   loss_fn=H.make_loss(LCFG))
    return sf, model


@pytest.mark.parametrize("opt_kind", ["sgd", "sgdm", "adamw"])
def test_lr_schedule_and_step_count_survive_graph_replay(opt_kind):
    """6 steps with a different learning rate each; steps 3-6 are graph replays.  A frozen lr
    or a frozen AdamW step count would diverge from the oracle."""
    B, H = _imports()
    from besskge_b200.bess import training_model
    from besskge_b200.optim import SGD, AdamW
    fam, p, d, n_rel = "DistMult", 2, 32, 5
    sh, ent, rel, batches, _ = _problem(fam, p, d=d, n_rel=n_rel)
    lrs = [0.1, 0.05, 0.2, 0.01, 0.15, 0.07]
    ocfg = {"sgd": dict(kind="sgd", lr=lrs[0]),
            "sgdm": dict(kind="sgd", lr=lrs[0], momentum=0.9),
            "adamw": dict(kind="adamw", lr=lrs[0], weight_decay=0.01, eps=ADAM_EPS)}[opt_kind]
    want = O.training_steps(H.score_cfg(fam, d, p), H.oracle_loss_cfg(LCFG), ocfg, ent, rel,
                            batches, "t", True, True, "mean", lr_schedule=lrs)
    sf, model = _model(H, fam, p, sh, n_rel, d, ent, rel)
    opt = (AdamW(lr=lrs[0], weight_decay=0.01, eps=ADAM_EPS) if opt_kind == "adamw"
           else SGD(lr=lrs[0], momentum=0.9 if opt_kind == "sgdm" else 0.0))
    step = training_model(model, opt)
    assert step.cuda_graph
    adam = opt_kind == "adamw"
    for s, b in enumerate(batches):
        opt.lr = lrs[s]
        res = step(**b)
        assert_close(res["loss"].cpu(), want["loss"][s], rtol=2e-5 if adam else 1e-5, atol=1e-4)
    torch.cuda.synchronize()
    assert len([g for g in step._graphs.values() if g != "warm"]) == 1  # replays did happen
    tol = dict(rtol=1e-4, atol=2e-5) if adam else dict(rtol=1e-5, atol=2e-6)
    assert_close(sf.entity_embedding.detach().cpu(), want["ent"], **tol)
    assert_close(sf.relation_embedding.detach().cpu(), want["rel"], **tol)


def test_buffer_growth_drops_captured_graphs():
    """small batch x3 (captured) -> larger batch (re-allocates workspaces and input buffers) ->
    a validation forward with an even larger batch on the same module -> small batch again.
    Every step must match the oracle; a stale replay would read freed / re-used buffers."""
    B, H = _imports()
    from besskge_b200.bess import training_model
    from besskge_b200.optim import SGD
    fam, p, d, n_rel = "TransE", 1, 32, 5
    sh, ent, rel, small, mk = _problem(fam, p, d=d, n_rel=n_rel, n_batch=5)
    big = [mk(24, 40) for _ in range(2)]
    seq = small[:3] + big[:1] + small[3:4] + big[1:] + small[4:]
    want = O.training_steps(H.score_cfg(fam, d, p), H.oracle_loss_cfg(LCFG), dict(kind="sgd", lr=0.1),
                            ent, rel, seq, "t", True, True, "mean")
    sf, model = _model(H, fam, p, sh, n_rel, d, ent, rel)
    step = training_model(model, SGD(lr=0.1))
    gens = []
    for s, b in enumerate(seq):
        if s == 4:  # validation forward on the same module, bigger than anything so far
            model(**mk(40, 64))
        res = step(**b)
        gens.append(step._graph_gen)
        assert_close(res["loss"].cpu(), want["loss"][s], rtol=1e-5, atol=1e-4,
                     msg=lambda m: f"step {s}: {m}")
    torch.cuda.synchronize()
    assert len(set(gens)) > 1  # the generation did move, i.e. graphs were invalidated
    assert_close(sf.entity_embedding.detach().cpu(), want["ent"], rtol=1e-5, atol=2e-6)
    assert_close(sf.relation_embedding.detach().cpu(), want["rel"], rtol=1e-5, atol=2e-6)


def test_optimizer_state_dict_round_trip():
    B, H = _imports()
    from besskge_b200.bess import training_model
    from besskge_b200.optim import AdamW
    fam, p, d, n_rel = "ComplEx", 2, 16, 4
    sh, ent, rel, batches, _ = _problem(fam, p, d=d, n_rel=n_rel, n_batch=4)
    ocfg = dict(kind="adamw", lr=0.01, weight_decay=0.01)
    want = O.training_steps(H.score_cfg(fam, d, p), H.oracle_loss_cfg(LCFG), ocfg, ent, rel, batches,
                            "t", True, True, "mean")
    sf, model = _model(H, fam, p, sh, n_rel, d, ent, rel)
    step = training_model(model, AdamW(lr=0.01, weight_decay=0.01))
    for b in batches[:2]:
        step(**b)
    state = step.state_dict()
    assert state["step"] == 2 and {"ent_s0_0", "ent_s1_1", "rel_s0", "rel_s1"} <= set(state)
    # resume in a fresh module / wrapper from the checkpointed tables + optimizer state
    sf2, model2 = _model(H, fam, p, sh, n_rel, d, sf.entity_embedding.detach().cpu(),
                         sf.relation_embedding.detach().cpu())
    step2 = training_model(model2, AdamW(lr=0.01, weight_decay=0.01))
    step2.load_state_dict(state)
    for b in batches[2:]:
        step2(**b)
    torch.cuda.synchronize()
    assert_close(sf2.entity_embedding.detach().cpu(), want["ent"], rtol=1e-4, atol=1e-5)
    assert_close(sf2.relation_embedding.detach().cpu(), want["rel"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("opt_kind,k,reduction", [("sgd", 2, "mean"), ("sgd", 3, "sum"),
                                                  ("adamw", 2, "mean"), ("sgdm", 3, "mean")])
def test_gradient_accumulation_vs_oracle(opt_kind, k, reduction):
    """k micro-batches from the same weights, one optimizer step (nb1 cell 26 / nb2 cell 14)."""
    B, H = _imports()
    from besskge_b200.bess import training_model
    from besskge_b200.optim import SGD, AdamW
    fam, p, d, n_rel = "RotatE", 1, 16, 5
    sh, ent, rel, batches, _ = _problem(fam, p, d=d, n_rel=n_rel, n_batch=6)
    ocfg = {"sgd": dict(kind="sgd", lr=0.1), "sgdm": dict(kind="sgd", lr=0.1, momentum=0.9),
            "adamw": dict(kind="adamw", lr=0.01, weight_decay=0.01, eps=ADAM_EPS)}[opt_kind]
    want = O.training_steps(H.score_cfg(fam, d, p), H.oracle_loss_cfg(LCFG), ocfg, ent, rel, batches,
                            "t", True, True, "mean", accumulate=k, accumulation_reduction=reduction)
    sf, model = _model(H, fam, p, sh, n_rel, d, ent, rel)
    opt = (AdamW(lr=0.01, weight_decay=0.01, eps=ADAM_EPS) if opt_kind == "adamw"
           else SGD(lr=0.1, momentum=0.9 if opt_kind == "sgdm" else 0.0))
    step = training_model(model, opt, gradient_accumulation=k, accumulation_reduction=reduction)
    adam = opt_kind == "adamw"
    for s, b in enumerate(batches):
        res = step(**b)
        assert_close(res["loss"].cpu(), want["loss"][s], rtol=2e-5 if adam else 1e-5, atol=1e-4)
    torch.cuda.synchronize()
    tol = dict(rtol=1e-4, atol=2e-5) if adam else dict(rtol=1e-5, atol=2e-6)
    assert_close(sf.entity_embedding.detach().cpu(), want["ent"], **tol)
    assert_close(sf.relation_embedding.detach().cpu(), want["rel"], **tol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_dense_optimizer_kernel_vs_torch_optim(dtype):
    """bess_opt_dense (vectorised) against torch.optim on a table with an odd row count, with
    sparse segment gradients; SGD-momentum and AdamW over 3 steps."""
    B, H = _imports()
    from besskge_b200 import _lib as L, kernels as K
    Es, W, G = 1001, 72, 257
    gen = torch.Generator().manual_seed(1)
    for kind, mk in ((L.OPT_SGDM, lambda ps: torch.optim.SGD(ps, lr=0.05, momentum=0.9,
                                                            weight_decay=0.01)),
                     (L.OPT_ADAMW, lambda ps: torch.optim.AdamW(ps, lr=0.01, weight_decay=0.02))):
        w0 = (torch.randn(Es, W, generator=gen)).to(dtype)
        ref = w0.float().clone().requires_grad_(True)
        opt = mk([ref])
        table = w0.clone().cuda()
        s0 = torch.zeros(Es, W, device="cuda")
        s1 = torch.zeros(Es, W, device="cuda")
        hyper = torch.zeros(L.HYPER_COUNT, device="cuda")
        for step_no in range(1, 4):
            rows = torch.randperm(Es, generator=gen)[:G]
            seg = torch.randn(G, W, generator=gen)
            r2s = torch.full((Es,), -1, dtype=torch.int32)
            r2s[rows] = torch.arange(G, dtype=torch.int32)
            dense = torch.zeros(Es, W)
            dense[rows] = seg
            ref.grad = dense
            opt.step()
            if kind == L.OPT_SGDM:
                K.set_hyper(hyper, 0.05, 0.9, 0.0, 0.0, 0.0, 0.0, 0.01, step_no)
            else:
                K.set_hyper(hyper, 0.01, 0.0, 0.0, 0.9, 0.999, 1e-8, 0.02, step_no)
            # by-value arguments deliberately wrong: the device array must win
            K.opt_dense(kind, table, seg.cuda(), r2s.cuda(), s0, s1, 123.0, 0.0, 0.0, 0.0, 0.0, 1.0,
                        0.0, 1, hyper)
            if dtype != torch.float32:  # the oracle's weights live in the table dtype too
                with torch.no_grad():
                    ref.copy_(ref.to(dtype).float())
        torch.cuda.synchronize()
        tol = dict(rtol=1e-5, atol=1e-6) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)
        assert_close(table.float().cpu(), ref.detach(), **tol)


def test_utils_device_helpers():
    """gather_indices / complex_multiplication / complex_rotation (utils.py:10-33, 72-112)."""
    B, H = _imports()
    from besskge_b200.utils import complex_multiplication, complex_rotation, gather_indices
    gen = torch.Generator().manual_seed(0)
    for a, b in ((7, 7), (1, 5), (6, 1)):
        for dt in (torch.float32, torch.int32, torch.bool, torch.float16, torch.int64):
            x = torch.randn(a, 19, generator=gen)
            x = (x > 0) if dt == torch.bool else (x * 100).to(dt)
            idx = torch.randint(19, (b, 4), generator=gen, dtype=torch.int32)
            got = gather_indices(x.cuda(), idx.cuda()).cpu()
            xs = x.expand(max(a, b), 19) if a == 1 else x
            ix = idx.expand(max(a, b), 4) if b == 1 else idx
            assert torch.equal(got, torch.gather(xs, 1, ix.long()))
    for dt, tol in ((torch.float32, 1e-6), (torch.bfloat16, 1e-2), (torch.float16, 2e-3)):
        v1 = torch.randn(33, 10, generator=gen).to(dt)
        v2 = torch.randn(33, 10, generator=gen).to(dt)
        r = torch.randn(33, 5, generator=gen).to(dt)
        e = 5
        f1, f2 = v1.float(), v2.float()
        want = torch.cat([f1[:, :e] * f2[:, :e] - f1[:, e:] * f2[:, e:],
                          f1[:, :e] * f2[:, e:] + f1[:, e:] * f2[:, :e]], -1)
        got = complex_multiplication(v1.cuda(), v2.cuda())
        assert got.dtype == dt
        assert_close(got.float().cpu(), want, rtol=tol, atol=tol)
        rc = torch.cat([torch.cos(r), torch.sin(r)], -1).float()
        want = torch.cat([f1[:, :e] * rc[:, :e] - f1[:, e:] * rc[:, e:],
                          f1[:, :e] * rc[:, e:] + f1[:, e:] * rc[:, :e]], -1)
        assert_close(complex_rotation(v1.cuda(), r.cuda()).float().cpu(), want, rtol=tol, atol=tol)
    with pytest.raises(B.BessLibraryError):
        gather_indices(torch.zeros(2, 2), torch.zeros(2, 1, dtype=torch.int32))
