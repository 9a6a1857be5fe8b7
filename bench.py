#!/usr/bin/env python
"""Benchmark of the BESS sharded KGE training step (BASELINE.json metric:
train triples/s incl. negative scoring; gather HBM GB/s vs peak).

    python bench.py --gpus N --steps K --warmup W          # B200 arm
    python bench.py --impl reference ...                   # reference CPU arm

A step = one micro-batch per GPU: gather -> exchange -> score -> loss ->
backward -> scatter + optimizer, `--shard-bs` positive triples per GPU, each
scored against `--negatives` shared negatives.  Workload (default) =
BASELINE.json configs[1]: ogbl-biokg-shaped synthetic KG, DistMult d=256, fp32,
one entity shard per GPU (n_shard = N), LogSigmoid loss, SGD.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (dataset shape, family, embedding_size, norm, dtype, default shard_bs, negatives/triple)
    "biokg-distmult-d256-fp32": ("ogbl-biokg", "DistMult", 256, 2, "fp32", 16384, 2048),
    "biokg-transe-l2-d128-fp32": ("ogbl-biokg", "TransE", 128, 2, "fp32", 8192, 64),
    "wikikg2-transe-l1-d256-bf16": ("ogbl-wikikg2", "TransE", 256, 1, "bf16", 8192, 256),
    # BASELINE.json configs[2]: top-10 tail prediction against ALL entities (inference);
    # metric = queries/s; shard_bs = queries per GPU per step; "negatives" unused
    "yago-complex-d256-topk": ("yago3-10", "ComplEx", 256, 2, "fp32", 2048, 0),
    # BASELINE.json configs[4]: score-moving inference against 500 triple-specific candidate
    # tails (TripleBasedShardedNegativeSampler, negatives scored where they are stored);
    # metric = queries/s + gather GB/s; shard_bs = queries per GPU per step
    "wikikg2-rotate-d512-scoremoving": ("ogbl-wikikg2", "RotatE", 512, 1, "fp32", 2048, 500),
    "wikikg2-pairre-d512-scoremoving": ("ogbl-wikikg2", "PairRE", 512, 1, "fp32", 2048, 500),
}
DTYPES = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}


def parse() -> argparse.Namespace:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="biokg-distmult-d256-fp32", choices=list(WORKLOADS))
    ap.add_argument("--shard-bs", type=int, default=0)
    ap.add_argument("--negatives", type=int, default=0)
    ap.add_argument("--n-triple", type=int, default=1 << 21,
                    help="synthetic training triples (sampled with replacement)")
    ap.add_argument("--ref-shard-bs", type=int, default=4096,
                    help="micro-batch of the reference CPU arm / cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks() -> dict:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops"], tensor_sustained=d["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu: int) -> None:
        self.gpu = gpu
        self.rows = []
        self.proc = None

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self) -> None:
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm))


# ---------------------------------------------------------------------------
def build_problem(args, n_shard: int):
    """Synthetic graph of the named shape + sharding + samplers (host side)."""
    from besskge_b200.batch_sampler import RandomShardedBatchSampler
    from besskge_b200.dataset import synthetic_kg
    from besskge_b200.negative_sampler import RandomShardedNegativeSampler
    from besskge_b200.sharding import PartitionedTripleSet, Sharding

    shape, fam, d, p, dt, sbs, nneg = WORKLOADS[args.workload]
    shard_bs = args.shard_bs or sbs
    negatives = args.negatives or nneg
    ds = synthetic_kg(shape, seed=1234, n_triple=args.n_triple)
    sh = Sharding.create(ds.n_entity, n_shard, seed=1234)
    pts = PartitionedTripleSet.create_from_dataset(ds, "train", sh)
    ns = RandomShardedNegativeSampler(max(negatives // n_shard, 1), sh, 1234, "t", False, True)
    bs = RandomShardedBatchSampler(pts, ns, shard_bs=shard_bs, batches_per_step=1, seed=1234)
    return dict(ds=ds, sh=sh, ns=ns, bs=bs, fam=fam, d=d, p=p, dtype=dt, shard_bs=shard_bs,
                negatives=max(negatives // n_shard, 1) * n_shard, shape=shape)


def make_model(prob, device, rank_tables_only=False):
    from besskge_b200 import scoring
    from besskge_b200.bess import EmbeddingMovingBessKGE
    from besskge_b200.loss import LogSigmoidLoss

    torch.manual_seed(1234)
    cls = getattr(scoring, prob["fam"])
    if prob["fam"] in ("DistMult", "ComplEx"):
        sf = cls(True, prob["sh"], prob["ds"].n_relation_type, prob["d"])
    else:
        sf = cls(True, prob["p"], prob["sh"], prob["ds"].n_relation_type, prob["d"])
    sf = sf.to(device=device, dtype=DTYPES[prob["dtype"]])
    model = EmbeddingMovingBessKGE(prob["ns"], sf, loss_fn=LogSigmoidLoss(12.0, True))
    return model, sf


def flat_batch(batch):
    return {k: v.flatten(end_dim=1) for k, v in batch.items()}


def time_kernel(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for _ in range(iters):
        fn()
    en.record()
    torch.cuda.synchronize()
    return st.elapsed_time(en) / iters * 1e-3  # seconds per launch


def kernel_rooflines(model, sf, prob, pk):
    """Stand-alone timing of the dominant negative-scoring kernel and of the gather
    (CUDA events on the launching stream); algorithmic flops/bytes per launch are
    those of SURVEY.md §8(d)."""
    from besskge_b200 import _lib as L, kernels as K

    dev = sf.entity_embedding.device
    S, N = prob["shard_bs"], prob["negatives"]
    ent = sf.entity_embedding.data[0]
    W = ent.shape[1]
    es = ent.element_size()
    cfg = sf.kernel_cfg()
    dt = L.dtype_code(ent.dtype)
    g = torch.Generator(device="cpu").manual_seed(0)
    idx = torch.randint(ent.shape[0], (2 * S + N,), generator=g, dtype=torch.int32).to(dev)
    out = torch.empty(2 * S + N, W, dtype=ent.dtype, device=dev)
    t_gather = time_kernel(lambda: K.gather_rows(ent, idx, out))
    rows = 2 * S + N
    gather_bytes = rows * (W * es + 4) + rows * W * es
    nvec = K.call("bess_query_nvec", L.C.byref(cfg))
    qv = torch.randn(S, nvec, W, device=dev)
    cand = out[2 * S:]
    scores = torch.empty(S, N, device=dev)
    if prob["fam"] in ("DistMult", "ComplEx"):
        # dominant kernel: the tcgen05 contraction (forward scores = Q C^T; the two backward
        # contractions have the same flop count)
        from besskge_b200.bess import _TcOperand
        ws = K.Workspace(dev)
        q_op = _TcOperand(ws, "bq", S, W, ent.dtype, False)
        q_op.fill(L.F32, L.rows(qv.view(S, W)), dt, None, dev)
        c_op = _TcOperand(ws, "bc", N, W, ent.dtype, False)
        c_op.fill(dt, L.rows(cand), dt, None, dev)
        gws = torch.empty(max(K.dot_gemm_workspace(S, N, W) // 4, 1), device=dev)
        t_score = time_kernel(lambda: K.dot_gemm(dt, q_op.hi, q_op.lo, q_op.ld, c_op.hi, c_op.lo, c_op.ld,
                                                 S, N, W, scores, L.IDENT, N, 0, False, gws))
        passes = 3 if ent.dtype == torch.float32 else 1
        work = 2.0 * S * N * W
        roof = dict(kernel="gemm_tc_kernel (tcgen05 shared-negative scores = Q C^T, "
                           + ("3xTF32: 3 tf32 MMAs per product = 6 bf16-equivalent passes"
                              if passes == 3 else "one kind::f16 MMA per product") + ")",
                    bound="tensor", achieved=work / t_score / 1e12, peak=pk["tensor"],
                    unit="TFLOP/s", traffic=None, mma_passes=passes,
                    tensor_pipe_tflops=work * passes * (2 if passes == 3 else 1) / t_score / 1e12)
    elif prob["p"] == 2:
        t_score = time_kernel(lambda: K.shared_fwd(cfg, dt, L.MODE_TAILS, qv, S, L.rows(cand), None, N,
                                                   scores, L.IDENT, N, 0, None))
        work = 2.0 * S * N * W
        roof = dict(kernel="pair_fwd_kernel (shared-negative L2 scoring, CUDA-core fp32 path)",
                    bound="tensor", achieved=work / t_score / 1e12, peak=pk["tensor"],
                    unit="TFLOP/s", traffic=None)
    else:
        t_score = time_kernel(lambda: K.shared_fwd(cfg, dt, L.MODE_TAILS, qv, S, L.rows(cand), None, N,
                                                   scores, L.IDENT, N, 0, None))
        work = (S * nvec * W * 4) + N * W * es + S * N * 4
        # the contract's roofline object (HBM bytes) + the roofline that actually binds this kernel:
        # S*N*W pair elements x 2 FADD on the FP32 pipe (148 SMs x 128 lanes x SM clock), the
        # algorithmic bytes being only ~17 MB per launch
        try:
            sm_ghz = torch.cuda.get_device_properties(dev).clock_rate * 1e-6
        except AttributeError:
            sm_ghz = 1.965
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        fp32_peak = n_sm * 128 * sm_ghz * 1e9
        fp32_ach = 2.0 * S * N * W / t_score
        roof = dict(kernel="pair_fwd_kernel (shared-negative L1 scoring)", bound="hbm",
                    achieved=work / t_score / 1e9, peak=pk["hbm"], unit="GB/s", traffic=None,
                    note="bytes are negligible here; the binding roofline is fp32_pipe",
                    fp32_pipe=dict(achieved=fp32_ach / 1e12, peak=fp32_peak / 1e12,
                                   unit="T lane-instr/s", frac=fp32_ach / fp32_peak,
                                   work="S*N*W pair elements x 2 FADD"))
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists() and prob["fam"] == "DistMult" and (S, N, W) == (16384, 2048, 256) and es == 4:
        roof["traffic"] = json.loads(tf.read_text()).get("gemm_tc_kernel<TF32X3> fwd S=16384 N=2048 W=256")
    if tf.exists() and prob["fam"] == "TransE" and prob["p"] == 1 and (S, N, W) == (8192, 256, 256) and es == 2:
        # (the score matrix written by this launch stays in L2: the capture shows 0 bytes written to DRAM)
        roof["traffic"] = json.loads(tf.read_text()).get("pair_fwd_kernel L1 bf16 S=8192 N=256 W=256")
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof["launch_us"] = t_score * 1e6
    roof["peak_source"] = pk["source"]
    gather = dict(kernel="gather_route_kernel", bound="hbm", achieved=gather_bytes / t_gather / 1e9,
                  peak=pk["hbm"], unit="GB/s", launch_us=t_gather * 1e6, rows=rows,
                  row_bytes=W * es)
    gather["frac"] = gather["achieved"] / gather["peak"]
    return roof, gather


def cpu_port_steps(prob, n_steps: int, shard_bs: int, threads: int):
    """The reference's plain-PyTorch CPU path (n_shard = 1 arithmetic of bess.py +
    scoring.py + loss.py, torch autograd, dense torch.optim.SGD) restated by the
    oracle; returns seconds per step."""
    from besskge_b200.batch_sampler import RandomShardedBatchSampler
    from besskge_b200.negative_sampler import RandomShardedNegativeSampler
    from besskge_b200.sharding import PartitionedTripleSet, Sharding
    from oracle import besskge_oracle as O

    torch.set_num_threads(threads)
    ds = prob["ds"]
    sh = Sharding.create(ds.n_entity, 1, seed=1234)
    pts = PartitionedTripleSet.create_from_dataset(ds, "train", sh)
    ns = RandomShardedNegativeSampler(prob["negatives"], sh, 1234, "t", False, True)
    bs = RandomShardedBatchSampler(pts, ns, shard_bs=shard_bs, batches_per_step=1, seed=1234)
    fam, d = prob["fam"], prob["d"]
    W = d * (2 if fam in ("RotatE", "ComplEx", "BoxE") else 1)
    Wr = {"ComplEx": 2 * d, "PairRE": 2 * d, "BoxE": 4 * d + 2}.get(fam, d)
    torch.manual_seed(1234)
    ent = (torch.rand(1, sh.max_entity_per_shard, W) * 2 - 1).div_(W).requires_grad_(True)
    rel = (torch.rand(ds.n_relation_type, Wr) * 2 - 1).div_(Wr).requires_grad_(True)
    opt = torch.optim.SGD([ent, rel], lr=1e-3)
    cfg = dict(family=fam, d=d, norm_p=prob["p"])
    lcfg = dict(kind="logsigmoid", margin=12.0, adversarial=True, adv_scale=1.0)
    w = torch.tensor([1.0])
    times = []
    for i in range(n_steps):
        b = {k: v[0] for k, v in bs[[i]].items()}
        t0 = time.perf_counter()
        opt.zero_grad()
        pos, neg = O.embedding_moving_forward(cfg, ent, rel, b["head"], b["relation"], b["tail"],
                                              b["negative"], "t", True, True)
        loss = O.loss_value(lcfg, pos[0].float(), neg[0].float(), w)
        loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    prob = build_problem(args, 1)
    threads = os.cpu_count() or 1
    sbs = min(args.ref_shard_bs, prob["shard_bs"])
    times = cpu_port_steps(prob, args.warmup + args.steps, sbs, threads)
    t = float(np.mean(times[args.warmup:]))
    value = sbs / t
    line = {
        "impl": "reference", "metric": "train_triples_per_sec", "value": value, "unit": "triples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "dataset_shape": prob["shape"],
                   "n_entity": prob["ds"].n_entity, "n_shard": 1, "shard_bs": sbs,
                   "negatives_per_triple": prob["negatives"], "loss": "LogSigmoid(12, adversarial)",
                   "optimizer": "SGD(1e-3)", "score_fn": prob["fam"], "embedding_size": prob["d"],
                   "sample": f"bounded sample of the workload: micro-batches of {sbs} triples "
                             f"(the B200 arm uses {prob['shard_bs']}) on the host cores"},
        "cpu_baseline": {"value": value, "unit": "triples/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of shard_bs={sbs} (oracle port of the reference's "
                                   "plain-PyTorch n_shard=1 path: forward + autograd + dense SGD)"},
        "e2e": {"value": value, "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else any library prints to
    fd 1 during the run (e.g. NCCL's version banner) was redirected to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def run_topk(args) -> None:
    """configs[2]: TopKQueryBessKGE, ComplEx d=256 (row 512) fp32, k=10, every query scored
    against all entities of a YAGO3-10-shaped graph; one step = shard_bs queries per GPU."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    import besskge_b200  # noqa: F401
    from besskge_b200 import _lib as L_, scoring
    from besskge_b200.bess import TopKQueryBessKGE
    from besskge_b200.dataset import DATASET_SHAPES
    from besskge_b200.metric import Evaluation
    from besskge_b200.negative_sampler import PlaceholderNegativeSampler
    from besskge_b200.sharding import Sharding

    pk = peaks()
    shape, fam, d, _, dt, sbs, _ = WORKLOADS[args.workload]
    S = args.shard_bs or sbs
    n_entity, n_rel, _ = DATASET_SHAPES[shape]
    n = world
    sh = Sharding.create(n_entity, n, seed=1234)
    torch.manual_seed(1234)
    sf = scoring.ComplEx(True, sh, n_rel, d).to(device=dev, dtype=DTYPES[dt])
    ev = Evaluation(["mrr", "hits@10"], worst_rank_infty=True, reduction="sum")
    model = TopKQueryBessKGE(k=10, candidate_sampler=PlaceholderNegativeSampler("t"), score_fn=sf,
                             evaluation=ev, return_scores=True, window_size=500)
    g = torch.Generator().manual_seed(1234 + rank)
    total = args.warmup + args.steps
    lo = int(sh.shard_counts.min())
    batches = []
    for _ in range(total):  # [n_shard, S] host layout; distributed ranks read their own row
        batches.append(dict(
            relation=torch.randint(n_rel, (n, S), generator=g, dtype=torch.int32).pin_memory(),
            head=torch.randint(lo, (n, S), generator=g, dtype=torch.int32).pin_memory(),
            tail=torch.randint(n_entity, (n, S), generator=g, dtype=torch.int32).pin_memory()))

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(read_back: bool):
        for i in range(args.warmup):
            model(**batches[i])
        barrier()
        c0 = L_.call("bess_launch_count")
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        d2h = 0
        for i in range(args.warmup, total):
            out = model(**batches[i])
            if read_back:
                ids = out["topk_global_id"].cpu()
                d2h = ids.numel() * ids.element_size()
        en.record()
        barrier()
        t = torch.tensor([st.elapsed_time(en) * 1e-3], device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item()), d2h, L_.call("bess_launch_count") - c0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_dev, _, launches = timed(False)
    clocks = sampler.stop() if rank == 0 else None
    t_e2e, d2h, _ = timed(True)
    if rank == 0:
        Es, W = sh.max_entity_per_shard, 2 * d
        local = 1 if world > 1 else n
        # per GPU and step: all n*S queries against the local shard(s)
        flops = 2.0 * (n * S) * Es * W * local
        passes = 3 if dt == "fp32" else 1
        queries = n * S * args.steps
        line = {
            "metric": "topk_queries_per_sec", "value": queries / t_dev, "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "f32"}.get(dt, dt), "data": "synthetic",
            "config": {"workload": args.workload, "dataset_shape": shape, "n_entity": n_entity,
                       "n_shard": n, "queries_per_gpu_per_step": S, "k": 10, "score_fn": fam,
                       "embedding_size": d, "candidates": "all entities",
                       "l2": "no flush: every step streams the whole shard "
                             f"({Es * W * 4 / 1e6:.0f} MB of fp32 rows + operand copies) > 126 MB L2"
                             if Es * W * 12 > 126e6 else "shard operands are L2-resident "
                             f"({Es * W * 12 / 1e6:.0f} MB incl. hi/lo copies); score windows "
                             f"({n * S * 4096 * 4 / 1e6:.0f} MB each) are not"},
            "e2e": {"value": queries / t_e2e, "unit": "queries/s",
                    "h2d_bytes_per_step": 3 * S * 4 * local, "d2h_bytes_per_step": d2h,
                    "ms_per_step": t_e2e / args.steps * 1e3},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"kernel": "gemm_tc_kernel (tcgen05 window scores = Q C^T) over the whole step",
                         "bound": "tensor", "achieved": flops * args.steps / t_dev / 1e12,
                         "peak": pk["tensor_sustained"], "unit": "TFLOP/s", "traffic": None,
                         "mma_passes": passes,
                         "frac": flops * args.steps / t_dev / 1e12 / pk["tensor_sustained"],
                         "note": "whole-step figure (GEMM + operand split + top-k merge); scores/s = "
                                 f"{(n * S) * Es * local * args.steps / t_dev:.3e}",
                         "peak_source": pk["source"]},
            "cpu_baseline": None,
        }
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


def run_scoremoving(args) -> None:
    """configs[4]: ScoreMovingBessKGE, RotatE / PairRE d=512 fp32 (entity rows 1024 / 512 wide),
    500 candidate tails per query split by owning shard (TripleBasedShardedNegativeSampler),
    ranks -> MRR / Hits@10.  The dominant kernel is the fused gather + score stream over the
    candidate rows (pertriple_fwd): Q * Nn rows of W * 4 bytes read in place from the shard."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    import besskge_b200  # noqa: F401
    from besskge_b200 import _lib as L_, kernels as K, scoring
    from besskge_b200.batch_sampler import RigidShardedBatchSampler
    from besskge_b200.bess import ScoreMovingBessKGE
    from besskge_b200.dataset import synthetic_kg
    from besskge_b200.metric import Evaluation
    from besskge_b200.negative_sampler import TripleBasedShardedNegativeSampler
    from besskge_b200.sharding import PartitionedTripleSet, Sharding

    pk = peaks()
    shape, fam, d, p, dt, sbs, n_cand = WORKLOADS[args.workload]
    S = args.shard_bs or sbs
    n = world
    total = args.warmup + args.steps
    n_query = n * S * total
    ds = synthetic_kg(shape, seed=1234, n_triple=n_query)
    ds.neg_tails = {"train": np.random.default_rng(4321).integers(
        ds.n_entity, size=(n_query, args.negatives or n_cand), dtype=np.int32)}
    sh = Sharding.create(ds.n_entity, n, seed=1234)
    pts = PartitionedTripleSet.create_from_dataset(ds, "train", sh)
    ns = TripleBasedShardedNegativeSampler(pts.neg_heads, pts.neg_tails, sh, "t", 1234)
    bs = RigidShardedBatchSampler(pts, ns, shard_bs=S, batches_per_step=1, seed=1234)
    torch.manual_seed(1234)
    sf = getattr(scoring, fam)(False, p, sh, ds.n_relation_type, d)
    sf = sf.to(device=dev, dtype=DTYPES[dt])
    ev = Evaluation(["mrr", "hits@10"], reduction="sum")
    model = ScoreMovingBessKGE(ns, sf, evaluation=ev)
    size = bs.partition_sample_size
    span = len(bs)
    batches = []
    for i in range(total):
        idx = [(i * size + j) % span for j in range(size)]
        batches.append({k: v.flatten(end_dim=1).pin_memory() for k, v in bs[idx].items()})
    Nn = int(ns.padded_shard_length)
    W = sf.entity_embedding.shape[-1]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(read_back: bool):
        for i in range(args.warmup):
            model(**batches[i])
        barrier()
        c0 = L_.call("bess_launch_count")
        st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        st.record()
        d2h = 0
        for i in range(args.warmup, total):
            out = model(**batches[i])
            if read_back:
                m = out["metrics"].cpu()
                d2h = m.numel() * m.element_size()
        en.record()
        barrier()
        t = torch.tensor([st.elapsed_time(en) * 1e-3], device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item()), d2h, L_.call("bess_launch_count") - c0

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_dev, _, launches = timed(False)
    clocks = sampler.stop() if rank == 0 else None
    t_e2e, d2h, _ = timed(True)
    if rank == 0:
        # dominant kernel alone: the fused gather + score stream of one step on this GPU
        cfg = sf.kernel_cfg()
        dtc = L_.dtype_code(sf.entity_embedding.dtype)
        nvec = K.call("bess_query_nvec", L_.C.byref(cfg))
        local = 1 if world > 1 else n
        Q = n * S  # every query of every shard is scored against this shard's candidates
        table = sf.entity_embedding.data[0]
        g = torch.Generator().manual_seed(0)
        idx = torch.randint(int(sh.shard_counts.min()), (Q * Nn,), generator=g, dtype=torch.int32).to(dev)
        qv = torch.randn(Q, nvec, W, device=dev)
        sc = torch.empty(Q, Nn, device=dev)
        t_k = time_kernel(lambda: K.pertriple_fwd(cfg, dtc, L_.MODE_TAILS, qv, Q, L_.rows(table, idx=idx),
                                                  Nn, Nn, sc, L_.IDENT, Nn, 0, None))
        es = table.element_size()
        kbytes = Q * Nn * (W * es + 4) + Q * Nn * 4 + Q * nvec * W * 4
        h2d = sum(v.numel() * v.element_size() for v in batches[0].values()) // (n if world > 1 else 1)
        queries = n * S * args.steps
        line = {
            "metric": "scoremoving_queries_per_sec", "value": queries / t_dev, "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "f32"}.get(dt, dt), "data": "synthetic",
            "config": {"workload": args.workload, "dataset_shape": shape, "n_entity": ds.n_entity,
                       "n_shard": n, "queries_per_gpu_per_step": S, "candidates_per_query": n_cand,
                       "padded_candidates_per_shard": Nn, "score_fn": fam, "embedding_size": d,
                       "entity_row_bytes": W * es, "negative_sample_sharing": False,
                       "l2": f"no flush: one step streams {kbytes * local / 1e6:.0f} MB of randomly "
                             "placed candidate rows per GPU (> 126 MB L2)"},
            "e2e": {"value": queries / t_e2e, "unit": "queries/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e / args.steps * 1e3},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"kernel": "pertriple_fwd_row_kernel (fused gather + score of per-query candidates)",
                         "bound": "hbm", "achieved": kbytes / t_k / 1e9, "peak": pk["hbm"],
                         "unit": "GB/s", "traffic": None, "frac": kbytes / t_k / 1e9 / pk["hbm"],
                         "launch_us": t_k * 1e6, "rows_per_launch": Q * Nn,
                         "step_share": t_k * local / (t_dev / args.steps),
                         "peak_source": pk["source"]},
            "cpu_baseline": None,
        }
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


def main() -> None:
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.workload.endswith("-scoremoving"):
        if args.impl == "reference":
            emit({"impl": "reference", "unavailable": "the reference arm is defined on the training "
                                                      "workloads (configs[1]); run without --workload"})
            return
        run_scoremoving(args)
        return
    if args.workload.endswith("-topk"):
        if args.impl == "reference":
            emit({"impl": "reference", "unavailable": "the reference arm is defined on the training "
                                                      "workloads (configs[1]); run without --workload"})
            return
        run_topk(args)
        return
    if args.impl == "reference":
        run_reference(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    import besskge_b200  # noqa: F401  (fails loudly without the CUDA library)
    from besskge_b200.bess import training_model
    from besskge_b200.optim import SGD

    pk = peaks()
    n_shard = world
    prob = build_problem(args, n_shard)
    model, sf = make_model(prob, dev)
    step = training_model(model, SGD(lr=1e-3))
    S, N = prob["shard_bs"], prob["negatives"]

    total = args.warmup + args.steps
    # every rank draws the same batches (same seeds); distinct batch per step
    host_batches = []
    for i in range(total):
        b = flat_batch(prob["bs"][[i]])
        host_batches.append({k: v.pin_memory() for k, v in b.items()})

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # -------- kernel-only leg: inputs resident in HBM ------------------------
    from besskge_b200 import _lib as L_
    staged = [step.stage(**b) for b in host_batches]
    torch.cuda.synchronize()
    launches_per_step = 0
    for i in range(args.warmup):
        c0 = L_.call("bess_launch_count")
        step.run_staged(staged[i])
        if i == 0:  # the first call runs eagerly: every kernel launch of one step is counted
            launches_per_step = L_.call("bess_launch_count") - c0
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st.record()
    for i in range(args.warmup, total):
        out = step.run_staged(staged[i])
    en.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t_dev = torch.tensor([st.elapsed_time(en) * 1e-3], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_dev, op=torch.distributed.ReduceOp.MAX)
    t_dev = float(t_dev.item())
    loss_last = float(out["loss"].sum().item())

    # -------- end-to-end leg: host batch in, loss out, every step -------------
    for i in range(min(2, args.warmup)):
        step(**host_batches[i])["loss"].cpu()
    barrier()
    st.record()
    d2h = 0
    for i in range(args.warmup, total):
        l = step(**host_batches[i])["loss"].cpu()
        d2h = l.numel() * l.element_size()
    en.record()
    barrier()
    t_e2e = torch.tensor([st.elapsed_time(en) * 1e-3], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_e2e, op=torch.distributed.ReduceOp.MAX)
    t_e2e = float(t_e2e.item())
    h2d = staged[0].h2d_bytes

    if rank == 0:
        roof, gather = kernel_rooflines(model, sf, prob, pk)
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sbs = min(args.ref_shard_bs, S)
            times = cpu_port_steps(prob, 5, sbs, threads)
            tc = float(np.mean(times[2:]))
            cpu = {"value": sbs / tc, "unit": "triples/s", "cores": threads, "kind": "port",
                   "sample": f"3 timed steps (2 warm-up) of shard_bs={sbs}, {N} shared negatives, "
                             "n_shard=1: oracle port of the reference's plain-PyTorch path"}
        triples = world * S * args.steps
        ws_bytes = model._ws.bytes()
        line = {
            "metric": "train_triples_per_sec", "value": triples / t_dev, "unit": "triples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "f32"}.get(prob["dtype"], prob["dtype"]),
            "data": "synthetic",
            "config": {"workload": args.workload, "dataset_shape": prob["shape"],
                       "n_entity": prob["ds"].n_entity, "n_shard": n_shard, "shard_bs": S,
                       "negatives_per_triple": N, "loss": "LogSigmoid(12, adversarial)",
                       "optimizer": "SGD(1e-3)", "score_fn": prob["fam"], "embedding_size": prob["d"],
                       "l2": f"no flush: per-step working set {ws_bytes / 1e6:.0f} MB of buffers "
                             f"+ table > 126 MB L2",
                       "final_loss": loss_last},
            "e2e": {"value": triples / t_e2e, "unit": "triples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": t_e2e / args.steps * 1e3},
            "gpu_launches": None,
            "clocks": clocks,
            "roofline": roof,
            "gather": gather,
            "cpu_baseline": cpu,
        }
        # kernels of this library per step (counted by the library on the eager first step;
        # later steps replay the same launches from the captured CUDA graph) x timed steps
        line["gpu_launches"] = int(launches_per_step) * args.steps
        line["gpu_launches_per_step"] = int(launches_per_step)
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
