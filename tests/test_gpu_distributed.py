"""-m gpu: the multi-rank path, collected by pytest.

Spawns `tests/dist_gpu_worker.py` under torchrun with 2 ranks: distributed EmbeddingMoving
training through the peer-memory exchange (csrc/peer.cu: remote gather into the peers'
receive buffers, signal / wait handshakes, gradient push — early, from the side stream, for
the tensor-core 't' path —, relation partial push + reduce; the whole step replayed from one
CUDA graph per rank) against the CPU oracle.  With >= 2 GPUs: one rank per GPU over NVLink and
NCCL, and the distributed inference modules as well.  On a 1-GPU box: both ranks share cuda:0
(symmetric memory across two processes of one device, rendezvous over gloo) — same kernels,
same protocol, no NVLink.
Reference behaviour matched: /root/reference/tests/test_bess.py:54-275 (4 replicas on the IPU
model) — here against the oracle that is pinned to fixtures of that code."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_worker(n_ranks: int, same_device: bool, timeout: int = 420):
    env = dict(os.environ)
    env["BESS_TEST_SAME_DEVICE"] = "1" if same_device else "0"
    env["BESS_PEER_TIMEOUT_S"] = "60"  # a protocol bug must end in a CUDA error, not a hang
    env.setdefault("OMP_NUM_THREADS", "2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={n_ranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "dist_gpu_worker.py")]
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_two_rank_peer_exchange_training_vs_oracle():
    """>= 2 GPUs: one rank per GPU.  On a 1-GPU box the two ranks would have to spin on each
    other's flags as separate processes time-sliced on ONE GPU, which the B200 driver does not
    guarantee to co-schedule (context-switch timeouts have been seen): that mode only runs when
    BESS_ALLOW_SAME_DEVICE_RANKS=1 asks for it; the protocol itself is covered on one GPU by
    test_peer_protocol_emulated_ranks_one_process below."""
    n_gpu = torch.cuda.device_count()
    same_device = n_gpu < 2
    if same_device and os.environ.get("BESS_ALLOW_SAME_DEVICE_RANKS", "0") != "1":
        pytest.skip("needs 2 GPUs (set BESS_ALLOW_SAME_DEVICE_RANKS=1 to time-slice 2 ranks on one)")
    torch.cuda.synchronize()
    res = _run_worker(2, same_device)
    tail = (res.stdout[-3000:] + "\n--- stderr ---\n" + res.stderr[-3000:])
    assert res.returncode == 0, tail
    for needle in ("distributed parity ok: TransE t", "distributed parity ok: DistMult t",
                   "distributed parity ok: DistMult ht", "distributed parity ok: RotatE h"):
        assert needle in res.stdout, tail
    if not same_device:
        assert "distributed TopKQuery == local" in res.stdout, tail
        assert "distributed ScoreMoving == local" in res.stdout, tail


def test_two_rank_peer_exchange_ipc_transport():
    """Same worker over this library's own CUDA-IPC mapping instead of torch symmetric memory."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    os.environ["BESS_PEER_TRANSPORT"] = "ipc"
    try:
        res = _run_worker(2, False)
    finally:
        del os.environ["BESS_PEER_TRANSPORT"]
    tail = (res.stdout[-3000:] + "\n--- stderr ---\n" + res.stderr[-3000:])
    assert res.returncode == 0, tail
    assert "distributed parity ok: DistMult t" in res.stdout, tail


@pytest.mark.parametrize("n", [2, 4])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_peer_protocol_emulated_ranks_one_process(n, dtype):
    """The peer-exchange kernels (csrc/peer.cu + the routed gather of csrc/rows.cu) through the
    C-ABI with n emulated ranks in ONE process on ONE stream: every rank's receive buffer, flag
    rows and counters are ordinary allocations, 'peer pointers' point at the other ranks'
    buffers, and the launches are ordered so that every wait follows the signals it needs —
    nothing spins.  Checks, for three consecutive steps (sequence numbers, buffer reuse):
    the routed gather puts row (src rank r -> dst j, slot r) exactly where the AllToAll of
    bess.py:348-350 would; the gradient push is its transpose; the relation partials reduce to
    the same bits on every rank, in rank order."""
    from besskge_b200 import kernels as K
    g = torch.Generator().manual_seed(n)
    Es, W, S, per, cnt = 257, 64, 24, 40, 1024
    dev = "cuda"
    tables = [torch.randn(Es, W, generator=g).to(dtype).to(dev) for _ in range(n)]
    tn = [torch.zeros(n, per, W, dtype=dtype, device=dev) for _ in range(n)]       # receive buffers
    gback = [torch.zeros(n, per, W, device=dev) for _ in range(n)]                 # gradient receive
    rel_slots = [torch.zeros(n, cnt, device=dev) for _ in range(n)]
    flags = [torch.zeros(3, 64, dtype=torch.int32, device=dev) for _ in range(n)]
    counters = [torch.zeros(3, dtype=torch.int32, device=dev) for _ in range(n)]
    heads = [torch.empty(S, W, dtype=dtype, device=dev) for _ in range(n)]
    for step in range(3):
        idx = [torch.randint(Es, (S + n * per,), generator=g, dtype=torch.int32).to(dev) for _ in range(n)]
        # ---- forward: every rank gathers and stores straight into every rank's receive buffer
        for r in range(n):
            K.gather_route(tables[r], idx[r], S, per, heads[r], [t.data_ptr() for t in tn], r)
            K.peer_signal(counters[r][0:1], [f[0].data_ptr() for f in flags], r)
        for r in range(n):
            K.peer_wait(counters[r][0:1], flags[r][0], n, timeout_ms=2000)
        torch.cuda.synchronize()
        for r in range(n):
            assert torch.equal(heads[r], tables[r][idx[r][:S].long()])
            for j in range(n):  # what rank j received from rank r
                want = tables[r][idx[r][S + j * per:S + (j + 1) * per].long()]
                assert torch.equal(tn[j][r], want)
            assert flags[r][0][:n].tolist() == [step + 1] * n
        # ---- backward: gradient blocks go back to their owners (block j of rank r -> slot r of j)
        dtn = [torch.randn(n, per, W, generator=g).to(dev) for _ in range(n)]
        blk = per * W * 4
        for r in range(n):
            K.peer_push(dtn[r], blk, [b.data_ptr() + r * blk for b in gback], blk)
            K.peer_signal(counters[r][1:2], [f[1].data_ptr() for f in flags], r)
        for r in range(n):
            K.peer_wait(counters[r][1:2], flags[r][1], n, timeout_ms=2000)
        # ---- relation partials: slot [rank] of every rank, reduced in rank order
        part = [torch.randn(cnt, generator=g).to(dev) for _ in range(n)]
        outs = [torch.empty(cnt, device=dev) for _ in range(n)]
        for r in range(n):
            K.peer_push(part[r], 0, [sl.data_ptr() + r * cnt * 4 for sl in rel_slots], cnt * 4)
            K.peer_signal(counters[r][2:3], [f[2].data_ptr() for f in flags], r)
        for r in range(n):
            K.peer_wait(counters[r][2:3], flags[r][2], n, timeout_ms=2000)
            K.peer_reduce(rel_slots[r], n, cnt, 1.0 / n, outs[r])
        torch.cuda.synchronize()
        for j in range(n):
            for r in range(n):
                assert torch.equal(gback[j][r], dtn[r][j])
        acc = torch.zeros(cnt, device=dev)
        for r in range(n):
            acc = acc + part[r]
        for r in range(n):
            assert torch.equal(outs[r], outs[0])
            assert torch.equal(outs[r], acc * (1.0 / n))
