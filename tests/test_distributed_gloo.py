"""world_size-2 gloo test (CPU) of the distributed-mode routing logic."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_two_rank_routing():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           str(ROOT / "tests" / "dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "rank 0 ok" in res.stdout and "rank 1 ok" in res.stdout
