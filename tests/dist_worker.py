"""Worker of tests/test_distributed_gloo.py (world_size 2, gloo, CPU tensors):
checks the block routing of distributed BESS — what each rank puts in its send
buffer, what all_to_all delivers, the reverse (gradient) exchange and which
staged rows a rank consumes — against the oracle's routing rule."""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))

from besskge_b200.bess import _Placement  # noqa: E402


def main() -> None:
    dist.init_process_group("gloo")
    rank, n = dist.get_rank(), dist.get_world_size()
    pl = _Placement(n)
    assert pl.distributed and pl.shards == [rank] and pl.n_local == 1
    g = torch.Generator().manual_seed(0)
    Es, W, p, B, Nn = 50, 8, 3, 2, 4
    per = p + B * Nn
    ent = torch.randn(n, Es, W, generator=g)
    tail = torch.randint(Es, (n, n, p), generator=g)  # [shard_t, shard_h, p]
    neg = torch.randint(Es, (n, n, B, Nn), generator=g)  # [src, dst, B, Nn]
    # what this rank's gather writes into SEND[j] (bess_gather_route, slot 0)
    send = torch.stack([ent[rank][torch.cat([tail[rank, j], neg[rank, j].flatten()])]
                        for j in range(n)])
    recv = torch.empty(n, per, W)
    pl.all_to_all(recv, send)
    # routing rule: replica r receives from shard j the tails of block (r, j)
    # and the negatives shard j drew for r
    for j in range(n):
        want = ent[j][torch.cat([tail[j, rank], neg[j, rank].flatten()])]
        assert torch.equal(recv[j], want), f"rank {rank}: block from shard {j} misrouted"
    # reverse exchange of gradients: rank r's d_recv[j] must come back to shard j slot r
    d_recv = torch.stack([torch.full((per, W), float(10 * rank + j)) for j in range(n)])
    d_back = torch.empty(n, per, W)
    pl.all_to_all(d_back, d_recv)
    for r2 in range(n):
        assert torch.all(d_back[r2] == float(10 * r2 + rank))
    # relation gradient all-reduce (sum) then mean
    gr = torch.full((3, 4), float(rank + 1))
    dist.all_reduce(gr)
    assert torch.all(gr == sum(range(1, n + 1)))
    # staged rows: rank takes rows rank::n of the (bps*n)-row host layout
    bps = 3
    host = torch.arange(bps * n)
    mine = host[pl.rank::n]
    assert mine.tolist() == [s * n + rank for s in range(bps)]
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
