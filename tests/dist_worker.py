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
    # score-moving / all-scores exchange (bess.py:519-592, :1023-1060): queries are replicated
    # with all_gather, every rank scores ALL n*S queries against ITS candidates, and
    # all_to_all_single of the [n*S, X] local scores gives the owner of query (r, q) the block
    # out[q, j, x] = score(query (r, q), candidate x of shard j)
    S, X = 5, 3
    q_mine = torch.arange(S, dtype=torch.float32) + 100.0 * rank  # this rank's queries
    q_all = torch.empty(n * S)
    dist.all_gather_into_tensor(q_all, q_mine)
    assert torch.equal(q_all.view(n, S), torch.stack(
        [torch.arange(S, dtype=torch.float32) + 100.0 * j for j in range(n)]))
    cand_mine = torch.arange(X, dtype=torch.float32) * 0.25 + 10.0 * rank
    sc_local = q_all.view(-1, 1) * 1000.0 + cand_mine.view(1, -1)  # [n*S, X]
    sc_recv = torch.empty(n, S, X)
    dist.all_to_all_single(sc_recv.view(-1), sc_local.contiguous().view(-1))
    out = sc_recv.transpose(0, 1)  # [S, n, X]
    for j in range(n):
        want = q_mine.view(-1, 1) * 1000.0 + (torch.arange(X, dtype=torch.float32) * 0.25 + 10.0 * j)
        assert torch.equal(out[:, j], want), f"rank {rank}: score block of shard {j} misrouted"
    # AllScoresPipeline row selection in distributed mode: this rank keeps the valid rows of ITS
    # shard, in (step, triple) order
    tmask = (torch.arange(bps * n * 4).view(bps, n, 4) % 3) != 0
    sel = tmask.clone()
    own = torch.zeros_like(sel)
    own[:, rank] = True
    sel &= own
    row_src = torch.nonzero(tmask[:, rank].reshape(-1)).reshape(-1)
    flat_ids = torch.arange(bps * n * 4).view(bps, n, 4)
    assert torch.equal(flat_ids[sel], flat_ids[:, rank].reshape(-1)[row_src])
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
