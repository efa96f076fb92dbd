"""world_size-2 gloo worker (CPU): exercises the host-side sharding / exchange plan used by bench.py at N > 1.

Each rank owns a contiguous range of groups, builds R shard planes whose bytes encode (home rank, replica,
group), runs the exchange rounds of summerset_b200.sharding (send plane r to rank (home + r) % world), then
returns an ack plane per received shard and checks that every plane landed where the plan says.
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from summerset_b200 import sharding  # noqa: E402


def xchg(outs, ins):
    """list all-to-all over point-to-point ops (gloo has no list all_to_all)"""
    world, rank = dist.get_world_size(), dist.get_rank()
    ops = []
    for peer in range(world):
        if peer == rank:
            if ins[peer].numel():
                outs[peer].copy_(ins[peer])
            continue
        if ins[peer].numel():
            ops.append(dist.P2POp(dist.isend, ins[peer], peer))
        if outs[peer].numel():
            ops.append(dist.P2POp(dist.irecv, outs[peer], peer))
    for w in (dist.batch_isend_irecv(ops) if ops else []):
        w.wait()


def main():
    dist.init_process_group("gloo")
    world, rank = dist.get_world_size(), dist.get_rank()
    R, total, ds = 5, 1001, 16
    lo, hi = sharding.group_range(total, rank, world)
    # ranges tile [0, total)
    spans = [sharding.group_range(total, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == total and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    n = hi - lo
    nmax = max(b - a for a, b in spans)
    g = torch.arange(lo, hi, dtype=torch.int64)
    shards = torch.zeros((R, nmax, ds), dtype=torch.int64)
    for r in range(R):
        shards[r, :n] = (rank * 1000003 + r * 10007 + g)[:, None]
    recv = torch.full((R, nmax, ds), -1, dtype=torch.int64)
    rounds = sharding.exchange_rounds(R, world, rank)
    assert len(rounds) == -(-R // world)
    for rd in rounds:
        ins = [shards[rd["send"][d]] if rd["send"][d] >= 0 else shards[0][:0] for d in range(world)]
        outs = [recv[rd["recv"][s]] if rd["recv"][s] >= 0 else recv[0][:0] for s in range(world)]
        xchg(outs, ins)
    # replica r held here came from home rank (rank - r) % world
    for r in range(R):
        home = (rank - r) % world
        assert sharding.replica_rank(home, r, world) == rank
        hlo, hhi = spans[home]
        want = home * 1000003 + r * 10007 + torch.arange(hlo, hhi, dtype=torch.int64)
        assert torch.equal(recv[r, :hhi - hlo, 0], want), (rank, r)
    # acks travel back: follower of (home, r) returns a plane to home
    acks = torch.zeros((R, nmax), dtype=torch.int64)
    for r in range(R):
        acks[r] = recv[r, :, 0] * 2 + 1
    ack_recv = torch.full((R, nmax), -1, dtype=torch.int64)
    for rd in rounds:
        ins = [acks[rd["recv"][s]] if rd["recv"][s] >= 0 else acks[0][:0] for s in range(world)]
        outs = [ack_recv[rd["send"][d]] if rd["send"][d] >= 0 else ack_recv[0][:0] for d in range(world)]
        xchg(outs, ins)
    for r in range(R):
        assert torch.equal(ack_recv[r, :n], shards[r, :n, 0] * 2 + 1), (rank, r)
    dist.barrier()
    if rank == 0:
        print("SHARD_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
