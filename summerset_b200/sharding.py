"""Host-side partitioning of replica groups across the GPUs of one box.

Every (group, slot) instance is independent, so groups shard across ranks with no data-path
collective for the tally / Raft scan.  The RS encode path has one real exchange step: replica r of a
group whose leader ("home") is on rank h is simulated on rank (h + r) % world, so shard plane r of
each rank's local groups travels to that rank (rspaxos/request.rs:127-142 sends shard r to peer r),
and the simulated follower's ack bit-plane travels back (rspaxos/durability.rs:101-118).
Pure index arithmetic -- shared by bench.py (NCCL) and the gloo CPU tests.
"""
from __future__ import annotations

from typing import Dict, List, Tuple


def group_range(total_groups: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of [0, total_groups) -- rank gets [lo, hi)."""
    base, rem = divmod(total_groups, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def replica_rank(home_rank: int, replica: int, world: int) -> int:
    """Rank that simulates replica `replica` of the groups led from `home_rank`."""
    return (home_rank + replica) % world


def exchange_rounds(n_replicas: int, world: int, rank: int) -> List[Dict[str, List[int]]]:
    """Plans the shard exchange as ceil(n_replicas / world) all-to-all rounds in which every rank
    sends at most one shard plane to every rank.

    Round k, on `rank`:
      send[dst]  = replica id whose plane this rank sends to dst (or -1)
      recv[src]  = replica id of the plane arriving from src (or -1)
    In round k rank h sends replica r = k*world + o to rank (h + o) % world for o in [0, world).
    """
    rounds = []
    k = 0
    while k * world < n_replicas:
        send = [-1] * world
        recv = [-1] * world
        for o in range(world):
            r = k * world + o
            if r >= n_replicas:
                continue
            send[(rank + o) % world] = r
            recv[(rank - o) % world] = r
        rounds.append({"send": send, "recv": recv})
        k += 1
    return rounds
