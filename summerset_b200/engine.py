"""Batched multi-group leader engine: the composition of the path's kernels that a host drives once per tick.

One object holds the leader-side state of G independent RSPaxos (or MultiPaxos) groups as the struct-of-arrays of
DESIGN.md section 2 and advances all of them together:

    propose(slot, payloads)      RSCodeword::from_data + compute_parity + per-peer subset_copy for every group
                                 (rspaxos/request.rs:72-142)  ->  ReedSolomon.encode_uniform into the shard planes;
                                 the instance enters Accepting under the group's prepared ballot
    on_accept_replies(records)   handle_msg_accept_reply's filters (rspaxos/messages.rs:395-437) -> ack_ingest
    tick()                       count() >= majority + fault_tolerance (rspaxos/messages.rs:438-440) and the commit_bar
                                 advance (rspaxos/durability.rs:144-186) -> tally_planes; committed instances leave
                                 Accepting, so later acks for them are dropped exactly as the reference drops them

This is host-side plumbing over the C ABI (PyTorch holds the device memory): every tally, filter and shard byte is
computed by the CUDA kernels.  A Rust host would make the same calls through INTEGRATION.md's extern block.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .api import Context, ReedSolomon, round_up, shard_len


class RSPaxosLeaderEngine:
    SLOTS = 64

    def __init__(self, ctx: Context, n_groups: int, population: int = 5, fault_tolerance: int = 1,
                 data_len: int = 4096, device: Optional[torch.device] = None, keep_slots: int = 1):
        self.ctx = ctx
        self.G = n_groups
        self.n = population
        self.majority = population // 2 + 1
        self.f = fault_tolerance
        if fault_tolerance > population - self.majority:
            raise ValueError("invalid fault_tolerance")          # rspaxos/mod.rs:599-605
        self.threshold = self.majority + self.f
        self.data_len = data_len
        self.dev = device or torch.device("cuda", ctx.device)
        self.rs = ReedSolomon(ctx, self.majority, population - self.majority)
        self.L = shard_len(data_len, self.majority)
        self.ds = round_up(self.L, 16)
        G = n_groups
        self.planes = torch.zeros((population, G), dtype=torch.int64, device=self.dev)
        self.bal_prepared = torch.ones(G, dtype=torch.int64, device=self.dev)
        self.inst_bal = torch.zeros(G * self.SLOTS, dtype=torch.int64, device=self.dev)
        self.accepting = torch.zeros(G, dtype=torch.int64, device=self.dev)
        self.committed = torch.zeros(G, dtype=torch.int64, device=self.dev)
        self.commit_bar = torch.zeros(G, dtype=torch.int32, device=self.dev)
        # shard store of the most recent `keep_slots` proposals: [slot % keep][d+p][G][ds]
        self.keep = keep_slots
        self.shards = torch.zeros((keep_slots, population, G, self.ds), dtype=torch.uint8, device=self.dev)

    def set_prepared_ballots(self, ballots: torch.Tensor) -> None:
        self.bal_prepared.copy_(ballots)

    def propose(self, slot: int, payloads: torch.Tensor) -> torch.Tensor:
        """payloads: uint8 [G, >= data_len] on the device.  Returns the shard planes [population, G, ds] of this slot:
        plane r is the packed buffer of Accept payloads for replica r."""
        assert 0 <= slot < self.SLOTS and payloads.shape[0] == self.G
        sh = self.shards[slot % self.keep]
        d = self.majority
        ps = self.G * self.ds
        from ._lib import check
        check(self.ctx.lib.ss_rs_encode_uniform_dev(self.rs.h, payloads.data_ptr(), payloads.shape[1], self.data_len, self.G,
                                                    sh[d].data_ptr(), ps, self.ds, 1 | 2))     # padded16 | emit data planes
        # the instance enters Accepting under bal_prepared (rspaxos/request.rs:97-99)
        bit = (1 << slot) if slot < 63 else -(1 << 63)
        self.accepting |= bit
        self.inst_bal.view(self.G, self.SLOTS)[:, slot] = self.bal_prepared
        # a new Accept round starts with an empty ack set
        self.planes &= ~bit
        self.committed &= ~bit
        return sh

    def on_accept_replies(self, rec_group: torch.Tensor, rec_slot: torch.Tensor, rec_peer: torch.Tensor,
                          rec_ballot: torch.Tensor) -> None:
        self.ctx.ack_ingest(rec_group, rec_slot, rec_peer, rec_ballot, self.bal_prepared, self.inst_bal, self.accepting,
                            self.n, self.planes)

    def tick(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """Tallies every window; newly committed instances leave Accepting.  Returns (committed words, commit_bar)."""
        newly, _ = self.ctx.tally_planes(self.planes, self.threshold, want_bar=False)
        newly &= self.accepting                      # only instances that were in Accepting can commit
        self.committed |= newly
        self.accepting &= ~newly                     # Status::Committed: later acks are ignored (messages.rs:419-424)
        # commit_bar over everything committed so far
        full, bar = self.ctx.tally_planes(self.committed.view(1, self.G), 1, want_bar=True, commit_bar=self.commit_bar)
        return self.committed, self.commit_bar
