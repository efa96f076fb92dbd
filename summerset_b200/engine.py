"""Python binding of the batched multi-group consensus engine (ss_engine_* in include/summerset_b200.h).

The engine itself is native: the state of G independent replica groups lives in HBM and every transition --
propose bookkeeping, AcceptReply / AppendEntriesReply ingest, the fused commit tick -- is a CUDA kernel behind the C
ABI (summerset_b200/csrc/engine.cu).  This module only marshals arguments and wraps the engine's device arrays as
torch tensors (zero-copy views) so tests can read results and install state; it runs no torch ops on them.

    LeaderEngine(ctx, protocol, G, population, ...)   protocol in {"multipaxos", "rspaxos", "crossword"}
        propose(slot, payloads[, policy])   multipaxos/request.rs:112-221, rspaxos/request.rs:72-142, crossword/request.rs:82-185
        on_accept_replies(records)          handle_msg_accept_reply filters (multipaxos/messages.rs:377-409, ...)
        tick()                              commit decision + commit_bar, one kernel
    RaftEngine(ctx, G, population[, fault_tolerance, craft])
        append(n_new) / on_append_replies(records) / tick()      raft/messages.rs:243-309, craft/messages.rs:300-308
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check
from .api import Context

_PROTO = {"multipaxos": _lib.SS_PROTO_MULTIPAXOS, "rspaxos": _lib.SS_PROTO_RSPAXOS, "crossword": _lib.SS_PROTO_CROSSWORD,
          "raft": _lib.SS_PROTO_RAFT, "craft": _lib.SS_PROTO_CRAFT}


class _DevArray:
    """__cuda_array_interface__ carrier for a raw device pointer owned by the engine"""

    def __init__(self, ptr: int, shape: Tuple[int, ...], typestr: str):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


class _EngineBase:
    def __init__(self, ctx: Context, cfg: _lib.EngineConfig, n_groups: int):
        self.ctx, self.lib, self.G = ctx, ctx.lib, n_groups
        h = C.c_void_p()
        check(self.lib.ss_engine_create(ctx.h, C.byref(cfg), n_groups, C.byref(h)))
        self.h = h
        self.view = _lib.EngineView()
        check(self.lib.ss_engine_view_get(self.h, C.byref(self.view)))
        self.dev = torch.device("cuda", ctx.device)

    def _tensor(self, ptr: Optional[int], shape: Tuple[int, ...], typestr: str, dtype: torch.dtype) -> Optional[torch.Tensor]:
        if not ptr:
            return None
        return torch.as_tensor(_DevArray(ptr, shape, typestr), device=self.dev).view(dtype).view(shape)

    def close(self) -> None:
        if getattr(self, "h", None) is not None and self.h:
            self.lib.ss_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LeaderEngine(_EngineBase):
    SLOTS = 64

    def __init__(self, ctx: Context, protocol: str, n_groups: int, population: int = 5, fault_tolerance: int = 0,
                 data_len: int = 0, keep_slots: int = 1, rs_total_shards: int = 0, rs_data_shards: int = 0):
        assert protocol in ("multipaxos", "rspaxos", "crossword")
        cfg = _lib.EngineConfig(_PROTO[protocol], population, fault_tolerance, data_len, rs_total_shards, rs_data_shards,
                                keep_slots, 0)
        super().__init__(ctx, cfg, n_groups)
        v, G = self.view, n_groups
        self.protocol, self.n = protocol, population
        self.threshold, self.d, self.T, self.L, self.ds = v.threshold, v.data_shards, v.total_shards, v.shard_len, v.shard_stride
        self.keep = max(1, keep_slots)
        self.planes = self._tensor(v.planes, (population, G), "<i8", torch.int64)
        self.bal_prepared = self._tensor(v.bal_prepared, (G,), "<i8", torch.int64)
        self.inst_bal = self._tensor(v.inst_bal, (G * 64,), "<i8", torch.int64)
        self.accepting = self._tensor(v.accepting, (G,), "<i8", torch.int64)
        self.committed = self._tensor(v.committed, (G,), "<i8", torch.int64)
        self.commit_bar = self._tensor(v.commit_bar, (G,), "<i4", torch.int32)
        self.shards = self._tensor(v.shards, (self.keep, self.T, G, self.ds), "|u1", torch.uint8) if v.shards else None
        self.policy_idx = self._tensor(v.policy_idx, (G * 64,), "|u1", torch.uint8)

    def set_prepared_ballots(self, ballots: torch.Tensor) -> None:
        assert ballots.is_cuda and ballots.dtype == torch.int64 and ballots.numel() == self.G
        check(self.lib.ss_engine_set_prepared_ballots(self.h, ballots.data_ptr()))

    def set_policies(self, policies: Sequence[Sequence[int]], balanced: bool) -> None:
        pol = np.ascontiguousarray(np.array(policies, dtype=np.uint32))
        assert pol.shape[1] == self.n
        check(self.lib.ss_engine_set_policies(self.h, pol.ctypes.data, pol.shape[0], 1 if balanced else 0))

    def propose(self, slot: int, payloads: Optional[torch.Tensor], policy: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """payloads uint8 [G, >= data_len] on the device (None for MultiPaxos).  Returns the proposal's shard planes
        [T, G, ds] (plane r = the packed Accept payloads for the peer(s) holding shard r), or None for MultiPaxos."""
        base = C.c_void_p()
        stride = payloads.shape[1] if payloads is not None else 0
        check(self.lib.ss_engine_propose(self.h, slot, payloads.data_ptr() if payloads is not None else None, stride,
                                         policy.data_ptr() if policy is not None else None, C.byref(base)))
        return self.shards[slot % self.keep] if self.shards is not None else None

    def on_accept_replies(self, rec_group: torch.Tensor, rec_slot: torch.Tensor, rec_peer: torch.Tensor, rec_ballot: torch.Tensor) -> None:
        assert rec_group.dtype == torch.int32 and rec_slot.dtype == torch.uint8 and rec_peer.dtype == torch.uint8 and rec_ballot.dtype == torch.int64
        check(self.lib.ss_engine_ingest(self.h, rec_group.data_ptr(), rec_slot.data_ptr(), rec_peer.data_ptr(), rec_ballot.data_ptr(),
                                        rec_group.numel()))

    def tick(self, newly: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """One fused kernel.  Returns (committed words, commit_bar): views of the engine state."""
        check(self.lib.ss_engine_tick(self.h, newly.data_ptr() if newly is not None else None))
        return self.committed, self.commit_bar


class RSPaxosLeaderEngine(LeaderEngine):
    """RSPaxos leader (rspaxos/request.rs:72-142, rspaxos/messages.rs:395-465)."""

    def __init__(self, ctx: Context, n_groups: int, population: int = 5, fault_tolerance: int = 1, data_len: int = 4096,
                 device=None, keep_slots: int = 1):
        super().__init__(ctx, "rspaxos", n_groups, population, fault_tolerance, data_len, keep_slots)


class RaftEngine(_EngineBase):
    def __init__(self, ctx: Context, n_groups: int, population: int = 7, fault_tolerance: int = 0, craft: bool = False,
                 window: int = 64):
        cfg = _lib.EngineConfig(_PROTO["craft" if craft else "raft"], population, fault_tolerance, 0, 0, 0, 0, window)
        super().__init__(ctx, cfg, n_groups)
        v, G, P = self.view, n_groups, population - 1
        self.P, self.W, self.threshold = P, v.raft_window, v.threshold
        self.match = self._tensor(v.match, (P, G), "<i4", torch.int32)
        self.next_slot = self._tensor(v.next_slot, (P, G), "<i4", torch.int32)
        self.last_commit = self._tensor(v.last_commit, (G,), "<i4", torch.int32)
        self.log_end = self._tensor(v.log_end, (G,), "<i4", torch.int32)
        self.curr_term = self._tensor(v.curr_term, (G,), "<i4", torch.int32)
        self.last_snap = self._tensor(v.last_snap, (G,), "<i4", torch.int32)
        self.terms = self._tensor(v.terms, (G, self.W), "<i4", torch.int32)

    def append(self, n_new: torch.Tensor) -> None:
        assert n_new.is_cuda and n_new.dtype == torch.int32 and n_new.numel() == self.G
        check(self.lib.ss_engine_raft_append(self.h, n_new.data_ptr()))

    def on_append_replies(self, rec_group: torch.Tensor, rec_peer: torch.Tensor, rec_end_slot: torch.Tensor) -> None:
        assert rec_group.dtype == torch.int32 and rec_peer.dtype == torch.uint8 and rec_end_slot.dtype == torch.int32
        check(self.lib.ss_engine_raft_ingest(self.h, rec_group.data_ptr(), rec_peer.data_ptr(), rec_end_slot.data_ptr(), rec_group.numel()))

    def tick(self) -> Tuple[torch.Tensor, torch.Tensor]:
        check(self.lib.ss_engine_tick(self.h, None))
        return self.last_commit, self.last_snap
