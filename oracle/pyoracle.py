"""ctypes/numpy wrapper of the CPU oracle (oracle/libss_oracle.so).

TEST INFRASTRUCTURE ONLY (see oracle/ss_oracle.h): imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs.  Never imported by summerset_b200/.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

_DIR = Path(__file__).resolve().parent
LIB_PATH = _DIR / "libss_oracle.so"
_lib = None


def build(force: bool = False) -> Path:
    src = [_DIR / "ss_oracle.c", _DIR / "ss_wire.c", _DIR / "ss_oracle.h", _DIR / "Makefile"]
    if force or not LIB_PATH.exists() or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in src):
        subprocess.run(["make", "-C", str(_DIR), "-B" if force else "-s"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        L = C.CDLL(str(LIB_PATH))
        L.ssor_gf_mul.restype = C.c_uint8
        L.ssor_gf_mul.argtypes = [C.c_uint8, C.c_uint8]
        L.ssor_gf_div.restype = C.c_uint8
        L.ssor_gf_div.argtypes = [C.c_uint8, C.c_uint8]
        L.ssor_gf_exp.restype = C.c_uint8
        L.ssor_gf_exp.argtypes = [C.c_uint8, C.c_uint]
        L.ssor_gf_log.restype = C.c_uint8
        L.ssor_gf_log.argtypes = [C.c_uint8]
        L.ssor_gf_exp_table.restype = C.c_uint8
        L.ssor_gf_exp_table.argtypes = [C.c_uint]
        L.ssor_cw_shard_len.restype = C.c_size_t
        L.ssor_cw_shard_len.argtypes = [C.c_size_t, C.c_int]
        L.ssor_commit_bar.restype = C.c_uint32
        L.ssor_commit_bar.argtypes = [C.c_uint64]
        L.ssor_cw_min_spr.restype = C.c_uint32
        L.ssor_cw_coverage.restype = C.c_uint32
        L.ssor_raft_scan.restype = C.c_uint32
        L.ssor_raft_snap_scan.restype = C.c_uint32
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- GF / matrix -------------------------------------------------------------------------------
def gf_mul(a: int, b: int) -> int:
    return int(lib().ssor_gf_mul(a, b))


def gf_exp(a: int, n: int) -> int:
    return int(lib().ssor_gf_exp(a, n))


def rs_matrix(d: int, p: int) -> np.ndarray:
    m = np.zeros((d + p, d), dtype=np.uint8)
    rc = lib().ssor_rs_build_matrix(d, p, _p(m))
    if rc != 0:
        raise ValueError(f"ssor_rs_build_matrix({d},{p}) -> {rc}")
    return m


def rs_new_rc(d: int, p: int) -> int:
    """error code ReedSolomon::new(d, p) would give (0 = ok)"""
    m = np.zeros((max(d + p, 1), max(d, 1)), dtype=np.uint8)
    return int(lib().ssor_rs_build_matrix(d, p, _p(m)))


def decode_matrix(d: int, p: int, present: Sequence[int]):
    pres = np.array(present, dtype=np.uint8)
    src = np.zeros(d, dtype=np.int32)
    dec = np.zeros((d, d), dtype=np.uint8)
    rc = lib().ssor_rs_decode_matrix(d, p, _p(pres), _p(src), _p(dec))
    return rc, src, dec


# ---- single codeword ---------------------------------------------------------------------------
def _ptrs(shards: Sequence[np.ndarray]):
    return (C.c_void_p * len(shards))(*[s.ctypes.data for s in shards])


def rs_encode(d: int, p: int, shards: List[np.ndarray]) -> int:
    return int(lib().ssor_rs_encode(d, p, _ptrs(shards), C.c_size_t(len(shards[0]))))


def rs_reconstruct(d: int, p: int, shards: List[Optional[np.ndarray]], data_only: bool) -> int:
    L = next(len(s) for s in shards if s is not None) if any(s is not None for s in shards) else 0
    pres = np.array([0 if s is None else 1 for s in shards], dtype=np.uint8)
    bufs = [s if s is not None else np.zeros(max(L, 1), dtype=np.uint8) for s in shards]
    rc = int(lib().ssor_rs_reconstruct(d, p, _ptrs(bufs), _p(pres), C.c_size_t(L), 1 if data_only else 0))
    if rc == 0:
        for i in range(len(shards)):
            if shards[i] is None and pres[i]:
                shards[i] = bufs[i]
    return rc


def rs_verify(d: int, p: int, shards: List[np.ndarray]):
    ok = C.c_int(0)
    rc = int(lib().ssor_rs_verify(d, p, _ptrs(shards), C.c_size_t(len(shards[0])), C.byref(ok)))
    return rc, bool(ok.value)


def cw_shard_len(data_len: int, d: int) -> int:
    return int(lib().ssor_cw_shard_len(data_len, d))


def cw_split(data: bytes, d: int) -> np.ndarray:
    L = cw_shard_len(len(data), d)
    out = np.zeros((d, L), dtype=np.uint8)
    src = np.frombuffer(data, dtype=np.uint8)
    lib().ssor_cw_split(_p(np.ascontiguousarray(src)), C.c_size_t(len(data)), d, _p(out))
    return out


# ---- batched -----------------------------------------------------------------------------------
def rs_encode_batch(d: int, p: int, data: np.ndarray, data_off: np.ndarray, data_len: np.ndarray,
                    parity: np.ndarray, plane_stride: int, par_off: np.ndarray, mode: int = 0,
                    threads: int = 1) -> None:
    assert data.dtype == np.uint8 and parity.dtype == np.uint8
    data_off = np.ascontiguousarray(data_off, dtype=np.uint64)
    data_len = np.ascontiguousarray(data_len, dtype=np.uint32)
    par_off = np.ascontiguousarray(par_off, dtype=np.uint64)
    rc = lib().ssor_rs_encode_batch(d, p, _p(data), _p(data_off), _p(data_len), C.c_uint64(len(data_len)), _p(parity),
                                    C.c_uint64(plane_stride), _p(par_off), mode, threads)
    if rc != 0:
        raise ValueError(f"ssor_rs_encode_batch -> {rc}")


def rs_encode_uniform(d: int, p: int, data: np.ndarray, data_len: int, mode: int = 0, threads: int = 1,
                      shard_stride: Optional[int] = None) -> np.ndarray:
    """data uint8 [n, stride] -> parity uint8 [p, n, shard_stride] (zero beyond L)."""
    n, stride = data.shape
    L = cw_shard_len(data_len, d)
    ss = shard_stride or (L + 15) // 16 * 16
    parity = np.zeros((p, n, ss), dtype=np.uint8)
    off = np.arange(n, dtype=np.uint64) * np.uint64(stride)
    lens = np.full(n, data_len, dtype=np.uint32)
    poff = np.arange(n, dtype=np.uint64) * np.uint64(ss)
    rs_encode_batch(d, p, data.reshape(-1), off, lens, parity.reshape(-1), n * ss, poff, mode, threads)
    return parity


def rs_reconstruct_batch(d: int, p: int, shards: np.ndarray, plane_stride: int, off: np.ndarray,
                         data_len: np.ndarray, present: np.ndarray, data_only: bool, mode: int = 0,
                         threads: int = 1) -> np.ndarray:
    off = np.ascontiguousarray(off, dtype=np.uint64)
    data_len = np.ascontiguousarray(data_len, dtype=np.uint32)
    present = np.ascontiguousarray(present, dtype=np.uint32)
    status = np.zeros(len(data_len), dtype=np.int32)
    rc = lib().ssor_rs_reconstruct_batch(d, p, _p(shards), C.c_uint64(plane_stride), _p(off), _p(data_len),
                                         _p(present), C.c_uint64(len(data_len)), 1 if data_only else 0, _p(status),
                                         mode, threads)
    if rc != 0:
        raise ValueError(f"ssor_rs_reconstruct_batch -> {rc}")
    return status


MODE_STATIC = 0x100     # SSOR_MODE_STATIC: contiguous block of codewords per OpenMP thread


def first_touch_fill(buf: np.ndarray, n_items: int, item_bytes: int, item_stride: int, seed: int, zero: bool,
                     threads: int) -> None:
    """Fill (or zero) buf in the static thread partition the encode loop uses, so pages are placed NUMA-locally."""
    assert buf.dtype == np.uint8 and buf.flags.c_contiguous and buf.size >= (n_items - 1) * item_stride + item_bytes
    lib().ssor_first_touch_fill(_p(buf), C.c_uint64(n_items), C.c_uint64(item_bytes), C.c_uint64(item_stride),
                                C.c_uint64(seed), 1 if zero else 0, threads)


def have_avx2() -> bool:
    return bool(lib().ssor_have_avx2())


def max_threads() -> int:
    return int(lib().ssor_max_threads())


# ---- tallies -----------------------------------------------------------------------------------
ST_NULL, ST_PREPARING, ST_ACCEPTING, ST_COMMITTED, ST_EXECUTED = range(5)


def tally_stream(rec_g, rec_s, rec_p, rec_b, S: int, population: int, threshold: int, bal_prepared, inst_bal,
                 status, acks) -> None:
    rec_g = np.ascontiguousarray(rec_g, dtype=np.uint32)
    rec_s = np.ascontiguousarray(rec_s, dtype=np.uint8)
    rec_p = np.ascontiguousarray(rec_p, dtype=np.uint8)
    rec_b = np.ascontiguousarray(rec_b, dtype=np.uint64)
    assert status.dtype == np.uint8 and acks.dtype == np.uint16
    lib().ssor_tally_stream(_p(rec_g), _p(rec_s), _p(rec_p), _p(rec_b), C.c_uint64(len(rec_g)), S, population,
                            threshold, _p(np.ascontiguousarray(bal_prepared, dtype=np.uint64)),
                            _p(np.ascontiguousarray(inst_bal, dtype=np.uint64)), _p(status), _p(acks))


def tally_stream_crossword(rec_g, rec_s, rec_p, rec_b, S: int, population: int, T: int, d: int, majority: int, f: int,
                           balanced: bool, policies: np.ndarray, policy_idx: np.ndarray, bal_prepared, inst_bal, status,
                           acks) -> None:
    """crossword/messages.rs:481-574 per record; policies uint32 [K, population], policy_idx uint8 [G*S]."""
    rec_g = np.ascontiguousarray(rec_g, dtype=np.uint32)
    rec_s = np.ascontiguousarray(rec_s, dtype=np.uint8)
    rec_p = np.ascontiguousarray(rec_p, dtype=np.uint8)
    rec_b = np.ascontiguousarray(rec_b, dtype=np.uint64)
    policies = np.ascontiguousarray(policies, dtype=np.uint32)
    policy_idx = np.ascontiguousarray(policy_idx, dtype=np.uint8)
    assert status.dtype == np.uint8 and acks.dtype == np.uint16 and policies.shape[1] == population
    lib().ssor_tally_stream_crossword(_p(rec_g), _p(rec_s), _p(rec_p), _p(rec_b), C.c_uint64(len(rec_g)), S, population, T, d,
                                      majority, f, 1 if balanced else 0, _p(policies), policies.shape[0], _p(policy_idx),
                                      _p(np.ascontiguousarray(bal_prepared, dtype=np.uint64)),
                                      _p(np.ascontiguousarray(inst_bal, dtype=np.uint64)), _p(status), _p(acks))


def raft_reply_stream(rec_g, rec_peer, rec_end, npeers: int, threshold: int, next_slot, match, last_commit, last_snap,
                      log_end, curr_term, terms) -> None:
    """raft/messages.rs:221-309 for successful replies, per record, in place on next_slot/match [P,G], last_commit, last_snap."""
    rec_g = np.ascontiguousarray(rec_g, dtype=np.uint32)
    rec_peer = np.ascontiguousarray(rec_peer, dtype=np.uint8)
    rec_end = np.ascontiguousarray(rec_end, dtype=np.uint32)
    G, W = terms.shape
    for a in (next_slot, match, last_commit, last_snap, log_end, curr_term, terms):
        assert a.dtype == np.uint32 and a.flags.c_contiguous
    lib().ssor_raft_reply_stream(_p(rec_g), _p(rec_peer), _p(rec_end), C.c_uint64(len(rec_g)), npeers, threshold, C.c_uint64(G),
                                 _p(next_slot), _p(match), _p(last_commit), _p(last_snap), _p(log_end), _p(curr_term),
                                 _p(terms), W)


def tally_planes(planes: np.ndarray, threshold: int, threads: int = 1):
    planes = np.ascontiguousarray(planes, dtype=np.uint64)
    R, G = planes.shape
    committed = np.zeros(G, dtype=np.uint64)
    bar = np.zeros(G, dtype=np.uint32)
    lib().ssor_tally_planes(_p(planes), R, C.c_uint64(G), threshold, _p(committed), _p(bar), threads)
    return committed, bar


def tally_masks(masks: np.ndarray, threshold: int) -> np.ndarray:
    masks = np.ascontiguousarray(masks, dtype=np.uint16)
    out = np.zeros(len(masks), dtype=np.uint8)
    lib().ssor_tally_masks(_p(masks), C.c_uint64(len(masks)), threshold, _p(out))
    return out


def commit_bar(word: int) -> int:
    return int(lib().ssor_commit_bar(C.c_uint64(word)))


def cw_brr_assignment(n: int, T: int, spr: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint32)
    lib().ssor_cw_brr_assignment(n, T, spr, _p(out))
    return out


def cw_min_spr(d: int, majority: int, f: int, alive: int) -> int:
    return int(lib().ssor_cw_min_spr(d, majority, f, alive))


def cw_coverage(T: int, n: int, ack_mask: int, assignment: np.ndarray, f: int, balanced: bool) -> int:
    a = np.ascontiguousarray(assignment, dtype=np.uint32)
    return int(lib().ssor_cw_coverage(T, n, ack_mask, _p(a), f, 1 if balanced else 0))


def cw_committed(T: int, n: int, d: int, majority: int, f: int, ack_mask: int, assignment: np.ndarray,
                 balanced: bool) -> bool:
    a = np.ascontiguousarray(assignment, dtype=np.uint32)
    return bool(lib().ssor_cw_committed(T, n, d, majority, f, ack_mask, _p(a), 1 if balanced else 0))


def raft_scan(match: Sequence[int], last_commit: int, log_end: int, curr_term: int, terms: np.ndarray,
              threshold: int) -> int:
    m = np.ascontiguousarray(match, dtype=np.uint32)
    t = np.ascontiguousarray(terms, dtype=np.uint32)
    return int(lib().ssor_raft_scan(_p(m), len(m), last_commit, log_end, curr_term, _p(t), threshold))


def raft_snap_scan(match: Sequence[int], last_snap: int, end_slot: int) -> int:
    m = np.ascontiguousarray(match, dtype=np.uint32)
    return int(lib().ssor_raft_snap_scan(_p(m), len(m), last_snap, end_slot))


def raft_scan_batch(match: np.ndarray, last_commit, log_end, curr_term, terms: np.ndarray, threshold: int,
                    threads: int = 1) -> np.ndarray:
    match = np.ascontiguousarray(match, dtype=np.uint32)
    P, G = match.shape
    terms = np.ascontiguousarray(terms, dtype=np.uint32)
    W = terms.shape[1]
    out = np.zeros(G, dtype=np.uint32)
    lib().ssor_raft_scan_batch(_p(match), P, C.c_uint64(G), _p(np.ascontiguousarray(last_commit, dtype=np.uint32)),
                               _p(np.ascontiguousarray(log_end, dtype=np.uint32)),
                               _p(np.ascontiguousarray(curr_term, dtype=np.uint32)), _p(terms), W, threshold, _p(out),
                               threads)
    return out


# ---- CRaft / prepare merge ---------------------------------------------------------------------
def craft_threshold(majority: int, f: int, full_copy: bool) -> int:
    lib().ssor_craft_threshold.restype = C.c_uint32
    return int(lib().ssor_craft_threshold(majority, f, 1 if full_copy else 0))


def craft_shadow_last_commit(match: Sequence[int], threshold: int) -> int:
    m = np.ascontiguousarray(match, dtype=np.uint32)
    lib().ssor_craft_shadow_last_commit.restype = C.c_uint32
    return int(lib().ssor_craft_shadow_last_commit(_p(m), len(m), threshold))


PM_USE, PM_NULL, PM_RECONSTRUCT, PM_PARITY = 1, 2, 4, 8


def prepare_merge_stream(has_vote, bal, mask):
    hv = np.ascontiguousarray(has_vote, dtype=np.uint8)
    b = np.ascontiguousarray(bal, dtype=np.uint64)
    m = np.ascontiguousarray(mask, dtype=np.uint32)
    mb = C.c_uint64(0); mm = C.c_uint32(0)
    lib().ssor_prepare_merge_stream(_p(hv), _p(b), _p(m), len(hv), C.byref(mb), C.byref(mm))
    return int(mb.value), int(mm.value)


def prepare_decide(merged: int, acks_cnt: int, d: int, population: int, f: int) -> int:
    lib().ssor_prepare_decide.restype = C.c_uint32
    return int(lib().ssor_prepare_decide(merged, acks_cnt, d, population, f))


def gossip_targets_excl(me: int, population: int, d: int, src_peer: int, avail: int, assignment, peer_alive: int):
    a = np.ascontiguousarray(assignment, dtype=np.uint32)
    excl = np.zeros(population, dtype=np.uint32)
    lib().ssor_gossip_targets_excl.restype = C.c_uint32
    t = int(lib().ssor_gossip_targets_excl(me, population, d, src_peer, avail, _p(a), peer_alive, _p(excl)))
    return t, excl


# ---- wire / WAL byte formats and reconstruct serving (oracle/ss_wire.c) ---------------------------------------
def _shard_ptrs(shards):
    bufs = [None if s is None else np.ascontiguousarray(np.frombuffer(bytes(s), dtype=np.uint8)) for s in shards]
    arr = (C.c_void_p * len(bufs))(*[None if b is None else b.ctypes.data for b in bufs])
    return arr, bufs


def varint(v: int) -> bytes:
    out = np.zeros(16, dtype=np.uint8)
    lib().ssor_varint_put.restype = C.c_size_t
    n = lib().ssor_varint_put(_p(out), C.c_uint64(v))
    return out[:n].tobytes()


def bitmap_encode(size: int, bits: int) -> bytes:
    out = np.zeros(64, dtype=np.uint8)
    lib().ssor_bitmap_encode.restype = C.c_size_t
    n = lib().ssor_bitmap_encode(size, C.c_uint64(bits), _p(out))
    return out[:n].tobytes()


def frame_accept(accept_variant: int, slot: int, ballot: int, d: int, p: int, data_len: int, shards, assignment=None,
                 assign_size: int = 0) -> bytes:
    """shards: d+p entries, bytes or None.  assignment: list of per-replica shard bitmasks (Crossword) or None."""
    L = next(len(s) for s in shards if s is not None)
    arr, keep = _shard_ptrs(shards)
    out = np.zeros(64 + (d + p) * (L + 16) + (len(assignment) * 16 if assignment else 0), dtype=np.uint8)
    asg = np.ascontiguousarray(np.array(assignment, dtype=np.uint32)) if assignment is not None else None
    lib().ssor_frame_accept.restype = C.c_size_t
    n = lib().ssor_frame_accept(accept_variant, C.c_uint64(slot), C.c_uint64(ballot), d, p, C.c_uint64(data_len), C.c_uint64(L), arr,
                                _p(asg), len(assignment) if assignment is not None else 0, assign_size, _p(out))
    return out[:n].tobytes()


def frame_accept_reply(reply_variant: int, slot: int, ballot: int, with_size: bool = False, size: int = 0) -> bytes:
    out = np.zeros(64, dtype=np.uint8)
    lib().ssor_frame_accept_reply.restype = C.c_size_t
    n = lib().ssor_frame_accept_reply(reply_variant, C.c_uint64(slot), C.c_uint64(ballot), 1 if with_size else 0, C.c_uint64(size), _p(out))
    return out[:n].tobytes()


def parse_accept_reply(frame: bytes, reply_variant: int, with_size: bool = False):
    """-> (frame_len, slot, ballot, size, kind); frame_len -1 = malformed, -2 = another message kind"""
    buf = np.frombuffer(frame, dtype=np.uint8)
    slot, ballot, size, kind = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
    lib().ssor_parse_accept_reply.restype = C.c_long
    n = lib().ssor_parse_accept_reply(_p(np.ascontiguousarray(buf)), C.c_size_t(len(frame)), reply_variant, 1 if with_size else 0,
                                      C.byref(slot), C.byref(ballot), C.byref(size), C.byref(kind))
    return int(n), slot.value, ballot.value, size.value, kind.value


def wal_accept_data(slot: int, ballot: int, d: int, p: int, data_len: int, shards) -> bytes:
    L = next(len(s) for s in shards if s is not None)
    arr, keep = _shard_ptrs(shards)
    out = np.zeros(64 + (d + p) * (L + 16), dtype=np.uint8)
    lib().ssor_wal_accept_data.restype = C.c_size_t
    n = lib().ssor_wal_accept_data(C.c_uint64(slot), C.c_uint64(ballot), d, p, C.c_uint64(data_len), C.c_uint64(L), arr, _p(out))
    return out[:n].tobytes()


def wal_commit_slot(slot: int) -> bytes:
    out = np.zeros(32, dtype=np.uint8)
    lib().ssor_wal_commit_slot.restype = C.c_size_t
    n = lib().ssor_wal_commit_slot(C.c_uint64(slot), _p(out))
    return out[:n].tobytes()


def decode_accept(frame: bytes, kind: int, with_assignment: bool = False):
    """kind 0 = peer Accept frame, 1 = WAL AcceptData.  -> dict or None when malformed."""
    buf = np.ascontiguousarray(np.frombuffer(frame, dtype=np.uint8))
    variant, d, p, n_assign, asz = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    slot, ballot, dl, sl = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    shard_at = np.zeros(256, dtype=np.uint64)
    asg = np.zeros(64, dtype=np.uint32)
    lib().ssor_decode_accept.restype = C.c_long
    n = lib().ssor_decode_accept(_p(buf), C.c_size_t(len(frame)), kind, C.byref(variant), C.byref(slot), C.byref(ballot), C.byref(d),
                                 C.byref(p), C.byref(dl), C.byref(sl), _p(shard_at), 256, 1 if with_assignment else 0, _p(asg),
                                 C.byref(n_assign), C.byref(asz))
    if n < 0:
        return None
    t = d.value + p.value
    L = sl.value
    shards = [None if shard_at[j] == 0 else frame[int(shard_at[j]):int(shard_at[j]) + L] for j in range(t)]
    return dict(length=int(n), variant=variant.value, slot=slot.value, ballot=ballot.value, d=d.value, p=p.value, data_len=dl.value,
                shard_len=L, shards=shards, assignment=[int(x) for x in asg[:n_assign.value]] if with_assignment else None,
                assign_size=asz.value)


def reconstruct_serve_mask(held: int, exclude: int, total_shards: int, status: int) -> int:
    lib().ssor_reconstruct_serve_mask.restype = C.c_uint32
    return int(lib().ssor_reconstruct_serve_mask(held, exclude, total_shards, status))
