"""Seeded synthetic inputs for BASELINE.json's configs.  numpy only.

Seeds: SEED_BASE + config number, PRNG = numpy Philox.  The same arrays feed the CPU oracle and
the GPU path, so every comparison is on identical bytes.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

SEED_BASE = 0x5EED0001
G_1M = 1 << 20
G_4M = 1 << 22


def rng_for(cfg: int, extra: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=SEED_BASE + cfg + (extra << 8)))


def bernoulli_words(rng: np.random.Generator, n_words: int, p: float) -> np.ndarray:
    """n_words uint64 whose bits are independently 1 with probability round(p*256)/256."""
    thr = int(round(p * 256))
    out = np.empty(n_words, dtype=np.uint64)
    chunk = 1 << 17
    for a in range(0, n_words, chunk):
        b = min(n_words, a + chunk)
        bits = rng.integers(0, 256, size=(b - a) * 64, dtype=np.uint8) < thr
        out[a:b] = np.packbits(bits.reshape(-1, 64), axis=1, bitorder="little").view(np.uint64).reshape(-1)
    return out


def cfg2_planes(G: int, R: int = 5, p_ack: float = 0.9, seed_extra: int = 0) -> np.ndarray:
    """MultiPaxos n=R: ack bit-planes uint64 [R, G]; replica 0 is the leader (always acked)."""
    rng = rng_for(2, seed_extra)
    planes = np.empty((R, G), dtype=np.uint64)
    planes[0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    for r in range(1, R):
        planes[r] = bernoulli_words(rng, G, p_ack)
    return planes


def ack_records(planes: np.ndarray, seed_extra: int = 0, dup_frac: float = 0.05, stale_frac: float = 0.05
                ) -> Dict[str, np.ndarray]:
    """Turns planes into an AcceptReply record stream (group, slot, peer, ballot) plus noise that
    handle_msg_accept_reply must ignore: duplicates, stale-ballot replies, replies for instances not
    in Accepting state, out-of-range peers.  Returns the stream and the leader state it refers to.
    The valid acks of instances that ARE accepting reproduce `planes & accepting`."""
    rng = rng_for(2, 100 + seed_extra)
    R, G = planes.shape
    bal_prepared = rng.integers(1, 1 << 40, size=G, dtype=np.uint64)
    # a few instances are not Accepting (already Committed / Null): their acks must be dropped
    accepting = bernoulli_words(rng, G, 0.97)
    inst_bal = np.repeat(bal_prepared, 64).reshape(G, 64).copy()
    # some instances were accepted under an older ballot (inst.bal < bal_prepared): still fine (>=)
    u = rng.random((G, 64))
    inst_bal[u < 0.05] -= np.uint64(1)
    # ... and a few under a NEWER one (ballot < inst.bal): their acks are dropped (messages.rs:394-399)
    inst_bal[u > 0.99] += np.uint64(1)
    gs, ss, ps = [], [], []
    for r in range(R):
        bits = np.unpackbits(planes[r].view(np.uint8).reshape(G, 8), axis=1, bitorder="little")
        g, s = np.nonzero(bits)
        gs.append(g.astype(np.uint32)); ss.append(s.astype(np.uint8)); ps.append(np.full(len(g), r, dtype=np.uint8))
    g = np.concatenate(gs); s = np.concatenate(ss); p = np.concatenate(ps)
    b = bal_prepared[g]
    n = len(g)
    # duplicates of valid acks
    di = rng.integers(0, n, size=int(n * dup_frac))
    # stale-ballot acks on random (g, s, peer): ballot != bal_prepared
    m = int(n * stale_frac)
    sg = rng.integers(0, G, size=m).astype(np.uint32)
    ssl = rng.integers(0, 64, size=m).astype(np.uint8)
    sp = rng.integers(0, R, size=m).astype(np.uint8)
    delta = rng.integers(1, 5, size=m) * np.where(rng.random(m) < 0.5, 1, -1)
    sb = (bal_prepared[sg].astype(np.int64) + delta).astype(np.uint64)
    # out-of-range peers
    k = max(1, n // 1000)
    og = rng.integers(0, G, size=k).astype(np.uint32)
    osl = rng.integers(0, 64, size=k).astype(np.uint8)
    op = np.full(k, R, dtype=np.uint8)
    ob = bal_prepared[og]
    rec_g = np.concatenate([g, g[di], sg, og])
    rec_s = np.concatenate([s, s[di], ssl, osl])
    rec_p = np.concatenate([p, p[di], sp, op])
    rec_b = np.concatenate([b, b[di], sb, ob])
    perm = rng.permutation(len(rec_g))
    return dict(rec_group=rec_g[perm], rec_slot=rec_s[perm], rec_peer=rec_p[perm], rec_ballot=rec_b[perm],
                bal_prepared=bal_prepared, inst_bal=inst_bal.reshape(-1), accepting=accepting)


def payload_uniform(n: int, data_len: int, stride: int = 0, alphanumeric: bool = False, cfg: int = 3,
                    seed_extra: int = 0) -> np.ndarray:
    """n payloads of data_len bytes at a stride (default: data_len rounded up to 16): uint8 [n, stride]."""
    rng = rng_for(cfg, seed_extra)
    stride = stride or (data_len + 15) // 16 * 16
    out = np.zeros((n, stride), dtype=np.uint8)
    if alphanumeric:   # benches/rse_bench.rs:37-43
        alpha = np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789", dtype=np.uint8)
        out[:, :data_len] = alpha[rng.integers(0, len(alpha), size=(n, data_len))]
    else:
        out[:, :data_len] = rng.integers(0, 256, size=(n, data_len), dtype=np.uint8)
    # bytes past data_len inside the stride are deliberately NON-zero: the kernel must mask them
    if stride > data_len:
        out[:, data_len:] = 0xA5
    return out


def erasure_patterns(n: int, d: int, p: int, seed_extra: int = 0) -> np.ndarray:
    """cfg 3b: present masks uint32 [n]: 50% nothing missing, 25% one data shard, 25% two shards (>= 1 data)."""
    rng = rng_for(3, 50 + seed_extra)
    t = d + p
    full = (1 << t) - 1
    present = np.full(n, full, dtype=np.uint32)
    kind = rng.integers(0, 4, size=n)
    one = kind == 2
    present[one] &= ~(np.uint32(1) << rng.integers(0, d, size=int(one.sum())).astype(np.uint32))
    two = kind == 3
    a = rng.integers(0, d, size=int(two.sum())).astype(np.uint32)
    b = rng.integers(0, t - 1, size=int(two.sum())).astype(np.uint32)
    b = np.where(b >= a, b + 1, b).astype(np.uint32)      # any other shard
    present[two] &= ~((np.uint32(1) << a) | (np.uint32(1) << b))
    return present


CFG4_SIZES = np.array([256 << i for i in range(9)], dtype=np.uint32)   # 256 .. 65536


def cfg4_lengths(n: int, seed_extra: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """Crossword mixed sizes: data_len uniform over powers of two 256..65536; spr uniform in {1,2,3}."""
    rng = rng_for(4, seed_extra)
    lens = CFG4_SIZES[rng.integers(0, len(CFG4_SIZES), size=n)]
    spr = rng.integers(1, 4, size=n).astype(np.uint8)
    return lens, spr


def ragged_layout(lens: np.ndarray, d: int, align: int = 16) -> Dict[str, np.ndarray]:
    """CSR offsets for a ragged batch: payload offsets (16-aligned) and padded parity offsets."""
    lens = lens.astype(np.uint64)
    L = (lens + np.uint64(d) - np.uint64(1)) // np.uint64(d)
    a = np.uint64(align)
    dpad = (lens + a - np.uint64(1)) // a * a
    ppad = (L + a - np.uint64(1)) // a * a
    data_off = np.concatenate([[np.uint64(0)], np.cumsum(dpad)[:-1]]).astype(np.uint64)
    par_off = np.concatenate([[np.uint64(0)], np.cumsum(ppad)[:-1]]).astype(np.uint64)
    return dict(L=L.astype(np.uint32), data_off=data_off, par_off=par_off, data_bytes=int(dpad.sum()),
                plane_bytes=int(ppad.sum()))


def cfg5_raft(G: int, n: int = 7, W: int = 64, seed_extra: int = 0) -> Dict[str, np.ndarray]:
    """Raft n replicas: per group n-1 peer match indices, last_commit, log_end = last_commit+1+W, a
    non-decreasing term window whose (possibly empty) suffix is curr_term."""
    rng = rng_for(5, seed_extra)
    P = n - 1
    last_commit = rng.integers(0, 1 << 20, size=G).astype(np.uint32)
    log_end = (last_commit + np.uint32(1 + W)).astype(np.uint32)
    adv = rng.geometric(0.05, size=(P, G)).astype(np.uint32) - np.uint32(1)
    match = np.minimum(last_commit[None, :] + adv, log_end[None, :] - np.uint32(1)).astype(np.uint32)
    curr_term = rng.integers(3, 1000, size=G).astype(np.uint32)
    suffix = rng.integers(0, W + 1, size=G)            # number of trailing entries in curr_term
    idx = np.arange(W)[None, :]
    is_cur = idx >= (W - suffix)[:, None]
    older = (curr_term[:, None] - np.uint32(1) - (rng.integers(0, 2, size=(G, W)).cumsum(axis=1)[:, ::-1] // 8).astype(np.uint32))
    terms = np.where(is_cur, curr_term[:, None], np.minimum(older, curr_term[:, None] - np.uint32(1))).astype(np.uint32)
    return dict(match=match, last_commit=last_commit, log_end=log_end, curr_term=curr_term, terms=terms)
