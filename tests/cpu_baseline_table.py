#!/usr/bin/env python
"""CPU baseline table of BASELINE.md section 3 (B1-B6): the oracle port of the reference path timed on this box's
host cores -- scalar MUL_TABLE loops (the crate's default) and AVX2 vpshufb nibble tables (what `rse-simd`
enables, scripts/utils/file.py:33), on 1 thread (the reference runs encode + tally inline on one event-loop
thread per replica) and on all threads.  Baseline only; never a fallback.  Writes a text table to stdout.

  python tests/cpu_baseline_table.py [--quick]
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import pyoracle as oracle  # noqa: E402
from summerset_b200 import workloads as wl  # noqa: E402


def best_of(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    return best


def main():
    quick = "--quick" in sys.argv
    nthr = oracle.max_threads()
    modes = [(0, "scalar")] + ([(1, "avx2")] if oracle.have_avx2() else [])
    thr = sorted({1, nthr})
    print(f"# host threads available: {nthr}; AVX2: {oracle.have_avx2()}")
    # B1: MultiPaxos tally, planes form
    G = 1 << (16 if quick else 20)
    planes = wl.cfg2_planes(G, 5, 0.9)
    print("\nB1 MultiPaxos tally n=5 thr=3, %d groups x 64 slots (bit-plane batch form)" % G)
    for t in thr:
        dt = best_of(lambda: oracle.tally_planes(planes, 3, threads=t))
        print(f"  threads={t:3d}  {G * 64 / dt / 1e9:8.3f} G slots/s   {G * 52 / dt / 1e9:7.2f} GB/s")
    # per-ack incremental handler (single thread by nature)
    Gs = 1 << 12
    pl = wl.cfg2_planes(Gs, 5, 0.9)
    rec = wl.ack_records(pl)
    status = np.full(Gs * 64, oracle.ST_ACCEPTING, dtype=np.uint8)
    acks = np.zeros(Gs * 64, dtype=np.uint16)
    t0 = time.perf_counter()
    oracle.tally_stream(rec["rec_group"], rec["rec_slot"], rec["rec_peer"], rec["rec_ballot"], 64, 5, 3,
                        rec["bal_prepared"], rec["inst_bal"], status, acks)
    dt = time.perf_counter() - t0
    print(f"  per-ack incremental handler restatement (1 thread): {len(rec['rec_group']) / dt / 1e6:.1f} M acks/s "
          f"= {Gs * 64 / dt / 1e6:.1f} M slots/s")
    # B2: RSPaxos RS(3,2) encode, 4 KB
    n = 1 << (14 if quick else 18)
    data = wl.payload_uniform(n, 4096, seed_extra=7)
    L = 1366
    print("\nB2 RS(3,2) encode, %d codewords x 4096 B (from_data split + compute_parity)" % n)
    ds_ = 1376
    par_ = np.zeros((2, n, ds_), dtype=np.uint8)          # preallocated and touched: no page faults in the timed call
    off_ = np.arange(n, dtype=np.uint64) * np.uint64(data.shape[1])
    lens_ = np.full(n, 4096, dtype=np.uint32)
    poff_ = np.arange(n, dtype=np.uint64) * np.uint64(ds_)
    for m, name in modes:
        for t in thr:
            dt = best_of(lambda: oracle.rs_encode_batch(3, 2, data.reshape(-1), off_, lens_, par_.reshape(-1), n * ds_, poff_, m, t), 3)
            print(f"  {name:6s} threads={t:3d}  {5 * L * n / dt / 1e9:7.2f} GB/s shard   {4096 * n / dt / 1e9:7.2f} GB/s payload")
    # B3: reconstruct_data with the cfg-3b erasure mix
    nd = 1 << (13 if quick else 17)
    dsub = data[:nd]
    ds = 1376
    full = np.zeros((5, nd, ds), dtype=np.uint8)
    padded = np.zeros((nd, 3 * L), dtype=np.uint8); padded[:, :4096] = dsub[:, :4096]
    for i in range(3):
        full[i, :, :L] = padded[:, i * L:(i + 1) * L]
    full[3:] = oracle.rs_encode_uniform(3, 2, dsub, 4096, mode=modes[-1][0], threads=nthr)
    present = wl.erasure_patterns(nd, 3, 2)
    off = np.arange(nd, dtype=np.uint64) * np.uint64(ds)
    lens = np.full(nd, 4096, dtype=np.uint32)
    miss = sum(((present >> i) & 1) == 0 for i in range(3)).astype(np.int64)
    alg = int(((miss > 0) * 3 * L + miss * L).sum())
    print("\nB3 RS(3,2) reconstruct_data, %d codewords, 50%% intact / 25%% one data / 25%% two shards missing" % nd)
    for m, name in modes:
        for t in thr:
            work = full.copy()
            dt = best_of(lambda: oracle.rs_reconstruct_batch(3, 2, work.reshape(-1), nd * ds, off, lens, present, True, mode=m, threads=t), 2)
            print(f"  {name:6s} threads={t:3d}  {alg / dt / 1e9:7.2f} GB/s   {nd / dt / 1e6:7.2f} M codewords/s")
    # B4: Crossword ragged encode
    n4 = 1 << (11 if quick else 15)
    lens4, spr = wl.cfg4_lengths(n4)
    lay = wl.ragged_layout(lens4, 3)
    arena = np.random.default_rng(1).integers(0, 256, lay["data_bytes"] + 64, dtype=np.uint8)
    par = np.zeros((2, lay["plane_bytes"]), dtype=np.uint8)
    alg4 = int((lay["L"].astype(np.int64) * 5).sum())
    print("\nB4 Crossword mixed sizes 256 B..64 KB, %d codewords (%.2f GB payload), RS(3,2) encode" % (n4, lens4.astype(np.int64).sum() / 1e9))
    for m, name in modes:
        for t in thr:
            dt = best_of(lambda: oracle.rs_encode_batch(3, 2, arena, lay["data_off"], lens4, par.reshape(-1), lay["plane_bytes"], lay["par_off"], m, t), 2)
            print(f"  {name:6s} threads={t:3d}  {alg4 / dt / 1e9:7.2f} GB/s shard")
    # B5: Raft scan
    G5 = 1 << (16 if quick else 20)
    w = wl.cfg5_raft(G5, 7, 64)
    print("\nB5 Raft n=7 commit scan, %d groups, 64-slot term window (loop restatement of raft/messages.rs:256-275)" % G5)
    for t in thr:
        dt = best_of(lambda: oracle.raft_scan_batch(w["match"], w["last_commit"], w["log_end"], w["curr_term"], w["terms"], 4, threads=t))
        print(f"  threads={t:3d}  {G5 / dt / 1e6:8.2f} M groups/s   {G5 * 296 / dt / 1e9:7.2f} GB/s at 296 B/group")
    # B6: criterion-comparable single codeword table (benches/rse_bench.rs:19-26)
    print("\nB6 RS(3,2) from_data + compute_parity of ONE codeword, 1 thread (mirrors benches/rse_bench.rs sizes)")
    for size in [4096, 16 << 10, 64 << 10, 256 << 10, 1 << 20, 4 << 20]:
        one = wl.payload_uniform(1, size, alphanumeric=True, seed_extra=size)
        row = []
        for m, name in modes:
            reps = max(3, (64 << 20) // size // 4)
            t0 = time.perf_counter()
            for _ in range(reps):
                oracle.rs_encode_uniform(3, 2, one, size, mode=m, threads=1)
            row.append(f"{name} {(time.perf_counter() - t0) / reps * 1e3:9.4f} ms")
        print(f"  {size:8d} B   " + "   ".join(row))


if __name__ == "__main__":
    main()
