"""Oracle self-consistency for the tally / Crossword / Raft restatements (no unit tests exist for
these handlers in the reference -- SURVEY.md 8c -- so the line-by-line incremental restatements are
cross-checked against independent closed forms and the TLA+ predicates)."""
import itertools

import numpy as np
import pytest

from summerset_b200 import workloads as wl


def test_stream_equals_planes_form(oracle):
    """End state of the per-ack handler == batch predicate on the surviving ack set."""
    G, R = 300, 5
    planes = wl.cfg2_planes(G, R, 0.6)
    rec = wl.ack_records(planes)
    for thr in (3, 4):
        status = np.zeros(G * 64, dtype=np.uint8)
        acc_bits = np.unpackbits(rec["accepting"].view(np.uint8).reshape(G, 8), axis=1, bitorder="little").reshape(-1)
        status[acc_bits == 1] = oracle.ST_ACCEPTING
        status[acc_bits == 0] = oracle.ST_COMMITTED
        acks = np.zeros(G * 64, dtype=np.uint16)
        oracle.tally_stream(rec["rec_group"], rec["rec_slot"], rec["rec_peer"], rec["rec_ballot"], 64, R, thr,
                            rec["bal_prepared"], rec["inst_bal"], status, acks)
        # independent model: surviving acks = planes & accepting & (inst_bal <= bal_prepared)
        ok = (rec["inst_bal"].reshape(G, 64) <= rec["bal_prepared"][:, None])
        okw = np.packbits(ok, axis=1, bitorder="little").view(np.uint64).reshape(-1)
        eff = planes & rec["accepting"][None, :] & okw[None, :]
        committed, _ = oracle.tally_planes(eff, thr)
        newly = (status.reshape(G, 64) == oracle.ST_COMMITTED) & (acc_bits.reshape(G, 64) == 1)
        neww = np.packbits(newly, axis=1, bitorder="little").view(np.uint64).reshape(-1)
        assert (neww == (committed & rec["accepting"])).all()
        # once committed, later acks are dropped: the recorded ack set never exceeds the threshold
        cnt = np.array([bin(int(a)).count("1") for a in acks])
        assert cnt.max() <= thr


def test_planes_vs_popcount(oracle):
    planes = wl.cfg2_planes(1000, 7, 0.5)
    bits = np.unpackbits(planes.view(np.uint8).reshape(7, 1000, 8), axis=2, bitorder="little").sum(axis=0)
    for thr in range(0, 9):
        c, bar = oracle.tally_planes(planes, thr)
        want = np.packbits(bits >= thr, axis=1, bitorder="little").view(np.uint64).reshape(-1)
        assert (c == want).all()
        for g in range(0, 1000, 37):
            w = int(c[g]); b = 0
            while b < 64 and (w >> b) & 1:
                b += 1
            assert bar[g] == b == oracle.commit_bar(w)


def test_crossword_balanced_closed_form_is_worst_case_of_enumeration(oracle):
    """crossword/messages.rs:28-33 vs :35-61.  The balanced closed form is the coverage of the WORST
    placement of the acked replicas (adjacent ones, whose round-robin shard runs overlap most), so it is
    a lower bound of the subset enumeration and equal to it whenever the acked replicas are adjacent;
    it is not clamped at T.  Both branches are restated verbatim: the GPU path must follow whichever the
    `assignment_balanced` flag selects, not pick one as "the truth"."""
    for n, T, d in [(5, 5, 3), (3, 3, 2), (7, 7, 4), (5, 10, 6), (5, 15, 9)]:
        majority = n // 2 + 1
        dj = T // n
        for f in range(0, n - majority + 1):
            for spr in range(dj, d + 1, dj):
                asg = oracle.cw_brr_assignment(n, T, spr)
                for ack in range(1 << n):
                    a = oracle.cw_coverage(T, n, ack, asg, f, True)
                    b = oracle.cw_coverage(T, n, ack, asg, f, False)
                    assert min(a, T) <= b, (n, T, d, f, spr, ack, a, b)
                    k = bin(ack).count("1")
                    if k > f:
                        assert a == (k - f - 1) * dj + spr
                        # adjacent (cyclic run of) replicas realise the bound exactly
                        run = sum(1 << ((r0 + i) % n) for r0 in [0] for i in range(k))
                        assert oracle.cw_coverage(T, n, run, asg, f, False) == min(a, T)
                    if oracle.cw_committed(T, n, d, majority, f, ack, asg, True):
                        assert oracle.cw_committed(T, n, d, majority, f, ack, asg, False)


def test_crossword_required_acks_n5(oracle):
    """n=5,d=3,T=5,f=2 => required acks 5/4/3 for spr=1/2/3."""
    for spr, need in [(1, 5), (2, 4), (3, 3)]:
        asg = oracle.cw_brr_assignment(5, 5, spr)
        for ack in range(32):
            na = bin(ack).count("1")
            assert oracle.cw_committed(5, 5, 3, 3, 2, ack, asg, True) == (na >= need)
    assert oracle.cw_min_spr(3, 3, 2, 5) == 1 and oracle.cw_min_spr(3, 3, 2, 4) == 2 and oracle.cw_min_spr(3, 3, 2, 3) == 3


def test_crossword_tla_committed_condition(oracle):
    """tla+/crossword/Crossword.tla:235-240: committed iff every (|acks|-f)-subset still covers >= d shards."""
    rng = np.random.default_rng(3)
    n, T, d, f, majority = 5, 5, 3, 1, 3
    for _ in range(200):
        asg = rng.integers(0, 1 << T, size=n).astype(np.uint32)
        for ack in range(1 << n):
            servers = [r for r in range(n) if (ack >> r) & 1]
            if len(servers) <= f:
                want_cov = 0
            else:
                want_cov = min(bin(int(np.bitwise_or.reduce(asg[list(sub)]))).count("1")
                               for sub in itertools.combinations(servers, len(servers) - f))
            assert oracle.cw_coverage(T, n, ack, asg, f, False) == want_cov


def test_raft_scan_closed_form(oracle):
    w = wl.cfg5_raft(2000, 7, 64)
    got = oracle.raft_scan_batch(w["match"], w["last_commit"], w["log_end"], w["curr_term"], w["terms"], 4)
    P, G = w["match"].shape
    for g in range(G):
        m = np.sort(w["match"][:, g])[::-1][4 - 2]          # (quorum-1)-th largest = 3rd largest of 6
        upper = min(int(m), int(w["log_end"][g]) - 1)
        lc = int(w["last_commit"][g])
        want = lc
        for slot in range(upper, lc, -1):
            if w["terms"][g, slot - lc - 1] == w["curr_term"][g]:
                want = slot
                break
        assert got[g] == want
    assert (got >= w["last_commit"]).all()
    assert (got > w["last_commit"]).mean() > 0.2           # the workload does commit things


def test_raft_edge_cases(oracle):
    terms = np.array([5, 5, 7, 7], dtype=np.uint32)
    # quorum 2 of 3: one peer suffices
    assert oracle.raft_scan([12, 10], 10, 15, 7, terms, 2) == 10   # slots 11,12 have term 5 -> skipped
    assert oracle.raft_scan([14, 10], 10, 15, 7, terms, 2) == 14
    assert oracle.raft_scan([13, 10], 10, 15, 7, terms, 2) == 13
    assert oracle.raft_scan([99, 99], 10, 15, 7, terms, 2) == 14   # clamped by log_end
    assert oracle.raft_scan([99, 99], 10, 15, 9, terms, 2) == 10   # no current-term entry
    assert oracle.raft_scan([99, 99], 10, 11, 7, terms, 2) == 10   # empty range
    assert oracle.raft_scan([], 10, 15, 7, terms, 1) == 14         # single replica
    assert oracle.raft_scan([99], 10, 15, 7, terms, 3) == 10       # unreachable threshold
    assert oracle.raft_snap_scan([12, 15], 10, 15) == 12           # raft/messages.rs:298-309


def test_gloo_two_rank_group_sharding():
    """groups shard across ranks with no data-path collective; the host-side partition
    logic used by bench.py is exercised with world_size 2 over gloo on CPU."""
    import subprocess, sys, os
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    env = dict(os.environ, PYTHONPATH=str(root))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617",
                        str(root / "tests" / "gloo_shard_worker.py")], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "SHARD_OK" in r.stdout
