"""The batched leader engine (summerset_b200/engine.py) over several ticks vs the oracle's incremental per-message
handlers on the identical interleaved stream: proposals enter Accepting, AcceptReplies (with duplicates, stale ballots,
replies for instances not yet proposed or already committed) arrive in batches, commit bitmaps / commit_bar and every
shard plane must match bit for bit (rspaxos/request.rs:72-142, rspaxos/messages.rs:395-465, rspaxos/durability.rs:144-186)."""
import numpy as np
import pytest
import torch

from summerset_b200 import workloads as wl

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint64:
        a = a.view(np.int64)
    elif a.dtype == np.uint32:
        a = a.view(np.int32)
    return torch.from_numpy(a).to(DEV)


@pytest.mark.parametrize("population,f,data_len", [(5, 1, 4096), (5, 0, 300), (7, 1, 1000), (3, 0, 64)])
def test_engine_multi_tick_matches_incremental_handlers(ctx, oracle, population, f, data_len):
    from summerset_b200.engine import RSPaxosLeaderEngine
    rng = np.random.default_rng(population * 10 + f)
    G, S = 1200, 64
    eng = RSPaxosLeaderEngine(ctx, G, population, f, data_len)
    d = population // 2 + 1
    p = population - d
    thr = d + f
    bal = rng.integers(1, 1 << 30, G).astype(np.uint64)
    eng.set_prepared_ballots(_t(bal))
    status = np.zeros(G * S, dtype=np.uint8)
    acks = np.zeros(G * S, dtype=np.uint16)
    inst_bal = np.zeros(G * S, dtype=np.uint64)
    n_rounds = 5
    for rnd in range(n_rounds):
        slot = rnd if rnd < 4 else 63                      # also exercise the top bit of the window
        payload = wl.payload_uniform(G, data_len, seed_extra=100 + rnd)
        sh = eng.propose(slot, _t(payload))
        torch.cuda.synchronize()
        # --- shard planes: data shards are the contiguous split, parity is the oracle's
        L = oracle.cw_shard_len(data_len, d)
        got = sh.cpu().numpy()
        want_par = oracle.rs_encode_uniform(d, p, payload, data_len)
        assert (got[d:] == want_par).all()
        for g in (0, 1, G // 2, G - 1):
            assert (got[:d, g, :L] == oracle.cw_split(payload[g, :data_len].tobytes(), d)).all()
        assert (got[:d, :, L:] == 0).all()
        # --- oracle: the instance enters Accepting
        status.reshape(G, S)[:, slot] = oracle.ST_ACCEPTING
        acks.reshape(G, S)[:, slot] = 0
        inst_bal.reshape(G, S)[:, slot] = bal
        # --- replies: for every proposed-or-future slot, random subsets of replicas, plus noise
        nrec = 40000
        rg = rng.integers(0, G, nrec).astype(np.uint32)
        cand = np.array([0, 1, 2, 3, 63, 5, 40], dtype=np.uint8)       # 5 and 40 are never proposed
        rs_ = cand[rng.integers(0, len(cand), nrec)]
        rp = rng.integers(0, population + 1, nrec).astype(np.uint8)    # population = out-of-range peer
        rb = bal[rg].copy()
        stale = rng.random(nrec) < 0.08
        rb[stale] += rng.integers(1, 3, int(stale.sum())).astype(np.uint64)
        for half in range(2):
            sl = slice(half * nrec // 2, (half + 1) * nrec // 2)
            eng.on_accept_replies(_t(rg[sl]), torch.from_numpy(rs_[sl]).to(DEV), torch.from_numpy(rp[sl]).to(DEV), _t(rb[sl]))
            committed, bar = eng.tick()
            torch.cuda.synchronize()
            oracle.tally_stream(rg[sl], rs_[sl], rp[sl], rb[sl], S, population, thr, bal, inst_bal, status, acks)
            want = np.packbits(status.reshape(G, S) == oracle.ST_COMMITTED, axis=1, bitorder="little").view(np.uint64).reshape(-1)
            assert (committed.cpu().numpy().view(np.uint64) == want).all(), (rnd, half)
            want_bar = np.array([oracle.commit_bar(int(w)) for w in want], dtype=np.uint32)
            assert (bar.cpu().numpy().view(np.uint32) == want_bar).all()
    # the stream did commit things, and not everything
    frac = (status == oracle.ST_COMMITTED).sum() / (G * n_rounds)
    assert 0.2 < frac <= 1.0
