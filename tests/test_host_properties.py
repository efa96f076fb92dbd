"""Property tests (hypothesis) of the host-side pieces that need no GPU: the bincode / safetcp encoders of
summerset_b200/wire.py, the Bitmap mirror (src/utils/bitmap.rs), the codeword geometry helpers, and the oracle's
Reed-Solomon restatement (encode -> erase <= p shards -> reconstruct is the identity; linearity of the code)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from summerset_b200 import wire
from summerset_b200.api import Bitmap, crossword_brr_assignment, round_up, shard_len

U64 = st.integers(min_value=0, max_value=(1 << 64) - 1)


@given(U64)
def test_varint_roundtrip_and_width(v):
    b = wire.varint(v)
    assert wire.read_varint(b + b"\xaa\xbb", 0) == (v, len(b))
    # bincode 2 "standard" varint widths: 1 / 3 / 5 / 9 bytes
    assert len(b) == (1 if v < 251 else 3 if v < (1 << 16) else 5 if v < (1 << 32) else 9)


@given(st.lists(U64, min_size=1, max_size=8))
def test_varint_stream(vs):
    blob = b"".join(wire.varint(v) for v in vs)
    pos, out = 0, []
    for _ in vs:
        v, pos = wire.read_varint(blob, pos)
        out.append(v)
    assert out == vs and pos == len(blob)


@given(st.integers(1, 6), st.integers(1, 4), st.integers(1, 70000), st.data())
@settings(max_examples=60, deadline=None)
def test_rscodeword_encode_decode(d, p, data_len, data):
    L = shard_len(data_len, d)
    present = data.draw(st.lists(st.booleans(), min_size=d + p, max_size=d + p))
    rng = np.random.default_rng(data_len * 31 + d)
    shards = [rng.integers(0, 256, L, dtype=np.uint8).tobytes() if pr else None for pr in present]
    blob = wire.encode_rscodeword(d, p, data_len, L, shards)
    cw, pos = wire.decode_rscodeword(blob + b"tail", 0)
    assert pos == len(blob)
    assert (cw["d"], cw["p"], cw["data_len"], cw["shard_len"], cw["has_copy"]) == (d, p, data_len, L, False)
    assert cw["shards"] == shards
    # an Accept frame is the 8-byte big-endian length of a body that ends with the codeword
    idx = data.draw(st.integers(0, d + p - 1))
    body_shard = rng.integers(0, 256, L, dtype=np.uint8).tobytes()
    f = wire.rspaxos_accept_frame(5, 77, d, p, data_len, idx, body_shard)
    assert int.from_bytes(f[:8], "big") == len(f) - 8
    assert f[8] == wire.PEER_MESSAGE_MSG and f[9] == wire.RSPAXOS_ACCEPT


@given(st.integers(1, 255), st.data())
def test_bitmap_mirror(size, data):
    ones = data.draw(st.sets(st.integers(0, size - 1)))
    b = Bitmap.from_indices(size, ones)
    assert b.size() == size and b.count() == len(ones)
    assert all(b.get(i) == (i in ones) for i in range(size))
    b.flip()
    assert b.count() == size - len(ones)
    b.flip()
    assert [i for i in range(size) if b.get(i)] == sorted(ones)


@given(st.integers(0, 1 << 20), st.integers(1, 32))
def test_geometry_helpers(data_len, d):
    L = shard_len(data_len, d)
    assert L * d >= data_len and (L == 0 or (L - 1) * d < data_len)
    assert round_up(L, 16) % 16 == 0 and 0 <= round_up(L, 16) - L < 16


@given(st.integers(1, 12), st.integers(1, 3), st.integers(1, 12))
def test_brr_assignment_shape(n, mult, spr):
    T = n * mult
    spr = min(spr, T)
    asg = crossword_brr_assignment(n, T, spr)
    assert len(asg) == n and all(bin(m).count("1") == spr for m in asg)
    # replica r's window starts at r * (T / n)
    assert all((m >> ((r * mult) % T)) & 1 for r, m in enumerate(asg))


@given(st.sampled_from([(2, 1), (3, 2), (4, 3), (5, 4), (6, 4), (3, 1), (4, 2)]), st.integers(1, 600), st.data())
@settings(max_examples=40, deadline=None)
def test_oracle_roundtrip_and_linearity(oracle, code, data_len, data):
    d, p = code
    rng = np.random.default_rng(data_len + 1000 * d)
    a = rng.integers(0, 256, (1, round_up(data_len, 16) + 16), dtype=np.uint8)
    b = rng.integers(0, 256, a.shape, dtype=np.uint8)
    pa = oracle.rs_encode_uniform(d, p, a, data_len)
    pb = oracle.rs_encode_uniform(d, p, b, data_len)
    pab = oracle.rs_encode_uniform(d, p, a ^ b, data_len)
    assert (pab == (pa ^ pb)).all()                              # the code is GF(2)-linear
    L = shard_len(data_len, d)
    sh = oracle.cw_split(a[0, :data_len].tobytes(), d)
    full = [sh[i].copy() for i in range(d)] + [pa[j, 0, :L].copy() for j in range(p)]
    nmiss = data.draw(st.integers(0, p))
    miss = data.draw(st.lists(st.integers(0, d + p - 1), min_size=nmiss, max_size=nmiss, unique=True))
    cw = [None if i in miss else full[i].copy() for i in range(d + p)]
    assert oracle.rs_reconstruct(d, p, cw, False) == 0
    assert all((x == y).all() for x, y in zip(cw, full))
