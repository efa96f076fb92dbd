"""Wire / WAL byte formats on either side of the hot path.  Host-side reference encoders.

What the reference puts on a TCP connection (src/utils/safetcp.rs:30-159): an 8-byte big-endian body length, then
the bincode 2 `config::standard()` encoding of `PeerMessage::Msg { msg: PeerMsg }` (src/server/transport.rs:37-52).
For RSPaxos (src/protocols/rspaxos/mod.rs:249-309) `PeerMsg::Accept { slot, ballot, reqs_cw }` is variant 2 and
`reqs_cw: RSCodeword<ReqBatch>` uses the hand-written Encode impl of src/utils/rscoding.rs:43-72:
    num_data_shards u8, num_parity_shards u8, data_len usize, shard_len usize,
    shards Vec<Option<Vec<u8>>>, data_copy Option<T>
WAL entries (rspaxos/mod.rs:212-232): PrepareBal = 0, AcceptData { slot, ballot, reqs_cw } = 1, CommitSlot { slot } = 2.

bincode 2 standard config facts used (little-endian varint ints: < 251 one byte, else tag 251/252/253 + u16/u32/u64;
enum variant = varint u32; Vec / String = varint length + items; Option = 0 / 1 tag; u8 = one raw byte) are from
knowledge of the crate -- the reference cannot be run here, so these bytes are UNPINNED against it.
The GPU frame packer (ss_frame_accept_batch_dev) is tested byte-for-byte against these encoders.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


def varint(v: int) -> bytes:
    if v < 0:
        raise ValueError("unsigned only")
    if v < 251:
        return bytes([v])
    if v < 1 << 16:
        return b"\xfb" + v.to_bytes(2, "little")
    if v < 1 << 32:
        return b"\xfc" + v.to_bytes(4, "little")
    if v < 1 << 64:
        return b"\xfd" + v.to_bytes(8, "little")
    return b"\xfe" + v.to_bytes(16, "little")


def read_varint(b: bytes, pos: int) -> Tuple[int, int]:
    t = b[pos]
    if t < 251:
        return t, pos + 1
    n = {251: 2, 252: 4, 253: 8, 254: 16}[t]
    return int.from_bytes(b[pos + 1:pos + 1 + n], "little"), pos + 1 + n


def encode_rscodeword(d: int, p: int, data_len: int, shard_len: int, shards: Sequence[Optional[bytes]],
                      data_copy: Optional[bytes] = None) -> bytes:
    """rscoding.rs:54-71; data_copy is the already-bincode-encoded T (None after subset_copy(.., false))."""
    out = bytearray([d, p]) + varint(data_len) + varint(shard_len) + varint(len(shards))
    for s in shards:
        if s is None:
            out += b"\x00"
        else:
            out += b"\x01" + varint(len(s)) + bytes(s)
    out += b"\x00" if data_copy is None else b"\x01" + data_copy
    return bytes(out)


def decode_rscodeword(b: bytes, pos: int = 0):
    d, p = b[pos], b[pos + 1]
    data_len, pos = read_varint(b, pos + 2)
    shard_len, pos = read_varint(b, pos)
    n, pos = read_varint(b, pos)
    shards: List[Optional[bytes]] = []
    for _ in range(n):
        tag = b[pos]; pos += 1
        if tag == 0:
            shards.append(None)
        else:
            ln, pos = read_varint(b, pos)
            shards.append(b[pos:pos + ln]); pos += ln
    has_copy = b[pos]; pos += 1
    return dict(d=d, p=p, data_len=data_len, shard_len=shard_len, shards=shards, has_copy=bool(has_copy)), pos


PEER_MESSAGE_MSG = 0            # transport.rs:37-40
RSPAXOS_ACCEPT, RSPAXOS_ACCEPT_REPLY = 2, 3        # rspaxos/mod.rs:249-309
WAL_PREPARE_BAL, WAL_ACCEPT_DATA, WAL_COMMIT_SLOT = 0, 1, 2   # rspaxos/mod.rs:212-232


def frame(body: bytes) -> bytes:
    """safetcp.rs: 8-byte big-endian length + body"""
    return len(body).to_bytes(8, "big") + body


def rspaxos_accept_body(slot: int, ballot: int, cw: bytes) -> bytes:
    return varint(PEER_MESSAGE_MSG) + varint(RSPAXOS_ACCEPT) + varint(slot) + varint(ballot) + cw


def rspaxos_accept_frame(slot: int, ballot: int, d: int, p: int, data_len: int, shard_idx: int, shard: bytes) -> bytes:
    """The frame replica `shard_idx` receives: its single shard (rspaxos/request.rs:127-142), data_copy None."""
    shards: List[Optional[bytes]] = [None] * (d + p)
    shards[shard_idx] = shard
    return frame(rspaxos_accept_body(slot, ballot, encode_rscodeword(d, p, data_len, len(shard), shards)))


def rspaxos_accept_reply_frame(slot: int, ballot: int) -> bytes:
    return frame(varint(PEER_MESSAGE_MSG) + varint(RSPAXOS_ACCEPT_REPLY) + varint(slot) + varint(ballot))


def wal_accept_data(slot: int, ballot: int, cw: bytes) -> bytes:
    return varint(WAL_ACCEPT_DATA) + varint(slot) + varint(ballot) + cw


def wal_commit_slot(slot: int) -> bytes:
    return varint(WAL_COMMIT_SLOT) + varint(slot)
