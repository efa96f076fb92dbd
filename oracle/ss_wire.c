/* ss_wire.c -- CPU restatement of the byte formats on either side of the hot path.
 * TEST INFRASTRUCTURE ONLY, like the rest of oracle/: the product's GPU packers / decoders are checked against it.
 *
 * What the reference puts on a TCP connection (src/utils/safetcp.rs:30-88): an 8-byte big-endian body length, then
 * bincode 2 `config::standard()` of PeerMessage::Msg { msg: PeerMsg } (src/server/transport.rs:33-40).  The WAL uses the
 * same 8-byte big-endian length prefix (src/server/storage.rs:300-305,333-337: write_u64 + encode_to_vec).
 *
 * bincode 2 standard config (from knowledge of the crate -- the reference cannot run here, so these bytes are UNPINNED
 * against it; tests/golden/verify_with_cargo.rs is the recipe that pins them on a box with Rust):
 *   unsigned ints: < 251 one byte; else tag 251/252/253 + u16/u32/u64 little-endian;  u8: one raw byte
 *   enum variant: varint u32;  Vec / slice: varint length + items;  Option: 0 / 1 tag;  tuple / struct: fields in order
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#include "ss_oracle.h"

size_t ssor_varint_put(uint8_t *out, uint64_t v) {
    if (v < 251) { out[0] = (uint8_t)v; return 1; }
    int nb; uint8_t tag;
    if (v < (1ull << 16)) { nb = 2; tag = 251; }
    else if (v < (1ull << 32)) { nb = 4; tag = 252; }
    else { nb = 8; tag = 253; }
    out[0] = tag;
    for (int i = 0; i < nb; i++) out[1 + i] = (uint8_t)(v >> (8 * i));
    return (size_t)(1 + nb);
}

/* returns bytes consumed, or 0 when truncated / not a u64-sized varint */
size_t ssor_varint_get(const uint8_t *in, size_t avail, uint64_t *v) {
    if (avail < 1) return 0;
    uint8_t t = in[0];
    if (t < 251) { *v = t; return 1; }
    size_t nb = t == 251 ? 2 : t == 252 ? 4 : t == 253 ? 8 : 0;
    if (nb == 0 || avail < 1 + nb) return 0;
    uint64_t x = 0;
    for (size_t i = 0; i < nb; i++) x |= (uint64_t)in[1 + i] << (8 * i);
    *v = x;
    return 1 + nb;
}

static size_t put_be64(uint8_t *out, uint64_t v) {
    for (int i = 0; i < 8; i++) out[i] = (uint8_t)(v >> (8 * (7 - i)));
    return 8;
}

/* src/utils/bitmap.rs:20-30: logical bit length, then the backing usize block slice (one block for size <= 64) */
size_t ssor_bitmap_encode(uint32_t size, uint64_t bits, uint8_t *out) {
    size_t n = ssor_varint_put(out, size);
    uint32_t nblocks = (size + 63) / 64;
    n += ssor_varint_put(out + n, nblocks);
    for (uint32_t b = 0; b < nblocks; b++) n += ssor_varint_put(out + n, b == 0 ? bits : 0);
    return n;
}

/* src/utils/rscoding.rs:43-72 (hand-written Encode): num_data_shards u8, num_parity_shards u8, data_len, shard_len,
 * shards Vec<Option<BytesMut>>, data_copy Option<T> (None after subset_copy(.., false), rspaxos/request.rs:127-142).
 * shards[j] == NULL encodes None. */
size_t ssor_rscodeword_encode(uint32_t d, uint32_t p, uint64_t data_len, uint64_t shard_len,
                              const uint8_t *const *shards, uint8_t *out) {
    size_t n = 0;
    out[n++] = (uint8_t)d;
    out[n++] = (uint8_t)p;
    n += ssor_varint_put(out + n, data_len);
    n += ssor_varint_put(out + n, shard_len);
    n += ssor_varint_put(out + n, d + p);
    for (uint32_t j = 0; j < d + p; j++) {
        if (shards[j] == NULL) { out[n++] = 0; continue; }
        out[n++] = 1;
        n += ssor_varint_put(out + n, shard_len);
        memcpy(out + n, shards[j], shard_len);
        n += shard_len;
    }
    out[n++] = 0;                                   /* data_copy: None */
    return n;
}

/* PeerMessage::Msg { msg: PeerMsg::Accept { slot, ballot, reqs_cw [, assignment] } } framed for the wire.
 * accept_variant: index of Accept in the protocol's PeerMsg enum (2 in rspaxos/mod.rs:249-309 and crossword/mod.rs:324-400).
 * assignment (Crossword only, crossword/mod.rs:356-362): n_assign Bitmaps of size assign_size; NULL = field absent. */
size_t ssor_frame_accept(uint32_t accept_variant, uint64_t slot, uint64_t ballot, uint32_t d, uint32_t p,
                         uint64_t data_len, uint64_t shard_len, const uint8_t *const *shards,
                         const uint32_t *assignment, uint32_t n_assign, uint32_t assign_size, uint8_t *out) {
    uint8_t *body = out + 8;
    size_t n = 0;
    n += ssor_varint_put(body + n, 0);                               /* PeerMessage::Msg (transport.rs:37-40) */
    n += ssor_varint_put(body + n, accept_variant);
    n += ssor_varint_put(body + n, slot);
    n += ssor_varint_put(body + n, ballot);
    n += ssor_rscodeword_encode(d, p, data_len, shard_len, shards, body + n);
    if (assignment != NULL) {
        n += ssor_varint_put(body + n, n_assign);
        for (uint32_t r = 0; r < n_assign; r++) n += ssor_bitmap_encode(assign_size, assignment[r], body + n);
    }
    put_be64(out, n);                                                /* safetcp.rs:30-88 */
    return 8 + n;
}

/* PeerMsg::AcceptReply { slot, ballot } (rspaxos, variant 3) or { slot, ballot, size, reply_ts: None } (crossword) */
size_t ssor_frame_accept_reply(uint32_t reply_variant, uint64_t slot, uint64_t ballot, int with_size, uint64_t size,
                               uint8_t *out) {
    uint8_t *body = out + 8;
    size_t n = 0;
    n += ssor_varint_put(body + n, 0);
    n += ssor_varint_put(body + n, reply_variant);
    n += ssor_varint_put(body + n, slot);
    n += ssor_varint_put(body + n, ballot);
    if (with_size) {
        n += ssor_varint_put(body + n, size);
        body[n++] = 0;                                               /* reply_ts: None */
    }
    put_be64(out, n);
    return 8 + n;
}

/* Parses one frame.  Returns the frame's total length (8 + body) and fills the fields when it is a well-formed
 * AcceptReply of this protocol; -2 when it is some other message (kind = variant index, host handles it); -1 when
 * malformed or truncated. */
long ssor_parse_accept_reply(const uint8_t *frame, size_t avail, uint32_t reply_variant, int with_size,
                             uint64_t *slot, uint64_t *ballot, uint64_t *size, uint32_t *kind) {
    if (avail < 8) return -1;
    uint64_t len = 0;
    for (int i = 0; i < 8; i++) len = (len << 8) | frame[i];
    if (len > avail - 8) return -1;
    const uint8_t *b = frame + 8;
    size_t n = 0, c;
    uint64_t v;
    if (!(c = ssor_varint_get(b + n, len - n, &v))) return -1;
    n += c;
    if (v != 0) { *kind = 0x80000000u | (uint32_t)v; return -2; }    /* LeaseMsg / Leave / LeaveReply */
    if (!(c = ssor_varint_get(b + n, len - n, &v))) return -1;
    n += c;
    *kind = (uint32_t)v;
    if (v != reply_variant) return -2;
    if (!(c = ssor_varint_get(b + n, len - n, slot))) return -1;
    n += c;
    if (!(c = ssor_varint_get(b + n, len - n, ballot))) return -1;
    n += c;
    *size = 0;
    if (with_size) {
        if (!(c = ssor_varint_get(b + n, len - n, size))) return -1;
        n += c;
        if (n >= len) return -1;
        uint8_t tag = b[n++];
        if (tag == 1) {                                              /* Some(SystemTime): u64 secs + u32 nanos since the epoch */
            if (!(c = ssor_varint_get(b + n, len - n, &v))) return -1;
            n += c;
            if (!(c = ssor_varint_get(b + n, len - n, &v))) return -1;
            n += c;
        } else if (tag != 0) return -1;
    }
    if (n != len) return -1;
    return (long)(8 + len);
}

/* WalEntry::AcceptData { slot, ballot, reqs_cw } (variant 1) / WalEntry::CommitSlot { slot } (variant 2)
 * (rspaxos/mod.rs:212-232), with the StorageHub length prefix (storage.rs:333-337) */
size_t ssor_wal_accept_data(uint64_t slot, uint64_t ballot, uint32_t d, uint32_t p, uint64_t data_len,
                            uint64_t shard_len, const uint8_t *const *shards, uint8_t *out) {
    uint8_t *body = out + 8;
    size_t n = 0;
    n += ssor_varint_put(body + n, 1);
    n += ssor_varint_put(body + n, slot);
    n += ssor_varint_put(body + n, ballot);
    n += ssor_rscodeword_encode(d, p, data_len, shard_len, shards, body + n);
    put_be64(out, n);
    return 8 + n;
}

size_t ssor_wal_commit_slot(uint64_t slot, uint8_t *out) {
    uint8_t *body = out + 8;
    size_t n = 0;
    n += ssor_varint_put(body + n, 2);
    n += ssor_varint_put(body + n, slot);
    put_be64(out, n);
    return 8 + n;
}

/* Decoder of what ssor_frame_accept / ssor_wal_accept_data wrote (round-trip pin: decode(encode(x)) == x).
 * kind: 0 = peer Accept frame, 1 = WAL AcceptData.  shard_at[j] = offset of shard j's bytes inside `frame`, or 0 for
 * None.  Returns total length or -1. */
long ssor_decode_accept(const uint8_t *frame, size_t avail, int kind, uint32_t *variant, uint64_t *slot, uint64_t *ballot,
                        uint32_t *d, uint32_t *p, uint64_t *data_len, uint64_t *shard_len, uint64_t *shard_at,
                        uint32_t max_shards, int with_assignment, uint32_t *assignment, uint32_t *n_assign,
                        uint32_t *assign_size) {
    if (avail < 8) return -1;
    uint64_t len = 0;
    for (int i = 0; i < 8; i++) len = (len << 8) | frame[i];
    if (len > avail - 8) return -1;
    const uint8_t *b = frame + 8;
    size_t n = 0, c;
    uint64_t v;
#define GETV(dst) do { if (!(c = ssor_varint_get(b + n, len - n, (dst)))) return -1; n += c; } while (0)
    if (kind == 0) { GETV(&v); if (v != 0) return -1; }
    GETV(&v); *variant = (uint32_t)v;
    GETV(slot); GETV(ballot);
    if (len - n < 2) return -1;
    *d = b[n++]; *p = b[n++];
    GETV(data_len); GETV(shard_len);
    GETV(&v);
    if (v != (uint64_t)(*d + *p) || v > max_shards) return -1;
    for (uint32_t j = 0; j < *d + *p; j++) {
        if (n >= len) return -1;
        uint8_t tag = b[n++];
        if (tag == 0) { shard_at[j] = 0; continue; }
        if (tag != 1) return -1;
        GETV(&v);
        if (v != *shard_len || len - n < v) return -1;
        shard_at[j] = 8 + n;
        n += v;
    }
    if (n >= len || b[n++] != 0) return -1;                           /* data_copy must be None */
    if (with_assignment) {
        GETV(&v); *n_assign = (uint32_t)v;
        for (uint32_t r = 0; r < *n_assign; r++) {
            uint64_t sz, nb, blk = 0;
            GETV(&sz); GETV(&nb);
            for (uint64_t k = 0; k < nb; k++) { uint64_t w; GETV(&w); if (k == 0) blk = w; }
            *assign_size = (uint32_t)sz;
            assignment[r] = (uint32_t)blk;
        }
    }
#undef GETV
    if (n != len) return -1;
    return (long)(8 + len);
}

/* Reconstruct serving (crossword/messages.rs:577-632, rspaxos/messages.rs:468-517): for one requested slot, the shards
 * this replica sends back = held & flip(exclude) when the instance is at least Accepting; 0 = no entry in the reply
 * (status below Accepting, or nothing left after the exclusion: `avail_shards() == 0 -> continue`). */
uint32_t ssor_reconstruct_serve_mask(uint32_t held, uint32_t exclude, uint32_t total_shards, int status) {
    if (status < SSOR_ST_ACCEPTING) return 0;
    uint32_t all = total_shards >= 32 ? 0xffffffffu : ((1u << total_shards) - 1u);
    return held & ~exclude & all;                                    /* subset.flip(); subset_copy(&subset, false) */
}
