//! verify_with_cargo.rs -- turns the committed vectors of tests/golden/rs_vectors.json into a REAL pin against the
//! reference (josehu07/summerset) and the crates it depends on.  Nothing in this repository can run it (no Rust toolchain
//! in the build image); a maintainer with `cargo` does, once:
//!
//!   1. copy this file to <summerset>/src/protocols/rspaxos/verify_b200_golden.rs and rs_vectors.json next to it;
//!   2. add `#[cfg(test)] mod verify_b200_golden;` at the end of <summerset>/src/protocols/rspaxos/mod.rs
//!      (the module must live inside `rspaxos` because `PeerMsg` and `WalEntry` are private to it);
//!   3. add `serde_json = "1"` and `hex = "0.4"` under [dev-dependencies];
//!   4. `cargo test -p summerset verify_b200_golden`.
//!
//! What a green run pins (every vector below was produced by oracle/ss_oracle.c + oracle/ss_wire.c, i.e. by this
//! repository's CPU restatement, NOT by the reference; "reqbatch_put" starts from a real ReqBatch whose bincode bytes
//! tests/golden/make_golden.py writes out by hand):
//!   * parity bytes of reed_solomon_erasure::galois_8::ReedSolomon for RS(3,2), (4,3), (5,4), (2,1), (6,4)
//!     -> DESIGN.md section 4 "parity unpinned" becomes "pinned";
//!   * RSCodeword::from_data's split (shard_len, zero padding) -- src/utils/rscoding.rs:165-243;
//!   * bincode 2 `config::standard()` bytes of Bitmap (src/utils/bitmap.rs:20-30), of
//!     PeerMessage::Msg{PeerMsg::Accept / AcceptReply} frames (src/server/transport.rs:37-40, rspaxos/mod.rs:283-291)
//!     and of WalEntry::{AcceptData, CommitSlot} (rspaxos/mod.rs:212-232), each behind the 8-byte big-endian length
//!     prefix of safetcp / StorageHub.
//! The Crossword Accept vector (assignment: Vec<Bitmap>) needs the same three lines inside crossword/mod.rs.

use super::*;
use crate::server::transport::PeerMessage;          // make `PeerMessage` pub(crate) for the test, or move the frame checks there
use crate::utils::{Bitmap, RSCodeword};
use reed_solomon_erasure::galois_8::ReedSolomon;

fn vectors() -> serde_json::Value {
    serde_json::from_str(include_str!("rs_vectors.json")).unwrap()
}

fn framed<T: bincode::Encode>(v: &T) -> Vec<u8> {
    let body = bincode::encode_to_vec(v, bincode::config::standard()).unwrap();
    let mut out = (body.len() as u64).to_be_bytes().to_vec(); // safetcp.rs / storage.rs: write_u64 (big-endian) + body
    out.extend_from_slice(&body);
    out
}

#[test]
fn parity_bytes_match_the_crate() {
    for case in vectors()["rs"].as_array().unwrap() {
        let (d, p) = (case["d"].as_u64().unwrap() as usize, case["p"].as_u64().unwrap() as usize);
        let payload = hex::decode(case["payload"].as_str().unwrap()).unwrap();
        let rs = ReedSolomon::new(d, p).unwrap();
        // RSCodeword::from_data's geometry (rscoding.rs:177-199), on raw bytes
        let data_len = payload.len();
        let shard_len = if data_len % d == 0 { data_len / d } else { data_len / d + 1 };
        let mut padded = payload.clone();
        padded.resize(shard_len * d, 0);
        let mut shards: Vec<Vec<u8>> = padded.chunks(shard_len).map(|c| c.to_vec()).collect();
        for (i, want) in case["data_shards"].as_array().unwrap().iter().enumerate() {
            assert_eq!(hex::encode(&shards[i]), want.as_str().unwrap(), "data shard {} of RS({},{})", i, d, p);
        }
        shards.extend((0..p).map(|_| vec![0u8; shard_len]));
        rs.encode(&mut shards).unwrap();
        for (j, want) in case["parity_shards"].as_array().unwrap().iter().enumerate() {
            assert_eq!(hex::encode(&shards[d + j]), want.as_str().unwrap(), "parity {} of RS({},{}) len {}", j, d, p, data_len);
        }
    }
}

#[test]
fn bitmap_bincode_matches() {
    for case in vectors()["bitmap"].as_array().unwrap() {
        let size = case["size"].as_u64().unwrap() as u8;
        let ones: Vec<u8> = case["ones"].as_array().unwrap().iter().map(|v| v.as_u64().unwrap() as u8).collect();
        let bm = Bitmap::from((size, ones));
        let bytes = bincode::encode_to_vec(&bm, bincode::config::standard()).unwrap();
        assert_eq!(hex::encode(bytes), case["bincode"].as_str().unwrap(), "Bitmap of size {}", size);
    }
}

#[test]
fn frames_match() {
    let v = vectors();
    let f = &v["frames"];
    let reply = PeerMessage::Msg { msg: PeerMsg::AcceptReply { slot: 300, ballot: 70000 } };
    assert_eq!(hex::encode(framed(&reply)), f["rspaxos_accept_reply_slot300_ballot70000"].as_str().unwrap());
    let commit = WalEntry::CommitSlot { slot: 300 };
    assert_eq!(hex::encode(framed(&commit)), f["rspaxos_wal_commit_slot300"].as_str().unwrap());
}

/// A real request batch end to end: bincode of the batch, RSCodeword::from_data's split, compute_parity's bytes, the
/// single-shard subset_copy a follower is sent (rspaxos/request.rs:127-142), its Accept frame and its WAL record.
#[test]
fn real_request_batch_matches() -> Result<(), SummersetError> {
    use crate::server::{ApiRequest, Command};
    let v = vectors();
    let g = &v["reqbatch_put"];
    let batch: ReqBatch = vec![(7, ApiRequest::Req { id: 1, cmd: Command::Put { key: "k1".into(), value: "value-0123456789abcdef".into() } })];
    let bytes = bincode::encode_to_vec(&batch, bincode::config::standard()).unwrap();
    assert_eq!(hex::encode(&bytes), g["bincode"].as_str().unwrap(), "bincode of the batch");

    let rs = ReedSolomon::new(3, 2).unwrap();
    let mut cw = RSCodeword::<ReqBatch>::from_data(batch, 3, 2)?;
    assert_eq!(cw.data_len() as u64, g["data_len"].as_u64().unwrap());
    assert_eq!(cw.shard_len() as u64, g["shard_len"].as_u64().unwrap());
    cw.compute_parity(Some(&rs))?;
    // every shard, through the single-shard copies the leader sends (the only public view of the shard bytes)
    for j in 0..5u8 {
        let one = cw.subset_copy(&Bitmap::from((5, vec![j])), false)?;
        let enc = bincode::encode_to_vec(&one, bincode::config::standard()).unwrap();
        let want = hex::decode(g["shards"][j as usize].as_str().unwrap()).unwrap();
        assert!(enc.windows(want.len()).any(|w| w == &want[..]), "shard {} bytes inside the encoded subset copy", j);
    }
    let sub = cw.subset_copy(&Bitmap::from((5, vec![1])), false)?;
    let accept = PeerMessage::Msg { msg: PeerMsg::Accept { slot: 300, ballot: 70000, reqs_cw: sub.clone() } };
    assert_eq!(hex::encode(framed(&accept)), g["accept_frame_shard1_slot300_ballot70000"].as_str().unwrap());
    let wal = WalEntry::AcceptData { slot: 300, ballot: 70000, reqs_cw: sub };
    assert_eq!(hex::encode(framed(&wal)), g["wal_accept_data_shard1_slot300_ballot70000"].as_str().unwrap());
    // the Crossword vector (g["crossword_accept_frame_shards12_..."]) is the same check inside crossword/mod.rs with
    // PeerMsg::Accept { slot, ballot, reqs_cw: cw.subset_copy(&Bitmap::from((5, vec![1, 2])), false)?,
    //                   assignment: vec![Bitmap::from((5, vec![0,1])), (5,[1,2]), (5,[2,3]), (5,[3,4]), (5,[4,0])] }
    Ok(())
}
