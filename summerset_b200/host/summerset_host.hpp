// summerset_host.hpp -- header-only C++17 host side above the C ABI (include/summerset_b200.h).
//
// The reference's host code for this path is Rust; no Rust toolchain exists in the build image, so the
// host-side mirror is C++ with the reference's names, argument meaning and error behaviour:
//   ssb::Bitmap        <- src/utils/bitmap.rs            (u8 id -> bool bitset)
//   ssb::ReedSolomon   <- reed_solomon_erasure::galois_8::ReedSolomon as used by src/utils/rscoding.rs
//   ssb::RSCodeword    <- src/utils/rscoding.rs          (T = an already serialised byte string)
// Only bookkeeping happens here.  Every shard byte is computed by the CUDA kernels behind the C ABI;
// a failing ABI call becomes a SummersetError (src/utils/error.rs:6-14) -- there is no CPU fallback.
#pragma once
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/summerset_b200.h"

namespace ssb {

struct SummersetError : std::runtime_error {
    int code;
    SummersetError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
    static SummersetError msg(const std::string &m) { return SummersetError(SS_ERR_INVALID_ARG, m); }
};

inline void check(int rc) {
    if (rc != SS_OK) throw SummersetError(rc, ss_last_error());
}

// ---------------------------------------------------------------------------------------------
// Bitmap (src/utils/bitmap.rs:17-149)
// ---------------------------------------------------------------------------------------------
class Bitmap {
  public:
    Bitmap(uint8_t size, bool ones) : size_(size), words_((size + 63u) / 64u, 0) {
        if (size == 0) throw std::logic_error("invalid bitmap size 0");   // bitmap.rs:64 (panic)
        if (ones)
            for (unsigned i = 0; i < size; ++i) words_[i / 64] |= 1ull << (i % 64);
    }
    static Bitmap from(uint8_t size, const std::vector<uint8_t> &ones) {       // bitmap.rs:151-185
        Bitmap b(size, false);
        for (uint8_t i : ones) b.set(i, true);
        return b;
    }
    void set(uint8_t idx, bool flag) {                                         // bitmap.rs:74-84
        if (idx >= size_) throw SummersetError(SS_ERR_INVALID_INDEX, "index " + std::to_string(idx) + " out of bound");
        if (flag) words_[idx / 64] |= 1ull << (idx % 64);
        else words_[idx / 64] &= ~(1ull << (idx % 64));
    }
    bool get(uint8_t idx) const {                                              // bitmap.rs:87-97
        if (idx >= size_) throw SummersetError(SS_ERR_INVALID_INDEX, "index " + std::to_string(idx) + " out of bound");
        return (words_[idx / 64] >> (idx % 64)) & 1ull;
    }
    uint8_t size() const { return size_; }
    uint8_t count() const {                                                    // bitmap.rs:111-113
        unsigned c = 0;
        for (uint64_t w : words_) c += static_cast<unsigned>(__builtin_popcountll(w));
        return static_cast<uint8_t>(c);
    }
    void flip() {                                                              // bitmap.rs:117-119
        for (unsigned i = 0; i < size_; ++i) words_[i / 64] ^= 1ull << (i % 64);
    }
    void union_with(const Bitmap &o) {                                         // bitmap.rs:123-135
        if (size_ != o.size_)
            throw SummersetError::msg("unioning sizes mismatch: " + std::to_string(size_) + " != " + std::to_string(o.size_));
        for (size_t i = 0; i < words_.size(); ++i) words_[i] |= o.words_[i];
    }
    void clear() { std::fill(words_.begin(), words_.end(), 0ull); }
    std::vector<uint8_t> to_vec() const {                                      // bitmap.rs:211-229
        std::vector<uint8_t> v;
        for (unsigned i = 0; i < size_; ++i)
            if (get(static_cast<uint8_t>(i))) v.push_back(static_cast<uint8_t>(i));
        return v;
    }
    uint64_t mask() const { return words_[0]; }                                // ids < 64: the device-side vote mask
    bool operator==(const Bitmap &o) const { return size_ == o.size_ && words_ == o.words_; }

  private:
    uint8_t size_;
    std::vector<uint64_t> words_;
};

// ---------------------------------------------------------------------------------------------
// ReedSolomon: RAII over ss_ctx + ss_rs_coder
// ---------------------------------------------------------------------------------------------
class ReedSolomon {
  public:
    ReedSolomon(size_t data_shards, size_t parity_shards, int device = 0) {
        check(ss_ctx_create(device, &ctx_));
        int rc = ss_rs_coder_create(ctx_, static_cast<int>(data_shards), static_cast<int>(parity_shards), &coder_);
        if (rc != SS_OK) {
            std::string m = ss_last_error();
            ss_ctx_destroy(ctx_);
            throw SummersetError(rc, m);
        }
    }
    ~ReedSolomon() {
        ss_rs_coder_destroy(coder_);
        ss_ctx_destroy(ctx_);
    }
    ReedSolomon(const ReedSolomon &) = delete;
    ReedSolomon &operator=(const ReedSolomon &) = delete;

    size_t data_shard_count() const { return static_cast<size_t>(ss_rs_data_shard_count(coder_)); }
    size_t parity_shard_count() const { return static_cast<size_t>(ss_rs_parity_shard_count(coder_)); }
    size_t total_shard_count() const { return static_cast<size_t>(ss_rs_total_shard_count(coder_)); }

    // rs.encode(slices): all d+p shards, parity overwritten
    void encode(std::vector<std::vector<uint8_t>> &shards) const {
        const size_t len = shards.empty() ? 0 : shards[0].size();
        for (auto &s : shards)
            if (s.size() != len) throw SummersetError(SS_ERR_INCORRECT_SHARD_SIZE, "incorrect shard size");
        std::vector<uint8_t *> p;
        for (auto &s : shards) p.push_back(s.data());
        check(ss_rs_encode(coder_, p.data(), p.size(), len));
    }
    void reconstruct(std::vector<std::optional<std::vector<uint8_t>>> &shards) const { recon(shards, false); }
    void reconstruct_data(std::vector<std::optional<std::vector<uint8_t>>> &shards) const { recon(shards, true); }
    bool verify(const std::vector<std::vector<uint8_t>> &shards) const {
        const size_t len = shards.empty() ? 0 : shards[0].size();
        std::vector<const uint8_t *> p;
        for (auto &s : shards) p.push_back(s.data());
        int ok = 0;
        check(ss_rs_verify(coder_, p.data(), p.size(), len, &ok));
        return ok != 0;
    }
    ss_rs_coder *handle() const { return coder_; }
    ss_ctx *context() const { return ctx_; }

  private:
    void recon(std::vector<std::optional<std::vector<uint8_t>>> &shards, bool data_only) const {
        size_t len = 0;
        for (auto &s : shards)
            if (s) { len = s->size(); break; }
        std::vector<uint8_t> present(shards.size());
        std::vector<std::vector<uint8_t>> tmp(shards.size());
        std::vector<uint8_t *> p(shards.size());
        for (size_t i = 0; i < shards.size(); ++i) {
            present[i] = shards[i] ? 1 : 0;
            if (shards[i]) p[i] = shards[i]->data();
            else { tmp[i].assign(len ? len : 1, 0); p[i] = tmp[i].data(); }
        }
        check(data_only ? ss_rs_reconstruct_data(coder_, p.data(), present.data(), p.size(), len)
                        : ss_rs_reconstruct(coder_, p.data(), present.data(), p.size(), len));
        for (size_t i = 0; i < shards.size(); ++i)
            if (!shards[i] && present[i]) shards[i] = std::move(tmp[i]);
    }
    ss_ctx *ctx_ = nullptr;
    ss_rs_coder *coder_ = nullptr;
};

// ---------------------------------------------------------------------------------------------
// RSCodeword (src/utils/rscoding.rs:15-606), T = serialized bytes
// ---------------------------------------------------------------------------------------------
class RSCodeword {
  public:
    using Bytes = std::vector<uint8_t>;

    static RSCodeword from_data(const Bytes &data, uint8_t d, uint8_t p) {    // rscoding.rs:223-243
        return internal_new(data, &data, data.size(), d, p);
    }
    static RSCodeword from_null(uint8_t d, uint8_t p) {                       // rscoding.rs:246-251
        return internal_new(std::nullopt, nullptr, 0, d, p);
    }

    RSCodeword subset_copy(const Bitmap &subset, bool copy_data) const {      // rscoding.rs:255-293
        if (data_len_ == 0) throw SummersetError::msg("codeword is null");
        RSCodeword out(*this);
        for (auto &s : out.shards) s.reset();
        for (unsigned i = 0; i < subset.size(); ++i) {
            if (!subset.get(static_cast<uint8_t>(i))) continue;
            if (i >= shards.size()) throw SummersetError::msg("shard index " + std::to_string(i) + " out-of-bound");
            out.shards[i] = shards[i];
        }
        if (!copy_data) out.data_copy.reset();
        return out;
    }

    void absorb_other(RSCodeword other) {                                     // rscoding.rs:296-345
        if (d_ != other.d_) throw SummersetError::msg("num_data_shards mismatch");
        if (p_ != other.p_) throw SummersetError::msg("num_parity_shards mismatch");
        if (data_len_ != 0 && data_len_ != other.data_len_) throw SummersetError::msg("data_len mismatch");
        if (shard_len_ != 0 && shard_len_ != other.shard_len_) throw SummersetError::msg("shard_len mismatch");
        if (data_len_ == 0) { data_len_ = other.data_len_; shard_len_ = other.shard_len_; }
        for (size_t i = 0; i < other.shards.size(); ++i)
            if (other.shards[i] && !shards[i]) shards[i] = std::move(other.shards[i]);
    }

    uint8_t num_data_shards() const { return d_; }
    uint8_t num_parity_shards() const { return p_; }
    uint8_t num_shards() const { return static_cast<uint8_t>(shards.size()); }
    uint8_t avail_data_shards() const { return count(0, d_); }
    uint8_t avail_parity_shards() const { return count(d_, shards.size()); }
    uint8_t avail_shards() const { return count(0, shards.size()); }
    Bitmap avail_shards_map() const {                                         // rscoding.rs:400-408
        Bitmap m(num_shards(), false);
        for (size_t i = 0; i < shards.size(); ++i)
            if (shards[i]) m.set(static_cast<uint8_t>(i), true);
        return m;
    }
    size_t data_len() const { return data_len_; }
    size_t shard_len() const { return shard_len_; }

    void compute_parity(const ReedSolomon *rs) {                              // rscoding.rs:447-486
        if (data_len_ == 0) throw SummersetError::msg("codeword is null");
        if (p_ == 0) return;
        if (!rs) throw SummersetError::msg("ReedSolomon coder is None");
        splits_match(*rs);
        if (avail_data_shards() < d_) throw SummersetError::msg("not all data shards present");
        std::vector<Bytes> all;
        for (size_t i = 0; i < shards.size(); ++i) all.push_back(shards[i] ? *shards[i] : Bytes(shard_len_, 0));
        rs->encode(all);
        for (size_t i = 0; i < shards.size(); ++i) shards[i] = std::move(all[i]);
    }
    void reconstruct_all(const ReedSolomon *rs) { reconstruct(rs, false); }   // rscoding.rs:524-529
    void reconstruct_data(const ReedSolomon *rs) { reconstruct(rs, true); }   // rscoding.rs:532-537

    bool verify_parity(const ReedSolomon *rs) {                               // rscoding.rs:542-576
        if (data_len_ == 0) throw SummersetError::msg("codeword is null");
        if (p_ == 0) {
            if (avail_data_shards() == d_) return true;
            throw SummersetError::msg("not all shards present");
        }
        if (!rs) throw SummersetError::msg("ReedSolomon is None");
        splits_match(*rs);
        if (avail_shards() < num_shards()) throw SummersetError::msg("not all shards present");
        std::vector<Bytes> all;
        for (auto &s : shards) all.push_back(*s);
        return rs->verify(all);
    }

    const Bytes &get_data() {                                                 // rscoding.rs:581-606
        if (data_len_ == 0) throw SummersetError::msg("codeword is null");
        if (avail_data_shards() < d_) throw SummersetError::msg("not all data shards present");
        if (!data_copy) {                                                      // ShardsReader, rscoding.rs:649-682
            Bytes cat;
            for (unsigned i = 0; i < d_; ++i) cat.insert(cat.end(), shards[i]->begin(), shards[i]->end());
            cat.resize(data_len_);
            data_copy = std::move(cat);
        }
        return *data_copy;
    }

    std::vector<std::optional<Bytes>> shards;
    std::optional<Bytes> data_copy;

  private:
    static RSCodeword internal_new(std::optional<Bytes> copy, const Bytes *bytes, size_t data_len, uint8_t d,
                                   uint8_t p) {                               // rscoding.rs:165-220
        if (d == 0) throw SummersetError::msg("num_data_shards is zero");
        RSCodeword cw;
        cw.d_ = d; cw.p_ = p; cw.data_len_ = data_len;
        cw.shard_len_ = (data_len % d == 0) ? data_len / d : data_len / d + 1;
        cw.shards.resize(size_t(d) + p);
        if (bytes) {
            Bytes padded(*bytes);
            padded.resize(cw.shard_len_ * d, 0);
            for (unsigned i = 0; i < d; ++i)
                cw.shards[i] = Bytes(padded.begin() + long(i * cw.shard_len_), padded.begin() + long((i + 1) * cw.shard_len_));
        }
        cw.data_copy = std::move(copy);
        return cw;
    }
    uint8_t count(size_t a, size_t b) const {
        unsigned c = 0;
        for (size_t i = a; i < b; ++i) c += shards[i] ? 1u : 0u;
        return static_cast<uint8_t>(c);
    }
    void splits_match(const ReedSolomon &rs) const {                          // rscoding.rs:424-443
        if (rs.data_shard_count() != d_) throw SummersetError::msg("num_data_shards mismatch");
        if (rs.parity_shard_count() != p_) throw SummersetError::msg("num_parity_shards mismatch");
    }
    void reconstruct(const ReedSolomon *rs, bool data_only) {                 // rscoding.rs:490-520
        if (data_len_ == 0) throw SummersetError::msg("codeword is null");
        if (p_ == 0) {
            if (avail_data_shards() == d_) return;
            throw SummersetError::msg("insufficient data shards");
        }
        if (!rs) throw SummersetError::msg("ReedSolomon coder is None");
        splits_match(*rs);
        if (data_only) rs->reconstruct_data(shards);
        else rs->reconstruct(shards);
    }
    uint8_t d_ = 0, p_ = 0;
    size_t data_len_ = 0, shard_len_ = 0;
};

}  // namespace ssb
