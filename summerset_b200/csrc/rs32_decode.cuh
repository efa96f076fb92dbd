// rs32_decode.cuh -- compile-time decode rows of RS(3,2), the code of every 5-replica RSPaxos / Crossword / CRaft
// deployment (ReedSolomon::new(3, 2), rspaxos/mod.rs:597-609).
//
// reconstruct / reconstruct_data (rscoding.rs:490-537 -> crate ReedSolomon::reconstruct{,_data}) picks the FIRST d = 3
// present shards in index order, inverts the 3x3 sub-matrix of their rows and multiplies: the result depends only on
// the 5-bit present mask, so all of it -- source choice, inverse, output rows -- is evaluated by the compiler
// (constexpr GF(2^8) arithmetic, poly 0x11D) and each pattern's column routine is a fully unrolled packed Horner
// evaluation in which only SET coefficient bits cost an instruction.  Outputs that the all-ones parity row
// (d0 ^ d1 ^ d2 ^ p0 = 0) determines from values already at hand are computed as a 2-XOR chain instead of a dense row.
// tests/cpp/test_static_codes.cpp pins every table against gf256.hpp's run-time inverse.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SSB_HD __host__ __device__
#else
#define SSB_HD
#endif

namespace ssb {
namespace rs32 {

SSB_HD constexpr uint32_t cmul(uint32_t a, uint32_t b) {
    uint32_t r = 0;
    for (int i = 0; i < 8; ++i)
        if ((b >> i) & 1u) r ^= a << i;
    for (int i = 14; i >= 8; --i)
        if ((r >> i) & 1u) r ^= 0x11Du << (i - 8);
    return r & 0xffu;
}
SSB_HD constexpr uint32_t cinv(uint32_t a) {      // a^254
    uint32_t r = 1;
    for (int i = 0; i < 254; ++i) r = cmul(r, a);
    return r;
}
// coding matrix rows of RS(3,2): identity, {1,1,1}, {0f,08,06} (= vandermonde(5,3) * inverse(top), as the crate builds it)
SSB_HD constexpr uint32_t mrow(int shard, int col) {
    return shard < 3 ? (shard == col ? 1u : 0u) : shard == 3 ? 1u : (col == 0 ? 0x0fu : col == 1 ? 0x08u : 0x06u);
}

struct Decode {
    bool valid = false;          // >= 3 shards present
    int src[3] = {0, 0, 0};      // first three present shards, index order
    int n_out = 0;               // shards to regenerate: missing data first, then (full mode) missing parity
    int dst[2] = {0, 0};
    uint32_t coef[2][3] = {{0, 0, 0}, {0, 0, 0}};   // out_j = sum_i coef[j][i] * src_i
    // output 1 as an XOR chain: out_1 = val[a] ^ val[b] ^ val[c], val = {src0, src1, src2, out0}
    bool chain = false;
    int chain_term[3] = {0, 0, 0};
};

SSB_HD constexpr Decode make_decode(uint32_t present, bool data_only) {
    Decode D{};
    present &= 31u;
    int ns = 0;
    for (int i = 0; i < 5 && ns < 3; ++i)
        if ((present >> i) & 1u) D.src[ns++] = i;
    if (ns < 3) return D;
    D.valid = true;
    // sub-matrix of the source rows and its inverse (adjugate / determinant; characteristic 2: no signs)
    uint32_t a[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) a[r][c] = mrow(D.src[r], c);
    const uint32_t det = cmul(a[0][0], cmul(a[1][1], a[2][2]) ^ cmul(a[1][2], a[2][1])) ^
                         cmul(a[0][1], cmul(a[1][0], a[2][2]) ^ cmul(a[1][2], a[2][0])) ^
                         cmul(a[0][2], cmul(a[1][0], a[2][1]) ^ cmul(a[1][1], a[2][0]));
    const uint32_t di = cinv(det);       // det != 0: every 3x3 sub-matrix of an MDS code's generator is invertible
    uint32_t inv[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            // cofactor of a[c][r] (transposed): rows != c, cols != r
            const int r0 = c == 0 ? 1 : 0, r1 = c == 2 ? 1 : 2, c0 = r == 0 ? 1 : 0, c1 = r == 2 ? 1 : 2;
            inv[r][c] = cmul(di, cmul(a[r0][c0], a[r1][c1]) ^ cmul(a[r0][c1], a[r1][c0]));
        }
    // data shard k = sum_i inv[k][i] * src_i
    for (int k = 0; k < 3; ++k) {
        if ((present >> k) & 1u) continue;
        D.dst[D.n_out] = k;
        for (int i = 0; i < 3; ++i) D.coef[D.n_out][i] = inv[k][i];
        ++D.n_out;
    }
    if (!data_only) {
        for (int q = 3; q < 5; ++q) {
            if ((present >> q) & 1u) continue;
            D.dst[D.n_out] = q;
            for (int i = 0; i < 3; ++i) {
                uint32_t c = 0;
                for (int k = 0; k < 3; ++k) c ^= cmul(mrow(q, k), inv[k][i]);
                D.coef[D.n_out][i] = c;
            }
            ++D.n_out;
        }
    }
    // XOR chain for output 1: d0 ^ d1 ^ d2 ^ p0 = 0.  Applies when out_1 is one of {d0,d1,d2,p0} and the other three
    // are among the sources and out_0.
    if (D.n_out == 2 && D.dst[1] <= 3) {
        int terms = 0;
        bool ok = true;
        for (int s = 0; s <= 3 && ok; ++s) {
            if (s == D.dst[1]) continue;
            int at = -1;
            for (int i = 0; i < 3; ++i)
                if (D.src[i] == s) at = i;
            if (D.dst[0] == s) at = 3;
            if (at < 0) ok = false;
            else D.chain_term[terms++] = at;
        }
        D.chain = ok && terms == 3;
    }
    return D;
}

// work a pattern needs: 0 = nothing to regenerate, otherwise its outputs
SSB_HD constexpr bool needs_work(uint32_t present, bool data_only) {
    return data_only ? (present & 7u) != 7u : (present & 31u) != 31u;
}

#if defined(__CUDACC__)
// multiply four packed field elements by x
__device__ __forceinline__ uint32_t xt(uint32_t x) { return ((x * 2u) & 0xfefefefeu) ^ (ssb::dev::msb_mask(x) & 0x1d1d1d1du); }

// out = sum_i C_i * x_i, coefficients compile-time: Horner in x from the top set bit down
template <uint32_t C0, uint32_t C1, uint32_t C2>
__device__ __forceinline__ uint4 row3(const uint4 &x0, const uint4 &x1, const uint4 &x2) {
    constexpr uint32_t any = C0 | C1 | C2;
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
    bool started = false;
#pragma unroll
    for (int k = 7; k >= 0; --k) {
        if (!((any >> k) & 1u) && !started) continue;       // levels above the top set bit
        if (started) { acc.x = xt(acc.x); acc.y = xt(acc.y); acc.z = xt(acc.z); acc.w = xt(acc.w); }
        started = true;
        if ((C0 >> k) & 1u) { acc.x ^= x0.x; acc.y ^= x0.y; acc.z ^= x0.z; acc.w ^= x0.w; }
        if ((C1 >> k) & 1u) { acc.x ^= x1.x; acc.y ^= x1.y; acc.z ^= x1.z; acc.w ^= x1.w; }
        if ((C2 >> k) & 1u) { acc.x ^= x2.x; acc.y ^= x2.y; acc.z ^= x2.z; acc.w ^= x2.w; }
    }
    return acc;
}

// the outputs of one 16-byte column for a compile-time pattern
template <uint32_t PRESENT, bool DATA_ONLY>
__device__ __forceinline__ void decode_column(const uint4 &x0, const uint4 &x1, const uint4 &x2, uint4 &y0, uint4 &y1) {
    constexpr Decode D = make_decode(PRESENT, DATA_ONLY);
    y0 = row3<D.coef[0][0], D.coef[0][1], D.coef[0][2]>(x0, x1, x2);
    if constexpr (D.n_out == 2) {
        if constexpr (D.chain) {
            const uint4 v[4] = {x0, x1, x2, y0};
            const uint4 &a = v[D.chain_term[0]], &b = v[D.chain_term[1]], &c = v[D.chain_term[2]];
            y1 = make_uint4(a.x ^ b.x ^ c.x, a.y ^ b.y ^ c.y, a.z ^ b.z ^ c.z, a.w ^ b.w ^ c.w);
        } else {
            y1 = row3<D.coef[1][0], D.coef[1][1], D.coef[1][2]>(x0, x1, x2);
        }
    }
}
#endif

}  // namespace rs32
}  // namespace ssb
