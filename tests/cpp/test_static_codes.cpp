// Host-only check (no GPU, no CUDA): every compile-time parity row in summerset_b200/csrc/static_codes.hpp equals the
// coding matrix gf256.hpp builds the way the reed-solomon-erasure crate does (Vandermonde * inverse of its top square),
// for exactly the codes the protocols construct: ReedSolomon::new(majority, population - majority)
// (rspaxos/mod.rs:597-609).  The kernels only select a static code after the same comparison at coder creation; this
// test makes a table typo a CPU-suite failure instead of a silent fall-back to the run-time-mask kernels.
#include <cstdio>

#include "../../summerset_b200/csrc/gf256.hpp"
#include "../../summerset_b200/csrc/static_codes.hpp"
#include "../../summerset_b200/csrc/rs32_decode.cuh"

int main() {
    int failures = 0, checked = 0;
    for (int c = 0; c < ssb::kNumStaticCodes; ++c) {
        const int d = ssb::static_code_d(c), p = ssb::static_code_p(c);
        const auto M = ssb::gf::coding_matrix(d, p);
        for (int j = 0; j < p; ++j) {
            unsigned any = 0;
            for (int i = 0; i < d; ++i) {
                const unsigned want = M.at(d + j, i), got = ssb::static_code_coef(c, j, i);
                any |= want;
                ++checked;
                if (want != got) {
                    std::printf("code %d RS(%d,%d) row %d col %d: table %02x, matrix %02x\n", c, d, p, j, i, got, want);
                    ++failures;
                }
            }
            int top = 0;
            for (int k = 0; k < 8; ++k)
                if ((any >> k) & 1u) top = k;
            if (top != ssb::static_code_top(c, j)) {
                std::printf("code %d row %d: top %d, expected %d\n", c, j, ssb::static_code_top(c, j), top);
                ++failures;
            }
        }
    }
    // populations 3..9 map onto the table (5 -> RS(3,2) has its own hand-written kernel)
    const int pops[] = {3, 4, 6, 7, 9};
    for (int n : pops) {
        const int d = n / 2 + 1, p = n - d;
        bool found = false;
        for (int c = 0; c < ssb::kNumStaticCodes; ++c) found |= ssb::static_code_d(c) == d && ssb::static_code_p(c) == p;
        if (!found) { std::printf("population %d: RS(%d,%d) has no static code\n", n, d, p); ++failures; }
    }
    // RS(3,2) compile-time decode rows (rs32_decode.cuh) == run-time inverse of the first three present rows of the crate's
    // matrix, for every present mask, data-only and full; chains must reproduce the same shard through the parity relation
    {
        const auto M = ssb::gf::coding_matrix(3, 2);
        for (int c = 0; c < 3; ++c) {
            for (int r = 0; r < 5; ++r)
                if (ssb::rs32::mrow(r, c) != M.at(r, c)) { std::printf("rs32 mrow(%d,%d) differs from the coding matrix\n", r, c); ++failures; }
        }
        for (unsigned pat = 0; pat < 32; ++pat) {
            for (int mode = 0; mode < 2; ++mode) {
                const bool data_only = mode == 1;
                const ssb::rs32::Decode D = ssb::rs32::make_decode(pat, data_only);
                int src[3], ns = 0;
                for (int i = 0; i < 5 && ns < 3; ++i)
                    if ((pat >> i) & 1u) src[ns++] = i;
                if ((ns == 3) != D.valid) { std::printf("rs32 pattern %u: valid flag wrong\n", pat); ++failures; continue; }
                if (!D.valid) continue;
                ssb::gf::Matrix sub(3, 3), dec;
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) sub.at(r, c) = M.at(src[r], c);
                if (!sub.inverse(dec)) { std::printf("rs32 pattern %u: singular\n", pat); ++failures; continue; }
                int n_out = 0;
                auto check_row = [&](int shard, const unsigned (&want)[3]) {
                    if (n_out >= D.n_out || D.dst[n_out] != shard) { std::printf("rs32 pattern %u mode %d: output %d is not shard %d\n", pat, mode, n_out, shard); ++failures; }
                    else
                        for (int i = 0; i < 3; ++i) {
                            ++checked;
                            if (D.coef[n_out][i] != want[i]) { std::printf("rs32 pattern %u mode %d shard %d coef %d: %02x vs %02x\n", pat, mode, shard, i, D.coef[n_out][i], want[i]); ++failures; }
                        }
                    ++n_out;
                };
                for (int k = 0; k < 3; ++k) {
                    if ((pat >> k) & 1u) continue;
                    const unsigned want[3] = {dec.at(k, 0), dec.at(k, 1), dec.at(k, 2)};
                    check_row(k, want);
                }
                if (!data_only)
                    for (int q = 3; q < 5; ++q) {
                        if ((pat >> q) & 1u) continue;
                        unsigned want[3] = {0, 0, 0};
                        for (int i = 0; i < 3; ++i)
                            for (int k = 0; k < 3; ++k) want[i] ^= ssb::gf::mul(M.at(q, k), dec.at(k, i));
                        check_row(q, want);
                    }
                if (n_out != D.n_out) { std::printf("rs32 pattern %u mode %d: %d outputs, expected %d\n", pat, mode, D.n_out, n_out); ++failures; }
                if (D.chain) {
                    // out_1 as val[a]^val[b]^val[c] must have the same coefficients over the sources as the dense row
                    unsigned got[3] = {0, 0, 0};
                    for (int t = 0; t < 3; ++t) {
                        const int at = D.chain_term[t];
                        if (at < 3) got[at] ^= 1u;
                        else for (int i = 0; i < 3; ++i) got[i] ^= D.coef[0][i];
                    }
                    for (int i = 0; i < 3; ++i)
                        if (got[i] != D.coef[1][i]) { std::printf("rs32 pattern %u mode %d: chain disagrees with the dense row\n", pat, mode); ++failures; break; }
                }
            }
        }
    }
    std::printf("static codes: %d coefficients checked, %d failure(s)\n", checked, failures);
    return failures ? 1 : 0;
}
