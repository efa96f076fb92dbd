// Host-only check (no GPU, no CUDA): every compile-time parity row in summerset_b200/csrc/static_codes.hpp equals the
// coding matrix gf256.hpp builds the way the reed-solomon-erasure crate does (Vandermonde * inverse of its top square),
// for exactly the codes the protocols construct: ReedSolomon::new(majority, population - majority)
// (rspaxos/mod.rs:597-609).  The kernels only select a static code after the same comparison at coder creation; this
// test makes a table typo a CPU-suite failure instead of a silent fall-back to the run-time-mask kernels.
#include <cstdio>

#include "../../summerset_b200/csrc/gf256.hpp"
#include "../../summerset_b200/csrc/static_codes.hpp"

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
    std::printf("static codes: %d coefficients checked, %d failure(s)\n", checked, failures);
    return failures ? 1 : 0;
}
