// Prints what the PRODUCT's host-side field code (summerset_b200/csrc/gf256.hpp, the tables the GPU coefficient programs are
// built from) computes, so a Python test can compare it with an independent implementation:
//   gf_dump mul a b ...      -> products of consecutive pairs
//   gf_dump matrix d p       -> the (d+p) x d coding matrix, one row per line (hex)
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../summerset_b200/csrc/gf256.hpp"

int main(int argc, char **argv) {
    if (argc >= 2 && !strcmp(argv[1], "mul")) {
        for (int i = 2; i + 1 < argc; i += 2)
            std::printf("%u\n", ssb::gf::mul(static_cast<uint8_t>(atoi(argv[i])), static_cast<uint8_t>(atoi(argv[i + 1]))));
        return 0;
    }
    if (argc == 4 && !strcmp(argv[1], "matrix")) {
        const int d = atoi(argv[2]), p = atoi(argv[3]);
        const auto M = ssb::gf::coding_matrix(d, p);
        for (int r = 0; r < d + p; ++r) {
            for (int c = 0; c < d; ++c) std::printf("%02x ", M.at(r, c));
            std::printf("\n");
        }
        return 0;
    }
    return 2;
}
