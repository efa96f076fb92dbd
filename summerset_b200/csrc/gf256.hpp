// gf256.hpp -- host-side GF(2^8) arithmetic and coding-matrix construction for the product.
//
// Field and matrix are those of the crate the reference calls
// (`reed_solomon_erasure::galois_8::ReedSolomon`, src/utils/rscoding.rs:9; Cargo.toml:43):
// generating polynomial 0x11D, generator 2; M = vandermonde(d+p, d) * inverse(top d x d).
// Only tiny matrices are handled here (setup time); the byte loops run on the GPU.
// This file is NOT the test oracle (oracle/ is a separate C restatement); the two are
// compared against each other in tests/.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace ssb {
namespace gf {

struct Tables {
    std::array<uint8_t, 512> exp{};
    std::array<uint8_t, 256> log{};
    Tables() {
        unsigned v = 1;
        for (unsigned i = 0; i < 255; ++i) {
            exp[i] = static_cast<uint8_t>(v);
            log[v] = static_cast<uint8_t>(i);
            v = (v << 1) ^ ((v & 0x80u) ? 0x11Du : 0u);
        }
        for (unsigned i = 255; i < 512; ++i) exp[i] = exp[i - 255];
    }
};

inline const Tables &tables() {
    static const Tables t;
    return t;
}

inline uint8_t mul(uint8_t a, uint8_t b) {
    if (!a || !b) return 0;
    const Tables &t = tables();
    return t.exp[unsigned(t.log[a]) + unsigned(t.log[b])];
}

inline uint8_t inv(uint8_t a) {
    if (!a) throw std::domain_error("gf::inv(0)");
    const Tables &t = tables();
    return t.exp[255u - t.log[a]];
}

// a^n with the crate's conventions: a^0 = 1 (also for a = 0), 0^n = 0.
inline uint8_t pow(uint8_t a, unsigned n) {
    if (n == 0) return 1;
    if (a == 0) return 0;
    const Tables &t = tables();
    return t.exp[(unsigned(t.log[a]) * n) % 255u];
}

// Dense row-major matrix over GF(2^8).
struct Matrix {
    int rows = 0, cols = 0;
    std::vector<uint8_t> v;
    Matrix() = default;
    Matrix(int r, int c) : rows(r), cols(c), v(size_t(r) * size_t(c), 0) {}
    uint8_t &at(int r, int c) { return v[size_t(r) * size_t(cols) + size_t(c)]; }
    uint8_t at(int r, int c) const { return v[size_t(r) * size_t(cols) + size_t(c)]; }

    static Matrix identity(int n) {
        Matrix m(n, n);
        for (int i = 0; i < n; ++i) m.at(i, i) = 1;
        return m;
    }

    Matrix times(const Matrix &o) const {
        Matrix out(rows, o.cols);
        for (int r = 0; r < rows; ++r)
            for (int k = 0; k < cols; ++k) {
                const uint8_t a = at(r, k);
                if (!a) continue;
                for (int c = 0; c < o.cols; ++c) out.at(r, c) ^= mul(a, o.at(k, c));
            }
        return out;
    }

    // Gauss-Jordan inverse; returns false when singular.
    bool inverse(Matrix &out) const {
        const int n = rows;
        Matrix a = *this;
        out = identity(n);
        for (int col = 0; col < n; ++col) {
            int piv = col;
            while (piv < n && a.at(piv, col) == 0) ++piv;
            if (piv == n) return false;
            if (piv != col)
                for (int c = 0; c < n; ++c) {
                    std::swap(a.at(piv, c), a.at(col, c));
                    std::swap(out.at(piv, c), out.at(col, c));
                }
            const uint8_t s = inv(a.at(col, col));
            for (int c = 0; c < n; ++c) {
                a.at(col, c) = mul(a.at(col, c), s);
                out.at(col, c) = mul(out.at(col, c), s);
            }
            for (int r = 0; r < n; ++r) {
                if (r == col) continue;
                const uint8_t f = a.at(r, col);
                if (!f) continue;
                for (int c = 0; c < n; ++c) {
                    a.at(r, c) ^= mul(f, a.at(col, c));
                    out.at(r, c) ^= mul(f, out.at(col, c));
                }
            }
        }
        return true;
    }
};

// The (d+p) x d systematic coding matrix of ReedSolomon::new(d, p).
inline Matrix coding_matrix(int d, int p) {
    const int t = d + p;
    Matrix vm(t, d);
    for (int r = 0; r < t; ++r)
        for (int c = 0; c < d; ++c) vm.at(r, c) = pow(static_cast<uint8_t>(r), unsigned(c));
    Matrix top(d, d), top_inv;
    for (int r = 0; r < d; ++r)
        for (int c = 0; c < d; ++c) top.at(r, c) = vm.at(r, c);
    if (!top.inverse(top_inv)) throw std::runtime_error("vandermonde top block singular");
    return vm.times(top_inv);
}

}  // namespace gf
}  // namespace ssb
