// static_codes.hpp -- compile-time parity rows of the Reed-Solomon codes Summerset's protocols build.
// Shared by the CUDA kernels (rs_kernels.cu) and a host-only test (tests/cpp/test_static_codes.cpp) that pins every
// table entry against gf256.hpp's coding matrix, i.e. against the construction the reed-solomon-erasure crate uses.
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#endif

#if defined(__CUDACC__)
#define SSB_HD __host__ __device__
#else
#define SSB_HD
#endif

namespace ssb {

// The codes Summerset's protocols actually build are ReedSolomon::new(majority, population - majority)
// (rspaxos/mod.rs:597-609, crossword/mod.rs:742-750): population 3 -> RS(2,1), 5 -> RS(3,2), 7 -> RS(4,3), 9 -> RS(5,4)
// (4 -> RS(3,1), 6 -> RS(4,2)).  For those the parity rows are compile-time constants, so the Horner evaluation is
// fully unrolled and only the set coefficient bits cost an instruction.  The coder only selects a static code when its
// run-time matrix (built by gf256.hpp exactly as the crate builds it) equals the table below byte for byte.
enum : int { kCodeGeneric = -1, kCode21 = 0, kCode43 = 1, kCode54 = 2, kCode42 = 3, kCode31 = 4, kNumStaticCodes = 5,
             kCodeJit = 100 /* the code NVRTC specialises at run time: SS_JIT_D x SS_JIT_P coefficients SS_JIT_COEFS */ };
#if defined(SS_JIT_D)
SSB_HD constexpr uint32_t jit_coef(int idx) {
    constexpr uint32_t c[] = {SS_JIT_COEFS};
    return c[idx];
}
#else
SSB_HD constexpr uint32_t jit_coef(int) { return 0u; }
#define SS_JIT_D 0
#define SS_JIT_P 0
#endif
SSB_HD constexpr int static_code_d(int code) {
    if (code == kCodeJit) return SS_JIT_D;
    return code == kCode21 ? 2 : code == kCode43 ? 4 : code == kCode54 ? 5 : code == kCode42 ? 4 : code == kCode31 ? 3 : 0;
}
SSB_HD constexpr int static_code_p(int code) {
    if (code == kCodeJit) return SS_JIT_P;
    return code == kCode21 ? 1 : code == kCode43 ? 3 : code == kCode54 ? 4 : code == kCode42 ? 2 : code == kCode31 ? 1 : 0;
}
SSB_HD constexpr uint32_t static_code_coef(int code, int j, int i) {
    if (code == kCodeJit) return jit_coef(j * SS_JIT_D + i);
    if (code == kCode21) return i == 0 ? 0x03u : 0x02u;
    if (code == kCode31) return 0x01u;
    if (code == kCode43 || code == kCode42) {
        switch (j * 4 + i) {
            case 0: return 0x1bu; case 1: return 0x1cu; case 2: return 0x12u; case 3: return 0x14u;
            case 4: return 0x1cu; case 5: return 0x1bu; case 6: return 0x14u; case 7: return 0x12u;
            case 8: return 0x12u; case 9: return 0x14u; case 10: return 0x1bu; default: return 0x1cu;
        }
    }
    switch (j * 5 + i) {   // kCode54
        case 0: return 0x07u; case 1: return 0x07u; case 2: return 0x06u; case 3: return 0x06u; case 4: return 0x01u;
        case 5: return 0x09u; case 6: return 0x08u; case 7: return 0x09u; case 8: return 0x08u; case 9: return 0x01u;
        case 10: return 0x0fu; case 11: return 0x0eu; case 12: return 0x0eu; case 13: return 0x0fu; case 14: return 0x01u;
        case 15: return 0x02u; case 16: return 0x7du; case 17: return 0x95u; case 18: return 0xfdu; default: return 0x16u;
    }
}
SSB_HD constexpr int static_code_top(int code, int j) {
    uint32_t any = 0;
    for (int i = 0; i < static_code_d(code); ++i) any |= static_code_coef(code, j, i);
    int top = 0;
    for (int k = 0; k < 8; ++k)
        if ((any >> k) & 1u) top = k;
    return top;
}

}  // namespace ssb
