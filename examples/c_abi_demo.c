/* c_abi_demo.c -- the C ABI (include/summerset_b200.h) used from plain C, the way a Rust `extern "C"` shim
 * would drive it: no torch, no C++ -- pointers and sizes only.
 *
 * One tick of a batched RSPaxos leader for G groups:
 *   host payloads -> device arena                         (ss_copy_h2d)
 *   RS(3,2)-encode every group's request batch + tally the acks that arrived   (ss_accept_step_fused_dev)
 *   commit_bar[] back to the host                         (ss_copy_d2h)
 * then a follower-side reconstruct of a codeword that lost two shards (ss_rs_reconstruct, host slices).
 *
 * build: gcc -O2 examples/c_abi_demo.c -Iinclude -Lsummerset_b200 -lsummerset_b200 -Wl,-rpath,$PWD/summerset_b200 -o c_abi_demo
 * Without an sm_100 GPU it prints the library's refusal (there is no CPU fallback) and exits 3.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "summerset_b200.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != SS_OK) {                                                          \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, ss_last_error());          \
            return rc_ == SS_ERR_NO_DEVICE ? 3 : 1;                                  \
        }                                                                            \
    } while (0)

int main(void) {
    enum { G = 4096, LEN = 4096, D = 3, P = 2, R = 5 };
    const uint32_t L = (LEN + D - 1) / D, DS = (L + 15) / 16 * 16;     /* shard length 1366, padded slot 1376 */
    ss_ctx *ctx = NULL;
    ss_rs_coder *rs = NULL;
    CHECK(ss_ctx_create(0, &ctx));
    CHECK(ss_rs_coder_create(ctx, D, P, &rs));

    /* host side: pinned payloads and ack planes (leader always acks, followers ack 7 of 8 slots) */
    uint8_t *h_data = NULL; uint64_t *h_planes = NULL; uint32_t *h_bar = NULL;
    CHECK(ss_host_alloc(ctx, (size_t)G * LEN, (void **)&h_data));
    CHECK(ss_host_alloc(ctx, (size_t)R * G * 8, (void **)&h_planes));
    CHECK(ss_host_alloc(ctx, (size_t)G * 4, (void **)&h_bar));
    for (size_t i = 0; i < (size_t)G * LEN; ++i) h_data[i] = (uint8_t)(i * 2654435761u >> 13);
    for (int r = 0; r < R; ++r)
        for (int g = 0; g < G; ++g) h_planes[(size_t)r * G + g] = r == 0 ? ~0ull : ~(0x0101010101010101ull << ((g + r) & 7));

    /* device side */
    uint8_t *d_data = NULL, *d_parity = NULL; uint64_t *d_planes = NULL, *d_committed = NULL; uint32_t *d_bar = NULL;
    CHECK(ss_dev_alloc(ctx, (size_t)G * LEN, (void **)&d_data));
    CHECK(ss_dev_alloc(ctx, (size_t)P * G * DS, (void **)&d_parity));
    CHECK(ss_dev_alloc(ctx, (size_t)R * G * 8, (void **)&d_planes));
    CHECK(ss_dev_alloc(ctx, (size_t)G * 8, (void **)&d_committed));
    CHECK(ss_dev_alloc(ctx, (size_t)G * 4, (void **)&d_bar));
    CHECK(ss_copy_h2d(ctx, d_data, h_data, (size_t)G * LEN));
    CHECK(ss_copy_h2d(ctx, d_planes, h_planes, (size_t)R * G * 8));
    /* threshold = majority + fault_tolerance = 3 + 1 (rspaxos/messages.rs:438-440) */
    CHECK(ss_accept_step_fused_dev(rs, d_data, LEN, LEN, G, d_parity, (uint64_t)G * DS, DS, SS_RS_OUT_PADDED16, d_planes, R, 4,
                                   d_committed, d_bar));
    CHECK(ss_copy_d2h(ctx, h_bar, d_bar, (size_t)G * 4));
    CHECK(ss_ctx_sync(ctx));
    unsigned long long slots = 0;
    for (int g = 0; g < G; ++g) slots += h_bar[g];
    printf("fused accept step: %d groups, %llu slots in committed prefixes, kernels launched: %llu\n", G, slots,
           (unsigned long long)ss_ctx_launch_count(ctx));

    /* follower side, one codeword through the crate-shaped calls: encode, lose shards 0 and 3, reconstruct */
    uint8_t shard[D + P][1366], keep0[1366];
    uint8_t *ptr[D + P];
    uint8_t present[D + P] = {1, 1, 1, 1, 1};
    for (int j = 0; j < D + P; ++j) ptr[j] = shard[j];
    memset(shard, 0, sizeof(shard));
    memcpy(shard[0], h_data, L); memcpy(shard[1], h_data + L, L); memcpy(shard[2], h_data + 2 * L, LEN - 2 * L);
    CHECK(ss_rs_encode(rs, ptr, D + P, L));
    int ok = 0;
    CHECK(ss_rs_verify(rs, (const uint8_t *const *)ptr, D + P, L, &ok));
    memcpy(keep0, shard[0], L);
    memset(shard[0], 0xEE, L); memset(shard[3], 0xEE, L);
    present[0] = present[3] = 0;
    CHECK(ss_rs_reconstruct(rs, ptr, present, D + P, L));
    printf("verify after encode: %s; reconstruct of shards {0,3}: %s\n", ok ? "ok" : "MISMATCH",
           memcmp(keep0, shard[0], L) == 0 && present[0] && present[3] ? "bit-exact" : "MISMATCH");
    /* too few shards: refused, nothing written */
    present[0] = present[1] = present[4] = 0;
    int rc = ss_rs_reconstruct(rs, ptr, present, D + P, L);
    printf("3 of 5 missing -> %d (%s)\n", rc, ss_strerror(rc));

    ss_dev_free(ctx, d_data); ss_dev_free(ctx, d_parity); ss_dev_free(ctx, d_planes); ss_dev_free(ctx, d_committed); ss_dev_free(ctx, d_bar);
    ss_host_free(ctx, h_data); ss_host_free(ctx, h_planes); ss_host_free(ctx, h_bar);
    ss_rs_coder_destroy(rs);
    ss_ctx_destroy(ctx);
    return (ok && rc == SS_ERR_TOO_FEW_SHARDS_PRESENT) ? 0 : 1;
}
