// engine.cu -- the batched multi-group consensus engine behind the C ABI (ss_engine_*).
//
// One handle holds the leader-side state of G independent replica groups as a struct of arrays in HBM and advances
// all of them together; the host (the Rust shim of INTEGRATION.md, or summerset_b200/engine.py) calls
//
//   ss_engine_propose   RSCodeword::from_data + compute_parity + per-peer subset_copy for every group's new instance
//                       (multipaxos/request.rs:112-221, rspaxos/request.rs:72-142, crossword/request.rs:82-185): the
//                       RS encode kernel writes the shard planes, one bookkeeping kernel moves the instance to
//                       Accepting under the group's prepared ballot with an empty ack set
//   ss_engine_ingest    handle_msg_accept_reply's filters over a batch of AcceptReply records
//                       (multipaxos/messages.rs:377-409, rspaxos/messages.rs:402-437, crossword/messages.rs:489-530)
//   ss_engine_tick      the commit decision for every Accepting instance and the commit_bar advance, ONE kernel:
//                         MultiPaxos  accept_acks.count() >= quorum_cnt              (multipaxos/messages.rs:412-413)
//                         RSPaxos     count() >= majority + fault_tolerance          (rspaxos/messages.rs:438-440)
//                         Crossword   #acks >= majority && coverage_under_faults >= d (crossword/messages.rs:535-542)
//                       committed |= newly, accepting &= ~newly (a committed instance drops later replies, :394-399),
//                       commit_bar = committed prefix (multipaxos/durability.rs:161-170)
//   Raft / CRaft        ss_engine_raft_append (leader appends entries in its term), ss_engine_raft_ingest (successful
//                       AppendEntriesReply: next_slot / match_slot update, raft/messages.rs:243-252), ss_engine_tick
//                       (commit scan :254-275,295 and last_snap :298-309; CRaft thresholds craft/messages.rs:300-308)
//
// No tally, filter or shard byte is computed on the host, and nothing here is a torch op.
#include <cstring>
#include <new>

#include "device_common.cuh"
#include "ss_internal.hpp"

struct ss_engine {
    ss_ctx *ctx = nullptr;
    ss_rs_coder *coder = nullptr;
    ss_engine_config cfg{};
    uint64_t G = 0;
    uint32_t n = 0, majority = 0, threshold = 0, d = 0, T = 0;
    uint32_t L = 0, ds = 0;
    // Paxos family
    uint64_t *planes = nullptr, *bal_prepared = nullptr, *inst_bal = nullptr, *accepting = nullptr, *committed = nullptr;
    uint32_t *commit_bar = nullptr;
    uint8_t *shards = nullptr;
    // Crossword
    uint8_t *policy_idx = nullptr;          // [G*64]
    uint32_t *policies = nullptr, *lut_bits = nullptr;
    uint32_t n_policies = 0, lut_words = 0;
    int balanced = 1;
    // Raft
    uint32_t *match = nullptr, *next_slot = nullptr, *last_commit = nullptr, *log_end = nullptr, *curr_term = nullptr,
             *terms = nullptr, *last_snap = nullptr;
    uint8_t *touched = nullptr;
    uint32_t W = 0, n_peers = 0;
};

namespace ssb {

constexpr int kEngThreads = 256;

static inline uint32_t eng_grid(ss_ctx *ctx, uint64_t items) {
    uint64_t ctas = (items + kEngThreads - 1) / kEngThreads;
    const uint64_t cap = static_cast<uint64_t>(ctx->sm_count) * 8ull * 4ull;
    if (ctas > cap) ctas = cap;
    if (ctas == 0) ctas = 1;
    return static_cast<uint32_t>(ctas);
}

// ---- propose bookkeeping: the instance at `slot` of every group enters Accepting -------------------------------
__global__ void __launch_bounds__(kEngThreads)
engine_propose_kernel(uint64_t *__restrict__ planes, uint32_t R, uint64_t G, uint32_t slot, const uint64_t *__restrict__ bal_prepared,
                      uint64_t *__restrict__ inst_bal, uint64_t *__restrict__ accepting, uint64_t *__restrict__ committed,
                      uint8_t *__restrict__ policy_idx, const uint8_t *__restrict__ new_policy) {
    const uint64_t bit = 1ull << slot;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kEngThreads;
    for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * kEngThreads + threadIdx.x; g < G; g += stride) {
        accepting[g] |= bit;                                   // Status::Accepting (request.rs: inst.status = Accepting)
        committed[g] &= ~bit;                                  // a re-used window slot starts uncommitted
        inst_bal[g * 64 + slot] = bal_prepared[g];             // inst.bal = bal_prepared
        for (uint32_t r = 0; r < R; ++r) planes[static_cast<uint64_t>(r) * G + g] &= ~bit;   // fresh accept_acks
        if (policy_idx != nullptr) policy_idx[g * 64 + slot] = new_policy != nullptr ? new_policy[g] : 0;
    }
}

// ---- fused tick, threshold protocols: two groups per thread, 128-bit accesses ------------------------------------
template <int RT>
__global__ void __launch_bounds__(kEngThreads)
engine_tick_x2_kernel(const ulonglong2 *__restrict__ planes, uint32_t R, uint64_t G2, uint32_t threshold,
                      ulonglong2 *__restrict__ accepting, ulonglong2 *__restrict__ committed, uint2 *__restrict__ commit_bar,
                      ulonglong2 *__restrict__ newly_out) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kEngThreads;
    for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * kEngThreads + threadIdx.x; g < G2; g += stride) {
        uint64_t a[5] = {0, 0, 0, 0, 0}, b[5] = {0, 0, 0, 0, 0};
        auto add = [&](const ulonglong2 &x) {
            uint64_t c = x.x, t;
            t = a[0] & c; a[0] ^= c; c = t;
            t = a[1] & c; a[1] ^= c; c = t;
            t = a[2] & c; a[2] ^= c; c = t;
            t = a[3] & c; a[3] ^= c; c = t;
            a[4] ^= c;
            c = x.y;
            t = b[0] & c; b[0] ^= c; c = t;
            t = b[1] & c; b[1] ^= c; c = t;
            t = b[2] & c; b[2] ^= c; c = t;
            t = b[3] & c; b[3] ^= c; c = t;
            b[4] ^= c;
        };
        const ulonglong2 acc = accepting[g];
        ulonglong2 cm = committed[g];
        if constexpr (RT > 0) {
            ulonglong2 v[RT];
#pragma unroll
            for (int r = 0; r < RT; ++r) v[r] = __ldg(planes + static_cast<uint64_t>(r) * G2 + g);
#pragma unroll
            for (int r = 0; r < RT; ++r) add(v[r]);
        } else {
            for (uint32_t r = 0; r < R; ++r) add(__ldg(planes + static_cast<uint64_t>(r) * G2 + g));
        }
        auto ge = [&](const uint64_t (&cb)[5]) -> uint64_t {
            if (threshold == 0u) return ~0ull;
            if (threshold > 31u) return 0ull;
            uint64_t lt = 0ull, eq = ~0ull;
#pragma unroll
            for (int bb = 4; bb >= 0; --bb) {
                const uint64_t tb = ((threshold >> bb) & 1u) ? ~0ull : 0ull;
                lt |= eq & ~cb[bb] & tb;
                eq &= ~(cb[bb] ^ tb);
            }
            return ~lt;
        };
        ulonglong2 newly;
        newly.x = ge(a) & acc.x;                // only instances in Accepting can commit
        newly.y = ge(b) & acc.y;
        cm.x |= newly.x; cm.y |= newly.y;
        committed[g] = cm;
        accepting[g] = make_ulonglong2(acc.x & ~newly.x, acc.y & ~newly.y);
        commit_bar[g] = make_uint2(dev::commit_prefix(cm.x), dev::commit_prefix(cm.y));
        if (newly_out != nullptr) newly_out[g] = newly;
    }
}

__global__ void __launch_bounds__(kEngThreads)
engine_tick_kernel(const uint64_t *__restrict__ planes, uint32_t R, uint64_t G, uint32_t threshold, uint64_t *__restrict__ accepting,
                   uint64_t *__restrict__ committed, uint32_t *__restrict__ commit_bar, uint64_t *__restrict__ newly_out) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kEngThreads;
    for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * kEngThreads + threadIdx.x; g < G; g += stride) {
        const uint64_t acc = accepting[g];
        const uint64_t newly = dev::tally_word(planes, R, G, g, threshold) & acc;
        const uint64_t cm = committed[g] | newly;
        committed[g] = cm;
        accepting[g] = acc & ~newly;
        commit_bar[g] = dev::commit_prefix(cm);
        if (newly_out != nullptr) newly_out[g] = newly;
    }
}

// ---- fused tick, Crossword: coverage look-up per Accepting instance ------------------------------------------------
__global__ void __launch_bounds__(kEngThreads)
engine_tick_crossword_kernel(const uint64_t *__restrict__ planes, uint32_t R, uint64_t G, const uint8_t *__restrict__ policy_idx,
                             const uint32_t *__restrict__ lut_bits, uint32_t lut_words, uint32_t n_policies,
                             uint64_t *__restrict__ accepting, uint64_t *__restrict__ committed, uint32_t *__restrict__ commit_bar,
                             uint64_t *__restrict__ newly_out) {
    extern __shared__ uint32_t lut[];
    for (uint32_t i = threadIdx.x; i < lut_words; i += kEngThreads) lut[i] = lut_bits[i];
    __syncthreads();
    const uint32_t nmask = 1u << R;
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kEngThreads;
    for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * kEngThreads + threadIdx.x; g < G; g += stride) {
        const uint64_t acc = accepting[g];
        uint64_t newly = 0ull;
        if (acc != 0ull) {
            uint64_t pl[12];
#pragma unroll
            for (int r = 0; r < 12; ++r) pl[r] = static_cast<uint32_t>(r) < R ? __ldg(planes + static_cast<uint64_t>(r) * G + g) : 0ull;
            uint64_t rem = acc;
            while (rem) {
                const int s = __ffsll(static_cast<long long>(rem)) - 1;
                rem &= rem - 1ull;
                uint32_t ack = 0;
#pragma unroll
                for (int r = 0; r < 12; ++r) ack |= static_cast<uint32_t>((pl[r] >> s) & 1ull) << r;
                const uint32_t k = policy_idx[g * 64 + s];
                if (k < n_policies) {
                    const uint32_t e = k * nmask + ack;
                    if ((lut[e >> 5] >> (e & 31u)) & 1u) newly |= 1ull << s;
                }
            }
        }
        const uint64_t cm = committed[g] | newly;
        committed[g] = cm;
        accepting[g] = acc & ~newly;
        commit_bar[g] = dev::commit_prefix(cm);
        if (newly_out != nullptr) newly_out[g] = newly;
    }
}

// ---- Raft ----------------------------------------------------------------------------------------------------------
// leader appends n_new[g] entries in its current term (raft/request.rs: log.push(LogEntry{term: curr_term, ..}))
__global__ void __launch_bounds__(kEngThreads)
engine_raft_append_kernel(const uint32_t *__restrict__ n_new, uint64_t G, uint32_t W, const uint32_t *__restrict__ curr_term,
                          const uint32_t *__restrict__ last_commit, uint32_t *__restrict__ log_end, uint32_t *__restrict__ terms,
                          uint32_t *status) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kEngThreads;
    for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * kEngThreads + threadIdx.x; g < G; g += stride) {
        const uint32_t k = n_new[g];
        if (k == 0u) continue;
        const uint32_t le = log_end[g], ct = curr_term[g];
        // the ring holds the terms of slots last_commit+1 .. log_end-1: appending past W uncommitted entries would
        // overwrite one that the scan still needs
        if (le + k - last_commit[g] - 1u > W) { atomicOr(status, 2u); continue; }
        for (uint32_t i = 0; i < k; ++i) terms[g * W + ((le + i) & (W - 1u))] = ct;
        log_end[g] = le + k;
    }
}

// successful AppendEntriesReply (raft/messages.rs:243-252): ignored when next_slot[peer] > end_slot + 1, else
// next_slot = end_slot + 1 and match_slot = end_slot.  next only ever holds accepted end_slot + 1 values, so within a
// batch the outcome is the maximum in any order: atomicMax models it exactly.
__global__ void __launch_bounds__(kEngThreads)
engine_raft_ingest_kernel(const uint32_t *__restrict__ rec_group, const uint8_t *__restrict__ rec_peer,
                          const uint32_t *__restrict__ rec_end_slot, uint64_t n_rec, uint32_t P, uint64_t G,
                          uint32_t *next_slot, uint32_t *match, uint8_t *touched) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kEngThreads;
    for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * kEngThreads + threadIdx.x; i < n_rec; i += stride) {
        const uint64_t g = rec_group[i];
        const uint32_t peer = rec_peer[i], end = rec_end_slot[i];
        if (g >= G || peer >= P) continue;
        const uint64_t at = static_cast<uint64_t>(peer) * G + g;
        const uint32_t old = atomicMax(next_slot + at, end + 1u);
        if (old > end + 1u) continue;                          // :245-247
        atomicMax(match + at, end);
        touched[g] = 1;
    }
}

// last_snap advance (raft/messages.rs:298-309): after an accepted reply, the last slot every server stores
template <int NP>
__global__ void __launch_bounds__(kEngThreads)
engine_raft_snap_kernel(const uint32_t *__restrict__ match, uint32_t P, uint64_t G, uint8_t *__restrict__ touched,
                        uint32_t *__restrict__ last_snap) {
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kEngThreads;
    for (uint64_t g = static_cast<uint64_t>(blockIdx.x) * kEngThreads + threadIdx.x; g < G; g += stride) {
        if (!touched[g]) continue;
        touched[g] = 0;
        uint32_t mn = 0xffffffffu;
#pragma unroll
        for (int q = 0; q < NP; ++q)
            if (static_cast<uint32_t>(q) < P) { const uint32_t v = __ldg(match + static_cast<uint64_t>(q) * G + g); mn = v < mn ? v : mn; }
        if (P != 0u && mn > last_snap[g]) last_snap[g] = mn;
    }
}

template <typename T>
static int dalloc(T **p, size_t count, bool zero = true) {
    SS_CUDA(cudaMalloc(reinterpret_cast<void **>(p), count * sizeof(T) ? count * sizeof(T) : 1));
    if (zero) SS_CUDA(cudaMemset(*p, 0, count * sizeof(T)));
    return SS_OK;
}

}  // namespace ssb

using namespace ssb;

extern "C" {

int ss_engine_destroy(ss_engine *e) {
    if (e == nullptr) return SS_OK;
    if (e->ctx) {
        cudaSetDevice(e->ctx->device);
        cudaStreamSynchronize(e->ctx->stream);
    }
    void *bufs[] = {e->planes, e->bal_prepared, e->inst_bal, e->accepting, e->committed, e->commit_bar, e->shards, e->policy_idx,
                    e->policies, e->lut_bits, e->match, e->next_slot, e->last_commit, e->log_end, e->curr_term, e->terms,
                    e->last_snap, e->touched};
    for (void *b : bufs)
        if (b) cudaFree(b);
    if (e->coder) ss_rs_coder_destroy(e->coder);
    if (e->ctx) ctx_release(e->ctx);
    delete e;
    return SS_OK;
}

int ss_engine_create(ss_ctx *ctx, const ss_engine_config *cfg, uint64_t G, ss_engine **out) {
    if (out == nullptr) return set_error(SS_ERR_INVALID_ARG, "null out pointer");
    *out = nullptr;
    SS_TRY(ctx_bind(ctx));
    if (cfg == nullptr || G == 0) return set_error(SS_ERR_INVALID_ARG, "null config or zero groups");
    const uint32_t n = cfg->population;
    if (n == 0 || n > 12) return set_error(SS_ERR_INVALID_ARG, "population must be 1..12, got %u", n);
    const uint32_t majority = n / 2 + 1;
    if (cfg->protocol > SS_PROTO_CRAFT) return set_error(SS_ERR_INVALID_ARG, "unknown protocol %u", cfg->protocol);
    // fault_tolerance <= population - majority (rspaxos/mod.rs:599-605, crossword/mod.rs:752-760)
    if (cfg->protocol != SS_PROTO_MULTIPAXOS && cfg->protocol != SS_PROTO_RAFT && cfg->fault_tolerance > n - majority)
        return set_error(SS_ERR_INVALID_ARG, "invalid fault_tolerance %u for population %u", cfg->fault_tolerance, n);
    ss_engine *e = new (std::nothrow) ss_engine();
    if (!e) return set_error(SS_ERR_OUT_OF_MEMORY, "host allocation failed");
    e->ctx = ctx; e->cfg = *cfg; e->G = G; e->n = n; e->majority = majority;
    ctx_retain(ctx);
    int rc = SS_OK;
    auto fail = [&](int code) { ss_engine_destroy(e); return code; };
    const bool raft = cfg->protocol == SS_PROTO_RAFT || cfg->protocol == SS_PROTO_CRAFT;
    if (raft) {
        e->n_peers = n - 1;
        e->W = cfg->raft_window ? cfg->raft_window : 64;
        if ((e->W & (e->W - 1)) || e->W > 4096) return fail(set_error(SS_ERR_INVALID_ARG, "raft_window must be a power of two <= 4096"));
        // quorum_cnt (raft/mod.rs), or CRaft's majority + f (craft/messages.rs:300-308; full-copy mode = plain Raft)
        e->threshold = cfg->protocol == SS_PROTO_CRAFT ? majority + cfg->fault_tolerance : majority;
        if ((rc = dalloc(&e->match, size_t(e->n_peers) * G)) || (rc = dalloc(&e->next_slot, size_t(e->n_peers) * G)) ||
            (rc = dalloc(&e->last_commit, G)) || (rc = dalloc(&e->log_end, G)) || (rc = dalloc(&e->curr_term, G)) ||
            (rc = dalloc(&e->terms, G * e->W)) || (rc = dalloc(&e->last_snap, G)) || (rc = dalloc(&e->touched, G)))
            return fail(rc);
        // slot 0 is the sentinel entry (raft/mod.rs: log starts with a dummy entry): next_slot = 1, log_end = 1
        std::vector<uint32_t> ones(size_t(e->n_peers) * G > G ? size_t(e->n_peers) * G : G, 1u);
        cudaError_t ce = cudaMemcpy(e->next_slot, ones.data(), size_t(e->n_peers) * G * 4, cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) ce = cudaMemcpy(e->log_end, ones.data(), G * 4, cudaMemcpyHostToDevice);
        if (ce != cudaSuccess) return fail(cuda_error(ce, "initialise Raft state", __FILE__, __LINE__));
        *out = e;
        return SS_OK;
    }
    // ---- Paxos family ----
    switch (cfg->protocol) {
        case SS_PROTO_MULTIPAXOS: e->threshold = majority; e->d = 1; e->T = 1; break;                     // quorum_cnt (multipaxos/mod.rs:774)
        case SS_PROTO_RSPAXOS: e->threshold = majority + cfg->fault_tolerance; e->d = majority; e->T = n; break;
        default: {                                                                                       // crossword/mod.rs:805-830
            e->T = cfg->rs_total_shards ? cfg->rs_total_shards : n;
            e->d = cfg->rs_data_shards ? cfg->rs_data_shards : majority;
            if (e->T < n) e->T = n;
            if (e->d < majority) e->d = majority;
            if (e->T % n != 0 || e->d >= e->T || e->T > 32)
                return fail(set_error(SS_ERR_INVALID_ARG, "invalid Crossword code (T=%u, d=%u, n=%u)", e->T, e->d, n));
            e->threshold = 0;
        }
    }
    if ((rc = dalloc(&e->planes, size_t(n) * G)) || (rc = dalloc(&e->bal_prepared, G)) || (rc = dalloc(&e->inst_bal, G * 64)) ||
        (rc = dalloc(&e->accepting, G)) || (rc = dalloc(&e->committed, G)) || (rc = dalloc(&e->commit_bar, G)))
        return fail(rc);
    if (cfg->protocol != SS_PROTO_MULTIPAXOS) {
        if (cfg->data_len == 0) return fail(set_error(SS_ERR_INVALID_ARG, "data_len must be > 0 for an RS-coded protocol"));
        if ((rc = ss_rs_coder_create(ctx, static_cast<int>(e->d), static_cast<int>(e->T - e->d), &e->coder))) return fail(rc);
        e->L = (cfg->data_len + e->d - 1) / e->d;
        e->ds = (e->L + 15u) & ~15u;
        const uint32_t keep = cfg->keep_slots ? cfg->keep_slots : 1;
        e->cfg.keep_slots = keep;
        if ((rc = dalloc(&e->shards, size_t(keep) * e->T * G * e->ds))) return fail(rc);
    }
    if (cfg->protocol == SS_PROTO_CROSSWORD) {
        if ((rc = dalloc(&e->policy_idx, G * 64))) return fail(rc);
    }
    *out = e;
    return SS_OK;
}

int ss_engine_view_get(ss_engine *e, ss_engine_view *v) {
    if (e == nullptr || v == nullptr) return set_error(SS_ERR_INVALID_ARG, "null argument");
    memset(v, 0, sizeof(*v));
    v->n_groups = e->G; v->population = e->n; v->threshold = e->threshold; v->data_shards = e->d; v->total_shards = e->T;
    v->shard_len = e->L; v->shard_stride = e->ds; v->plane_stride = e->G * e->ds; v->slot_stride = uint64_t(e->T) * e->G * e->ds;
    v->planes = e->planes; v->bal_prepared = e->bal_prepared; v->inst_bal = e->inst_bal; v->accepting = e->accepting;
    v->committed = e->committed; v->commit_bar = e->commit_bar; v->shards = e->shards; v->policy_idx = e->policy_idx;
    v->match = e->match; v->next_slot = e->next_slot; v->last_commit = e->last_commit; v->log_end = e->log_end;
    v->curr_term = e->curr_term; v->terms = e->terms; v->last_snap = e->last_snap; v->raft_window = e->W;
    return SS_OK;
}

int ss_engine_set_prepared_ballots(ss_engine *e, const uint64_t *ballots_dev) {
    if (e == nullptr || ballots_dev == nullptr) return set_error(SS_ERR_INVALID_ARG, "null argument");
    if (e->bal_prepared == nullptr) return set_error(SS_ERR_INVALID_ARG, "not a Paxos-family engine");
    SS_TRY(ctx_bind(e->ctx));
    SS_CUDA(cudaMemcpyAsync(e->bal_prepared, ballots_dev, e->G * 8, cudaMemcpyDeviceToDevice, e->ctx->stream));
    return SS_OK;
}

int ss_engine_set_policies(ss_engine *e, const uint32_t *policies_host, uint32_t n_policies, int balanced) {
    if (e == nullptr || policies_host == nullptr) return set_error(SS_ERR_INVALID_ARG, "null argument");
    if (e->cfg.protocol != SS_PROTO_CROSSWORD) return set_error(SS_ERR_INVALID_ARG, "not a Crossword engine");
    if (n_policies == 0 || n_policies > 16) return set_error(SS_ERR_INVALID_ARG, "n_policies must be 1..16");
    ss_ctx *ctx = e->ctx;
    SS_TRY(ctx_bind(ctx));
    SS_CUDA(cudaStreamSynchronize(ctx->stream));
    if (e->policies) { cudaFree(e->policies); e->policies = nullptr; }
    if (e->lut_bits) { cudaFree(e->lut_bits); e->lut_bits = nullptr; }
    const uint32_t entries = n_policies << e->n;
    e->lut_words = (entries + 31) / 32;
    e->n_policies = n_policies;
    e->balanced = balanced;
    SS_TRY(dalloc(&e->policies, size_t(n_policies) * e->n, false));
    SS_TRY(dalloc(&e->lut_bits, e->lut_words));
    SS_CUDA(cudaMemcpy(e->policies, policies_host, size_t(n_policies) * e->n * 4, cudaMemcpyHostToDevice));
    return launch_crossword_lut(ctx, e->policies, n_policies, e->n, e->T, e->d, e->majority, e->cfg.fault_tolerance, balanced,
                                e->lut_bits);
}

int ss_engine_propose(ss_engine *e, uint32_t slot, const uint8_t *payloads, uint64_t payload_stride, const uint8_t *policy,
                      uint8_t **shard_planes) {
    if (e == nullptr) return set_error(SS_ERR_INVALID_ARG, "null engine");
    if (e->planes == nullptr) return set_error(SS_ERR_INVALID_ARG, "not a Paxos-family engine");
    if (slot >= 64) return set_error(SS_ERR_INVALID_ARG, "slot %u outside the 64-slot window", slot);
    ss_ctx *ctx = e->ctx;
    SS_TRY(ctx_bind(ctx));
    if (e->cfg.protocol == SS_PROTO_CROSSWORD && e->lut_bits == nullptr)
        return set_error(SS_ERR_INVALID_ARG, "ss_engine_set_policies must be called before proposing");
    if (e->coder != nullptr) {
        if (payloads == nullptr) return set_error(SS_ERR_INVALID_ARG, "null payloads");
        uint8_t *base = e->shards + uint64_t(slot % e->cfg.keep_slots) * e->T * e->G * e->ds;
        // from_data + compute_parity + the per-peer packed buffers of subset_copy: data shards into planes 0..d-1,
        // parity into planes d..T-1 (rspaxos/request.rs:72-77,127-142)
        SS_TRY(ss_rs_encode_uniform_dev(e->coder, payloads, payload_stride, e->cfg.data_len, e->G, base + uint64_t(e->d) * e->G * e->ds,
                                        e->G * e->ds, e->ds, SS_RS_OUT_PADDED16 | SS_RS_EMIT_DATA));
        if (shard_planes) *shard_planes = base;
    } else if (shard_planes) {
        *shard_planes = nullptr;                  // MultiPaxos sends the full batch to every peer: nothing to code
    }
    engine_propose_kernel<<<eng_grid(ctx, e->G), kEngThreads, 0, ctx->stream>>>(e->planes, e->n, e->G, slot, e->bal_prepared, e->inst_bal,
                                                                             e->accepting, e->committed, e->policy_idx, policy);
    SS_CUDA(cudaGetLastError());
    ctx->launches++;
    return SS_OK;
}

int ss_engine_ingest(ss_engine *e, const uint32_t *rec_group, const uint8_t *rec_slot, const uint8_t *rec_peer,
                     const uint64_t *rec_ballot, uint64_t n_records) {
    if (e == nullptr) return set_error(SS_ERR_INVALID_ARG, "null engine");
    if (e->planes == nullptr) return set_error(SS_ERR_INVALID_ARG, "not a Paxos-family engine");
    return ss_ack_ingest_dev(e->ctx, rec_group, rec_slot, rec_peer, rec_ballot, n_records, e->bal_prepared, e->inst_bal, e->accepting,
                             e->n, e->G, e->planes);
}

int ss_engine_tick(ss_engine *e, uint64_t *newly) {
    if (e == nullptr) return set_error(SS_ERR_INVALID_ARG, "null engine");
    ss_ctx *ctx = e->ctx;
    SS_TRY(ctx_bind(ctx));
    if (e->planes == nullptr) {
        // Raft / CRaft: commit scan in place (new_commit -> last_commit), then last_snap for the groups that ingested
        SS_TRY(launch_raft_scan_ring(ctx, e->match, e->n_peers, e->G, e->last_commit, e->log_end, e->curr_term, e->terms, e->W,
                                     e->threshold, e->last_commit));
        const uint32_t grid = eng_grid(ctx, e->G);
        if (e->n_peers <= 4) engine_raft_snap_kernel<4><<<grid, kEngThreads, 0, ctx->stream>>>(e->match, e->n_peers, e->G, e->touched, e->last_snap);
        else if (e->n_peers <= 8) engine_raft_snap_kernel<8><<<grid, kEngThreads, 0, ctx->stream>>>(e->match, e->n_peers, e->G, e->touched, e->last_snap);
        else engine_raft_snap_kernel<16><<<grid, kEngThreads, 0, ctx->stream>>>(e->match, e->n_peers, e->G, e->touched, e->last_snap);
        SS_CUDA(cudaGetLastError());
        ctx->launches++;
        return SS_OK;
    }
    if (e->cfg.protocol == SS_PROTO_CROSSWORD) {
        if (e->lut_bits == nullptr) return set_error(SS_ERR_INVALID_ARG, "ss_engine_set_policies must be called first");
        engine_tick_crossword_kernel<<<eng_grid(ctx, e->G), kEngThreads, e->lut_words * 4, ctx->stream>>>(
            e->planes, e->n, e->G, e->policy_idx, e->lut_bits, e->lut_words, e->n_policies, e->accepting, e->committed, e->commit_bar, newly);
    } else if ((e->G & 1ull) == 0ull && (newly == nullptr || (reinterpret_cast<uintptr_t>(newly) & 15u) == 0u)) {
        const uint64_t G2 = e->G / 2;
        const uint32_t grid = static_cast<uint32_t>((G2 + kEngThreads - 1) / kEngThreads);
        const ulonglong2 *p2 = reinterpret_cast<const ulonglong2 *>(e->planes);
        ulonglong2 *a2 = reinterpret_cast<ulonglong2 *>(e->accepting), *c2 = reinterpret_cast<ulonglong2 *>(e->committed);
        uint2 *b2 = reinterpret_cast<uint2 *>(e->commit_bar);
        ulonglong2 *n2 = reinterpret_cast<ulonglong2 *>(newly);
        if (e->n == 5) engine_tick_x2_kernel<5><<<grid, kEngThreads, 0, ctx->stream>>>(p2, e->n, G2, e->threshold, a2, c2, b2, n2);
        else if (e->n == 3) engine_tick_x2_kernel<3><<<grid, kEngThreads, 0, ctx->stream>>>(p2, e->n, G2, e->threshold, a2, c2, b2, n2);
        else if (e->n == 7) engine_tick_x2_kernel<7><<<grid, kEngThreads, 0, ctx->stream>>>(p2, e->n, G2, e->threshold, a2, c2, b2, n2);
        else engine_tick_x2_kernel<0><<<grid, kEngThreads, 0, ctx->stream>>>(p2, e->n, G2, e->threshold, a2, c2, b2, n2);
    } else {
        engine_tick_kernel<<<eng_grid(ctx, e->G), kEngThreads, 0, ctx->stream>>>(e->planes, e->n, e->G, e->threshold, e->accepting,
                                                                              e->committed, e->commit_bar, newly);
    }
    SS_CUDA(cudaGetLastError());
    ctx->launches++;
    return SS_OK;
}

int ss_engine_raft_append(ss_engine *e, const uint32_t *n_new) {
    if (e == nullptr || n_new == nullptr) return set_error(SS_ERR_INVALID_ARG, "null argument");
    if (e->match == nullptr) return set_error(SS_ERR_INVALID_ARG, "not a Raft engine");
    ss_ctx *ctx = e->ctx;
    SS_TRY(ctx_bind(ctx));
    engine_raft_append_kernel<<<eng_grid(ctx, e->G), kEngThreads, 0, ctx->stream>>>(n_new, e->G, e->W, e->curr_term, e->last_commit,
                                                                                 e->log_end, e->terms, ctx->dev_status);
    SS_CUDA(cudaGetLastError());
    ctx->launches++;
    return SS_OK;
}

int ss_engine_raft_ingest(ss_engine *e, const uint32_t *rec_group, const uint8_t *rec_peer, const uint32_t *rec_end_slot,
                          uint64_t n_records) {
    if (e == nullptr) return set_error(SS_ERR_INVALID_ARG, "null engine");
    if (e->match == nullptr) return set_error(SS_ERR_INVALID_ARG, "not a Raft engine");
    if (n_records == 0) return SS_OK;
    if (!rec_group || !rec_peer || !rec_end_slot) return set_error(SS_ERR_INVALID_ARG, "null buffer");
    ss_ctx *ctx = e->ctx;
    SS_TRY(ctx_bind(ctx));
    engine_raft_ingest_kernel<<<eng_grid(ctx, n_records), kEngThreads, 0, ctx->stream>>>(rec_group, rec_peer, rec_end_slot, n_records,
                                                                                      e->n_peers, e->G, e->next_slot, e->match, e->touched);
    SS_CUDA(cudaGetLastError());
    ctx->launches++;
    return SS_OK;
}

}  // extern "C"
