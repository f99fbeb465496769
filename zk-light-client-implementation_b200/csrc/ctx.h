// Context object behind b200_ctx and helpers shared by capi.cu / prover.cu.
#pragma once
#include "b200_internal.h"
#include <map>
#include <string>
#include <exception>
#include <new>
#include <memory>
#include <vector>
#include <chrono>

struct ApiErr { int code; std::string msg; };
struct b200_circuit;

struct b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    NttTables tw;
    std::map<std::pair<int, int>, u64 *> lde_scale;   // (log_n, rate_bits [+64: shift 1]) -> coset scaling table [2^rate_bits][n] (br storage)
    std::string err;
    void *arena = nullptr; size_t arena_bytes = 0;   // grow-only scratch for b200_prove
    u64 *tmp_q_big = nullptr;
    void *pinned = nullptr;                           // pinned staging for the small per-proof transfers
    cudaStream_t copy_stream = nullptr;               // H2D of host witnesses, overlapped with the LDE
    std::vector<b200_circuit *> circuit_cache; unsigned long long cache_tick = 0, cache_hits = 0, cache_misses = 0;   // b200_circuit_build
    void *fr_tw[2] = {nullptr, nullptr}; int fr_tw_log[2] = {0, 0};   // BN254 F_r twiddles (fr_ntt.cu): forward, inverse
    // optional phase timing (b200_ctx_enable_timing): CUDA events on the context stream around prover phases
    bool timing = false;
    std::vector<std::pair<std::string, std::pair<cudaEvent_t, cudaEvent_t>>> timing_events;
    std::string timing_json;
    double host_total_ms = 0, host_wait_ms = 0; int host_proofs = 0, host_syncs = 0;   // wall clock of prove() and of its stream syncs
    std::vector<double> host_seg_ms; size_t host_seg_i = 0; std::chrono::steady_clock::time_point host_seg_t0;   // host-busy time before the i-th sync of a proof
};
struct PhaseTimer {     // RAII: records start/stop events for one named phase when timing is enabled
    b200_ctx *ctx; size_t idx; bool on;
    PhaseTimer(b200_ctx *c, const char *name) : ctx(c), idx(0), on(c->timing) {
        if (!on) return;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, c->stream);
        idx = c->timing_events.size();
        c->timing_events.push_back({name, {a, b}});
    }
    ~PhaseTimer() { if (on) cudaEventRecord(ctx->timing_events[idx].second.second, ctx->stream); }
};

// grow-only device scratch owned by the context (one allocation reused by every prove / msm call)
u8 *ctx_arena(b200_ctx *ctx, size_t bytes);
// coset scaling table for the LDE, cached per size
const u64 *ctx_lde_scale(b200_ctx *ctx, int log_n, int rate_bits);
const u64 *ctx_lde_scale_unit(b200_ctx *ctx, int log_n, int rate_bits);     // shift 1 instead of the coset generator
// PolynomialBatch::from_values on device buffers: values [cols][n] -> coeffs_br [cols][n], lde [cols][n<<rate], digests
void commit_values(b200_ctx *ctx, const u64 *d_values, size_t cols, int log_n, int rate_bits, int cap_height,
                   u64 *d_coeffs_br, u64 *d_lde, u64 *d_digests, int hasher = 0);
// PolynomialBatch::from_coeffs: coeffs_br given
void commit_coeffs(b200_ctx *ctx, const u64 *d_coeffs_br, size_t cols, int log_n, int rate_bits, int cap_height,
                   u64 *d_lde, u64 *d_digests, int hasher = 0);
// the two halves of commit_values, for pipelining the host->device copy of the witness with the LDE
void commit_values_lde_only(b200_ctx *ctx, const u64 *d_values, size_t cols, int log_n, int rate_bits, u64 *d_coeffs_br, u64 *d_lde);
void commit_hash_only(b200_ctx *ctx, const u64 *d_lde, size_t cols, int log_n, int rate_bits, int cap_height, u64 *d_digests, int hasher);
void poseidon_permute_batch(u64 *d_states, size_t count, cudaStream_t st);
void ctx_drop_circuit_cache(b200_ctx *ctx);          // prover.cu
