// PolynomialBatch::from_values / from_coeffs on the GPU: iNTT -> coset LDE (rate 2^rate_bits, shift 7) ->
// Poseidon leaf hashing -> Merkle cap.  (upstream plonky2::fri::oracle::PolynomialBatch; SURVEY.md 3.2 steps 3,6,9)
#include "ctx.h"
#include "poseidon.cuh"

u8 *ctx_arena(b200_ctx *ctx, size_t bytes) {
    if (bytes > ctx->arena_bytes) {
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        if (ctx->arena) cudaFree(ctx->arena);
        ctx->arena = nullptr; ctx->arena_bytes = 0;
        CUDA_TRY(cudaMalloc(&ctx->arena, bytes));
        ctx->arena_bytes = bytes;
    }
    return (u8 *)ctx->arena;
}


static const u64 *lde_scale_cached(b200_ctx *ctx, int log_n, int rate_bits, bool unit_shift) {
    auto key = std::make_pair(log_n, unit_shift ? rate_bits + 64 : rate_bits);
    auto it = ctx->lde_scale.find(key);
    if (it != ctx->lde_scale.end()) return it->second;
    u64 *p = nullptr;
    CUDA_TRY(cudaMalloc(&p, (size_t(8) << (log_n + rate_bits))));
    lde_scale_table(p, log_n, rate_bits, unit_shift ? 1 : GL_GEN, ctx->stream);
    ctx->lde_scale[key] = p;
    return p;
}
const u64 *ctx_lde_scale(b200_ctx *ctx, int log_n, int rate_bits) { return lde_scale_cached(ctx, log_n, rate_bits, false); }
// the same table for the subgroup itself (shift 1): extends values on H_{2^log_n} to H_{2^(log_n + rate_bits)}
const u64 *ctx_lde_scale_unit(b200_ctx *ctx, int log_n, int rate_bits) { return lde_scale_cached(ctx, log_n, rate_bits, true); }

void commit_hash_only(b200_ctx *ctx, const u64 *d_lde, size_t cols, int log_n, int rate_bits, int cap_height, u64 *d_digests, int hasher) {
    if (hasher == HASHER_BN254) {
        { PhaseTimer t(ctx, cols >= 64 ? "leaf_hash_wide" : "leaf_hash"); merkle_leaves_colmajor_bn254(d_lde, cols, log_n + rate_bits, d_digests, ctx->stream); }
        { PhaseTimer t(ctx, "merkle_levels"); merkle_build_bn254(d_digests, log_n + rate_bits, cap_height, ctx->stream); }
        return;
    }
    { PhaseTimer t(ctx, cols >= 64 ? "leaf_hash_wide" : "leaf_hash"); merkle_leaves_colmajor(d_lde, cols, log_n + rate_bits, d_digests, ctx->stream); }
    { PhaseTimer t(ctx, "merkle_levels"); merkle_build(d_digests, log_n + rate_bits, cap_height, ctx->stream); }
}

void commit_values_lde_only(b200_ctx *ctx, const u64 *d_values, size_t cols, int log_n, int rate_bits, u64 *d_coeffs_br, u64 *d_lde) {
    ntt_tables_ensure(ctx->tw, log_n + rate_bits, ctx->stream);
    const u64 *scale = ctx_lde_scale(ctx, log_n, rate_bits);
    size_t n = size_t(1) << log_n;
    { PhaseTimer t(ctx, "intt"); ntt_dif(ctx->tw, d_values, n, false, d_coeffs_br, n, log_n, cols, 0, nullptr, true, gl_inv((u64)n % GL_P), ctx->stream); }
    { PhaseTimer t(ctx, "lde"); ntt_dif(ctx->tw, d_coeffs_br, n, true, d_lde, n << rate_bits, log_n, cols << rate_bits, rate_bits, scale, false, 1, ctx->stream); }
}

void commit_coeffs(b200_ctx *ctx, const u64 *d_coeffs_br, size_t cols, int log_n, int rate_bits, int cap_height,
                   u64 *d_lde, u64 *d_digests, int hasher) {
    ntt_tables_ensure(ctx->tw, log_n + rate_bits, ctx->stream);
    const u64 *scale = ctx_lde_scale(ctx, log_n, rate_bits);
    size_t n = size_t(1) << log_n;
    { PhaseTimer t(ctx, "lde"); ntt_dif(ctx->tw, d_coeffs_br, n, true, d_lde, n << rate_bits, log_n, cols << rate_bits, rate_bits, scale, false, 1, ctx->stream); }
    if (hasher == HASHER_BN254) {
        { PhaseTimer t(ctx, cols >= 64 ? "leaf_hash_wide" : "leaf_hash"); merkle_leaves_colmajor_bn254(d_lde, cols, log_n + rate_bits, d_digests, ctx->stream); }
        { PhaseTimer t(ctx, "merkle_levels"); merkle_build_bn254(d_digests, log_n + rate_bits, cap_height, ctx->stream); }
        return;
    }
    { PhaseTimer t(ctx, cols >= 64 ? "leaf_hash_wide" : "leaf_hash"); merkle_leaves_colmajor(d_lde, cols, log_n + rate_bits, d_digests, ctx->stream); }
    { PhaseTimer t(ctx, "merkle_levels"); merkle_build(d_digests, log_n + rate_bits, cap_height, ctx->stream); }
}

void commit_values(b200_ctx *ctx, const u64 *d_values, size_t cols, int log_n, int rate_bits, int cap_height,
                   u64 *d_coeffs_br, u64 *d_lde, u64 *d_digests, int hasher) {
    ntt_tables_ensure(ctx->tw, log_n + rate_bits, ctx->stream);
    size_t n = size_t(1) << log_n;
    { PhaseTimer t(ctx, "intt"); ntt_dif(ctx->tw, d_values, n, false, d_coeffs_br, n, log_n, cols, 0, nullptr, true, gl_inv((u64)n % GL_P), ctx->stream); }
    commit_coeffs(ctx, d_coeffs_br, cols, log_n, rate_bits, cap_height, d_lde, d_digests, hasher);
}

__global__ void __launch_bounds__(128) k_permute_batch(u64 *st, size_t count) {
    __shared__ u64 sm[12 * 128];
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const bool active = i < count;
    if (!active) i = count - 1;
#pragma unroll
    for (int k = 0; k < 12; k++) sm[k * 128 + threadIdx.x] = st[12 * i + k];
    poseidon_permute_smem(&sm[threadIdx.x], 128);
    if (active) {
#pragma unroll
        for (int k = 0; k < 12; k++) st[12 * i + k] = sm[k * 128 + threadIdx.x];
    }
}
void poseidon_permute_batch(u64 *d_states, size_t count, cudaStream_t st) {
    k_permute_batch<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(d_states, count);
    LAUNCHED(1);
}
