// Context object behind b200_ctx and helpers shared by capi.cu / prover.cu.
#pragma once
#include "b200_internal.h"
#include <map>
#include <string>
#include <exception>
#include <new>

struct ApiErr { int code; std::string msg; };

struct b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    NttTables tw;
    std::map<std::pair<int, int>, u64 *> lde_scale;   // (log_n, rate_bits) -> coset scaling table [2^rate_bits][n] (br storage)
    std::string err;
};

// coset scaling table for the LDE, cached per size
const u64 *ctx_lde_scale(b200_ctx *ctx, int log_n, int rate_bits);
// PolynomialBatch::from_values on device buffers: values [cols][n] -> coeffs_br [cols][n], lde [cols][n<<rate], digests
void commit_values(b200_ctx *ctx, const u64 *d_values, size_t cols, int log_n, int rate_bits, int cap_height,
                   u64 *d_coeffs_br, u64 *d_lde, u64 *d_digests);
// PolynomialBatch::from_coeffs: coeffs_br given
void commit_coeffs(b200_ctx *ctx, const u64 *d_coeffs_br, size_t cols, int log_n, int rate_bits, int cap_height,
                   u64 *d_lde, u64 *d_digests);
void poseidon_permute_batch(u64 *d_states, size_t count, cudaStream_t st);
