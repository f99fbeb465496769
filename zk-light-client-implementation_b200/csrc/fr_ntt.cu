// BN254 scalar-field NTT for sm_100a -- SURVEY.md section 8f rank 3 / BASELINE config 4: the FFTs of gnark's groth16.Prove
// (quotient H = (A.B - C)/Z: three inverse FFTs, three coset FFTs, one coset inverse FFT over a domain of >= nbConstraints
// points; gnark-crypto v0.12.2 ecc/bn254/fr/fft, not vendored; call site gnark-plonky2-verifier/cmd/web-api.go:77).
// Parameters restated from the published domain: generator 5, 2^28-th root of unity ROOT_2_28 (oracle/bn254_fr_ntt.py).
//
// Elements are 32 bytes (4 u64, canonical little-endian at the ABI, Montgomery inside).  One element = one 32-byte DRAM
// sector, so the radix-8 decimation-in-frequency passes work straight on global memory: a thread owns the 8 points
// base + e * 2^q, does three butterfly layers (12 Montgomery products, mont256.cuh) in registers and writes them back;
// consecutive threads take consecutive low indices, i.e. consecutive sectors.  log2(n)/3 passes of 64 n bytes each, then
// one bit-reversal pass that also leaves Montgomery form and applies 1/n and the coset factors.
// Twiddles: tw[h + j] = w_{2h}^j (Montgomery), built level by level (one product per two entries), cached per context.
#include "../../include/b200_plonky2.h"
#include "ctx.h"
#include "bn254.cuh"
#include <vector>

namespace {

const u64 FR_ROOT_2_28[4] = {0x9bd61b6e725b19f0ULL, 0x402d111e41112ed4ULL, 0x00e0a7eb8ef62abcULL, 0x2a3c09f0a58a7e85ULL};
// = 19103219067921713944291392827692070036145651957329286315305642004821462161904

struct FrLevels { fr_t w[29]; };          // w[b] = primitive 2^b-th root (Montgomery), direction-specific

__global__ void k_fr_tw_level(fr_t *tw, int b, fr_t w_next) {       // fills tw[2^b .. 2^(b+1)) from level b-1; w_next = w_{2^(b+1)}
    size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t h = size_t(1) << b;
    if (j >= h) return;
    fr_t x = tw[(h >> 1) + (j >> 1)];
    if (j & 1) x = fr_mul(x, w_next);
    tw[h + j] = x;
}
// powers of the coset generator: lo[i] = g^i (i < 4096), hi[i] = g^(4096 i); Montgomery
__global__ void k_fr_pow_tables(fr_t *lo, fr_t *hi, size_t n_hi, fr_t g, fr_t one) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= 4096 + n_hi) return;
    u64 e = t < 4096 ? t : (t - 4096) * 4096;
    fr_t r = one, b = g;
    while (e) { if (e & 1) r = fr_mul(r, b); b = fr_mul(b, b); e >>= 1; }
    if (t < 4096) lo[t] = r; else hi[t - 4096] = r;
}
__global__ void k_fr_load(const u64 *__restrict__ in, fr_t *__restrict__ data, size_t n, const fr_t *__restrict__ plo, const fr_t *__restrict__ phi) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 l[4] = {in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3]};
    fr_t x = fr_from_canonical(l);
    if (plo) x = fr_mul(x, fr_mul(plo[i & 4095], phi[i >> 12]));      // a[i] *= g^i (forward transform on the coset)
    data[i] = x;
}
// out[bitrev(i)] = data[i] * c, leaving Montgomery form; inverse transform on the coset: additionally * g^-k, k = bitrev(i)
__global__ void k_fr_store(const fr_t *__restrict__ data, u64 *__restrict__ out, int log_n, fr_t c_plain, const fr_t *__restrict__ plo, const fr_t *__restrict__ phi) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t n = size_t(1) << log_n;
    if (i >= n) return;
    size_t k = log_n ? (size_t)(__brevll((unsigned long long)i) >> (64 - log_n)) : 0;
    fr_t x = data[i];
    if (plo) x = fr_mul(x, fr_mul(plo[k & 4095], phi[k >> 12]));
    x = fr_mul(x, c_plain);                                            // (x R) * c / R = x c, canonical
#pragma unroll
    for (int j = 0; j < 4; j++) out[4 * k + j] = x.v[j];
}
// r butterfly layers on mid bits q+r-1 .. q, in place
template <int r>
__global__ void __launch_bounds__(128) k_fr_dif_pass(fr_t *__restrict__ data, const fr_t *__restrict__ tw, int log_n, int q) {
    size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t(1) << (log_n - r))) return;
    const size_t lo = t & ((size_t(1) << q) - 1), hi = t >> q;
    fr_t *p = data + ((hi << (q + r)) | lo);
    fr_t x[1 << r];
#pragma unroll
    for (int e = 0; e < (1 << r); e++) x[e] = p[(size_t)e << q];
#pragma unroll
    for (int s = r - 1; s >= 0; s--) {
        const size_t h = size_t(1) << (q + s);
#pragma unroll
        for (int e = 0; e < (1 << r); e++) {
            if (e & (1 << s)) continue;
            const size_t j = ((size_t)(e & ((1 << s) - 1)) << q) | lo;    // (index of the upper element) mod h
            fr_t u = x[e], v = x[e + (1 << s)];
            x[e] = fr_add(u, v);
            x[e + (1 << s)] = fr_mul(fr_sub(u, v), tw[h + j]);
        }
    }
#pragma unroll
    for (int e = 0; e < (1 << r); e++) p[(size_t)e << q] = x[e];
}

fr_t fr_inv_host(const fr_t &a) {          // Fermat
    u64 e[4] = {FR_M0 - 2, FR_M1, FR_M2, FR_M3};
    const u64 one_c[4] = {1, 0, 0, 0};
    fr_t r = fr_from_canonical(one_c), b = a;
    for (int i = 0; i < 256; i++) { if ((e[i / 64] >> (i % 64)) & 1) r = fr_mul(r, b); b = fr_mul(b, b); }
    return r;
}

}  // namespace

int ctx_fail(b200_ctx *ctx, int code, const std::string &msg);

// twiddle tables of one direction up to 2^log_n points, cached in the context
static const fr_t *fr_twiddles(b200_ctx *ctx, int log_n, bool inverse) {
    int &have = inverse ? ctx->fr_tw_log[1] : ctx->fr_tw_log[0];
    void *&tab = inverse ? ctx->fr_tw[1] : ctx->fr_tw[0];
    if (tab && have >= log_n) return (const fr_t *)tab;
    if (tab) { CUDA_TRY(cudaStreamSynchronize(ctx->stream)); cudaFree(tab); tab = nullptr; have = 0; }
    const int lg = log_n < 16 ? 16 : log_n;
    CUDA_TRY(cudaMalloc(&tab, (size_t(1) << lg) * sizeof(fr_t)));
    fr_t w[29];
    w[28] = fr_from_canonical(FR_ROOT_2_28);
    if (inverse) w[28] = fr_inv_host(w[28]);
    for (int b = 27; b >= 0; b--) w[b] = fr_mul(w[b + 1], w[b + 1]);
    fr_t *tw = (fr_t *)tab;
    CUDA_TRY(cudaMemcpyAsync(tw + 1, &w[0], sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));     // tw[1] = 1
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int b = 1; b < lg; b++) {
        size_t h = size_t(1) << b;
        k_fr_tw_level<<<(unsigned)((h + 255) / 256), 256, 0, ctx->stream>>>(tw, b, w[b + 1]);
        LAUNCHED(1);
    }
    have = lg;
    return tw;
}

static int fr_ntt_dev(b200_ctx *ctx, const uint64_t *d_in, uint32_t log_n, int inverse, int coset, uint64_t *d_out) {
    if (!ctx) return B200_ERR_BAD_ARG;
    try {
        cudaSetDevice(ctx->device);
        if (!d_in || !d_out || log_n > 28) throw ApiErr{B200_ERR_BAD_ARG, "b200_fr_ntt: bad arguments (log_n <= 28)"};
        cudaStream_t st = ctx->stream;
        const size_t n = size_t(1) << log_n, n_hi = (n + 4095) / 4096;
        const fr_t *tw = log_n ? fr_twiddles(ctx, (int)log_n, inverse != 0) : nullptr;
        auto rnd = [](size_t b) { return (b + 255) & ~size_t(255); };
        u8 *base = ctx_arena(ctx, rnd(n * sizeof(fr_t)) + rnd(4096 * sizeof(fr_t)) + rnd(n_hi * sizeof(fr_t)));
        fr_t *data = (fr_t *)base, *plo = (fr_t *)(base + rnd(n * sizeof(fr_t))), *phi = (fr_t *)((u8 *)plo + rnd(4096 * sizeof(fr_t)));
        const u64 one_c[4] = {1, 0, 0, 0}, gen_c[4] = {5, 0, 0, 0};
        if (coset) {
            fr_t g = fr_from_canonical(gen_c);
            if (inverse) g = fr_inv_host(g);
            PhaseTimer t(ctx, "fr_ntt_coset_tables");
            k_fr_pow_tables<<<(unsigned)((4096 + n_hi + 127) / 128), 128, 0, st>>>(plo, phi, n_hi, g, fr_from_canonical(one_c));
            LAUNCHED(1);
        }
        { PhaseTimer t(ctx, "fr_ntt_load");
          k_fr_load<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_in, data, n, (coset && !inverse) ? plo : nullptr, phi); LAUNCHED(1); }
        { PhaseTimer t(ctx, "fr_ntt_passes");
          int q = (int)log_n;
          const int rem = q % 3;
          if (rem == 1) { q -= 1; k_fr_dif_pass<1><<<(unsigned)(((n >> 1) + 127) / 128), 128, 0, st>>>(data, tw, (int)log_n, q); LAUNCHED(1); }
          else if (rem == 2) { q -= 2; k_fr_dif_pass<2><<<(unsigned)(((n >> 2) + 127) / 128), 128, 0, st>>>(data, tw, (int)log_n, q); LAUNCHED(1); }
          while (q > 0) { q -= 3; k_fr_dif_pass<3><<<(unsigned)(((n >> 3) + 127) / 128), 128, 0, st>>>(data, tw, (int)log_n, q); LAUNCHED(1); } }
        fr_t c; for (int i = 0; i < 4; i++) c.v[i] = one_c[i];
        if (inverse) {                                   // 1/n as a plain (non-Montgomery) integer: from_mont(inv(to_mont(n)))
            const u64 n_c[4] = {(u64)n, 0, 0, 0};
            fr_t ninv = fr_inv_host(fr_from_canonical(n_c));
            fr_to_canonical(ninv, c.v);
        }
        { PhaseTimer t(ctx, "fr_ntt_store");
          k_fr_store<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(data, d_out, (int)log_n, c, (coset && inverse) ? plo : nullptr, phi); LAUNCHED(1); }
        return B200_OK;
    } catch (const CudaErr &e) {
        char buf[256]; snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d", (int)e.e, cudaGetErrorString(e.e), e.file, e.line);
        cudaGetLastError(); return ctx_fail(ctx, e.e == cudaErrorMemoryAllocation ? B200_ERR_OOM : B200_ERR_CUDA, buf);
    } catch (const ApiErr &e) { return ctx_fail(ctx, e.code, e.msg); }
}

int fr_ntt_dev_internal(b200_ctx *ctx, const uint64_t *d_in, uint32_t log_n, int inverse, int coset, uint64_t *d_out) { return fr_ntt_dev(ctx, d_in, log_n, inverse, coset, d_out); }
extern "C" int b200_fr_ntt_dev(b200_ctx *ctx, const uint64_t *d_values, uint32_t log_n, int inverse, int coset, uint64_t *d_out) {
    return fr_ntt_dev(ctx, d_values, log_n, inverse, coset, d_out);
}
extern "C" int b200_fr_ntt(b200_ctx *ctx, const uint64_t *values, uint32_t log_n, int inverse, int coset, uint64_t *out) {
    if (!ctx || !values || !out || log_n > 28) return B200_ERR_BAD_ARG;
    try {
        cudaSetDevice(ctx->device);
        const size_t bytes = (size_t(32) << log_n);
        DevBuf a(bytes), b(bytes);
        CUDA_TRY(cudaMemcpyAsync(a.p, values, bytes, cudaMemcpyHostToDevice, ctx->stream));
        int rc = fr_ntt_dev(ctx, a.u(), log_n, inverse, coset, b.u());
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(out, b.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        return B200_OK;
    } catch (const CudaErr &e) { return ctx_fail(ctx, e.e == cudaErrorMemoryAllocation ? B200_ERR_OOM : B200_ERR_CUDA, "CUDA error in b200_fr_ntt"); }
}
