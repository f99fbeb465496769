// Poseidon permutation over Goldilocks (width 12, rate 8, x^7, 4 full + 22 partial + 4 full rounds),
// sponge helpers (hash_no_pad / hash_or_noop / two_to_one), host + device.
// Spec: plonky2::hash::poseidon (upstream, not vendored); algorithm as mirrored by
// gnark-plonky2-verifier/poseidon/goldilocks.go:30-37 (round structure), :92-115 (full / fast partial rounds),
// :251-331 (partial-first-constant layer, mds_partial_layer_init, mds_partial_layer_fast).
//
// Device layout: parameter tables live in __constant__ memory (every lane of a warp reads the same
// entry in the same cycle -> broadcast).  One thread runs one permutation with the 12-word state in
// registers; MDS uses the 32-bit-halves trick (sum of <= 13 products c*x with c < 2^6 and
// x < 2^32 fits a u64, IMAD.WIDE.U32 chain) followed by one 96-bit reduction per output.
#pragma once
#include "gl.cuh"
#include "poseidon_tables.cuh"

#define X(name, n, ...) static const u64 h_##name[n] = {__VA_ARGS__};
POSEIDON_TABLES(X)
#undef X
#ifdef __CUDACC__
#define X(name, n, ...) __constant__ u64 d_##name[n] = {__VA_ARGS__};
POSEIDON_TABLES(X)
#undef X
#endif

#ifdef __CUDA_ARCH__
#define PT(name) d_##name
#else
#define PT(name) h_##name
#endif

#define POSEIDON_WIDTH 12
#define POSEIDON_RATE 8

// x^7 with non-canonical intermediates (gl_mul_nc accepts any u64 inputs)
GL_HD u64 poseidon_sbox(u64 x) {
    u64 x2 = gl_mul_nc(x, x);
    u64 x4 = gl_mul_nc(x2, x2);
    u64 x3 = gl_mul_nc(x2, x);
    return gl_mul_nc(x4, x3);
}

// reduce (lo + hi * 2^32) with lo, hi < 2^46 -> [0, 2^64)
GL_HD u64 poseidon_reduce96(u64 lo, u64 hi) {
    u64 hh = hi >> 32;                 // * 2^64 == * EPS
    u64 hl = hi & GL_EPS;              // * 2^32
    u64 r = lo + (hl << 32);
    if (r < lo) r += GL_EPS;           // cannot wrap twice: lo < 2^46
    u64 t = hh * GL_EPS;               // hh < 2^14
    u64 r2 = r + t;
    if (r2 < t) r2 += GL_EPS;
    return r2;
}

// MDS: out[r] = sum_i circ[i] * s[(i+r)%12] + 8*s[0]*[r==0]; circ = [17,15,41,16,2,28,13,13,39,18,34,20].
// The circulant is a length-12 cyclic convolution; split indices mod 3, take a length-4 real FFT of every class and the
// convolution becomes three 3x3 "twisted" products whose constants are all powers of two for this matrix:
//   k=0 (Y=1):   (64, 128, 64)      k=2 (Y=-1): (-4, -32, 8)      k=1 (Y=-i): (4-2i, -8+2i, 32+2i)
// so the whole layer is adds and shifts on signed 64-bit integers (no multiplier).  Run separately on the low and the
// high 32-bit halves of the state (|intermediates| < 2^46), then recombine lo + hi*2^32 and reduce.
typedef long long i64;
GL_HD void poseidon_mds_half(const i64 x[12], i64 y[12]) {
    i64 U0[3], U2[3], Ur[3], Ui[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        i64 t0 = x[r] + x[r + 6], t1 = x[r] - x[r + 6], t2 = x[r + 3] + x[r + 9], t3 = x[r + 3] - x[r + 9];
        U0[r] = t0 + t2; U2[r] = t0 - t2; Ur[r] = t1; Ui[r] = -t3;
    }
    const i64 S = U0[0] + U0[1] + U0[2];
    i64 b0[3] = {(S + U0[2]) << 4, (S + U0[0]) << 4, (S + U0[1]) << 4};
    i64 b2[3] = {(U2[2] << 3) - U2[0] - (U2[1] << 1), -U2[1] - (U2[0] << 3) - (U2[2] << 1), (U2[0] << 1) - U2[2] - (U2[1] << 3)};
    // (K/2) * U with K0/2 = 2-i, K1/2 = -4+i, K2/2 = 16+i
#define K0RE(b) ((Ur[b] << 1) + Ui[b])
#define K0IM(b) ((Ui[b] << 1) - Ur[b])
#define K1RE(b) (-(Ur[b] << 2) - Ui[b])
#define K1IM(b) (Ur[b] - (Ui[b] << 2))
#define K2RE(b) ((Ur[b] << 4) - Ui[b])
#define K2IM(b) ((Ui[b] << 4) + Ur[b])
    // w*(u+vi) = v - ui  (w = -i)
    i64 Wre[3], Wim[3];
    { i64 ar = K1RE(2) + K2RE(1), ai = K1IM(2) + K2IM(1); Wre[0] = K0RE(0) + ai; Wim[0] = K0IM(0) - ar; }
    { i64 ar = K2RE(2), ai = K2IM(2); Wre[1] = K0RE(1) + K1RE(0) + ai; Wim[1] = K0IM(1) + K1IM(0) - ar; }
    Wre[2] = K0RE(2) + K1RE(1) + K2RE(0); Wim[2] = K0IM(2) + K1IM(1) + K2IM(0);
#undef K0RE
#undef K0IM
#undef K1RE
#undef K1IM
#undef K2RE
#undef K2IM
#pragma unroll
    for (int t = 0; t < 3; t++) {
        i64 s = b0[t] + b2[t], d = b0[t] - b2[t];
        y[t] = s + Wre[t]; y[t + 6] = s - Wre[t]; y[t + 3] = d - Wim[t]; y[t + 9] = d + Wim[t];
    }
}

// Same transform on 32-bit signed integers (inputs < 2^22: |intermediates| < 2^31), used on three 22-bit limbs so that
// every add is ONE instruction (the 64-bit version needs an add + add-with-carry pair per operation).
typedef int i32;
GL_HD void poseidon_mds_limb(const i32 x[12], i32 y[12]) {
    i32 U0[3], U2[3], Ur[3], Ui[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        i32 t0 = x[r] + x[r + 6], t1 = x[r] - x[r + 6], t2 = x[r + 3] + x[r + 9], t3 = x[r + 3] - x[r + 9];
        U0[r] = t0 + t2; U2[r] = t0 - t2; Ur[r] = t1; Ui[r] = -t3;
    }
    const i32 S = U0[0] + U0[1] + U0[2];
    i32 b0[3] = {(S + U0[2]) << 4, (S + U0[0]) << 4, (S + U0[1]) << 4};
    i32 b2[3] = {(U2[2] << 3) - U2[0] - (U2[1] << 1), -U2[1] - (U2[0] << 3) - (U2[2] << 1), (U2[0] << 1) - U2[2] - (U2[1] << 3)};
#define K0RE(b) ((Ur[b] << 1) + Ui[b])
#define K0IM(b) ((Ui[b] << 1) - Ur[b])
#define K1RE(b) (-(Ur[b] << 2) - Ui[b])
#define K1IM(b) (Ur[b] - (Ui[b] << 2))
#define K2RE(b) ((Ur[b] << 4) - Ui[b])
#define K2IM(b) ((Ui[b] << 4) + Ur[b])
    i32 Wre[3], Wim[3];
    { i32 ar = K1RE(2) + K2RE(1), ai = K1IM(2) + K2IM(1); Wre[0] = K0RE(0) + ai; Wim[0] = K0IM(0) - ar; }
    { i32 ar = K2RE(2), ai = K2IM(2); Wre[1] = K0RE(1) + K1RE(0) + ai; Wim[1] = K0IM(1) + K1IM(0) - ar; }
    Wre[2] = K0RE(2) + K1RE(1) + K2RE(0); Wim[2] = K0IM(2) + K1IM(1) + K2IM(0);
#undef K0RE
#undef K0IM
#undef K1RE
#undef K1IM
#undef K2RE
#undef K2IM
#pragma unroll
    for (int t = 0; t < 3; t++) {
        i32 s = b0[t] + b2[t], d = b0[t] - b2[t];
        y[t] = s + Wre[t]; y[t + 6] = s - Wre[t]; y[t + 3] = d - Wim[t]; y[t + 9] = d + Wim[t];
    }
}

// v = y0 + y1*2^22 + y2*2^44 (0 <= y_k < 2^31) reduced to [0, 2^64)
GL_HD u64 poseidon_recombine3(u32 y0, u32 y1, u32 y2) {
    u64 lo = (u64)y0 + ((u64)y1 << 22);            // < 2^54
    u64 t = (u64)y2 << 12;                         // y2 * 2^44 = t * 2^32,  t < 2^43
    return poseidon_reduce96(lo, t);
}

GL_HD void poseidon_mds(u64 s[12]) {
#ifdef __CUDA_ARCH__
    i32 l0[12], l1[12], l2[12], y0[12], y1[12], y2[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        l0[i] = (i32)((u32)s[i] & 0x3FFFFFu); l1[i] = (i32)((u32)(s[i] >> 22) & 0x3FFFFFu); l2[i] = (i32)(u32)(s[i] >> 44);
    }
    poseidon_mds_limb(l0, y0); poseidon_mds_limb(l1, y1); poseidon_mds_limb(l2, y2);
    y0[0] += l0[0] << 3; y1[0] += l1[0] << 3; y2[0] += l2[0] << 3;          // diagonal term 8*s[0]
#pragma unroll
    for (int r = 0; r < 12; r++) s[r] = poseidon_recombine3((u32)y0[r], (u32)y1[r], (u32)y2[r]);
#else
    i64 lo[12], hi[12], yl[12], yh[12];
    for (int i = 0; i < 12; i++) { lo[i] = (i64)(u32)s[i]; hi[i] = (i64)(s[i] >> 32); }
    poseidon_mds_half(lo, yl);
    poseidon_mds_half(hi, yh);
    yl[0] += lo[0] << 3; yh[0] += hi[0] << 3;          // diagonal term 8*s[0]
    for (int r = 0; r < 12; r++) s[r] = poseidon_reduce96((u64)yl[r], (u64)yh[r]);
#endif
}

GL_HD void poseidon_full_rounds(u64 s[12], int rc_off) {
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) {
            u64 c = PT(POSEIDON_RC)[rc_off + 12 * r + i];
            u64 x = s[i] + c;                       // s[i] may be non-canonical: handle wrap only
            if (x < c) x += GL_EPS;
            s[i] = poseidon_sbox(x);
        }
        poseidon_mds(s);
    }
}

// acc (3 words: lo, hi, top) += a*b  (used for the dot products of the fast partial rounds)
struct acc192 { u64 lo, hi; u32 top; };
GL_HD void acc192_mad(acc192 &A, u64 a, u64 b) {
    u64 lo, hi; gl_mul_wide(a, b, lo, hi);
#ifdef __CUDA_ARCH__
    asm("add.cc.u64 %0, %0, %3;\n\taddc.cc.u64 %1, %1, %4;\n\taddc.u32 %2, %2, 0;" : "+l"(A.lo), "+l"(A.hi), "+r"(A.top) : "l"(lo), "l"(hi));
#else
    u64 nl = A.lo + lo; u64 c = nl < lo;
    u64 nh = A.hi + hi; u32 c2 = nh < hi;
    u64 nh2 = nh + c; c2 += nh2 < c;
    A.lo = nl; A.hi = nh2; A.top += c2;
#endif
}
GL_HD u64 acc192_reduce(const acc192 &A) {
    // top * 2^128 : 2^128 = 2^64 * 2^64 == EPS * EPS == 2^64 - 2*2^32 + 1 == (2^32-1) - 2^33 + 2 = -2^32 (mod p)
    u64 r = gl_reduce128_nc(A.lo, A.hi);
    r = gl_canon(r);
    u64 t = (u64)A.top << 32;                       // top < 2^5
    return gl_sub(r, t);
}

GL_HD void poseidon_partial_rounds(u64 s[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add(gl_canon(s[i]), PT(POSEIDON_FP_FIRST_RC)[i]);
    {   // mds_partial_layer_init: res[d] = sum_{r=1..11} s[r] * INIT[r-1][d-1]
        u64 o[12]; o[0] = s[0];
#pragma unroll
        for (int d = 1; d < 12; d++) {
            acc192 A = {0, 0, 0};
#pragma unroll
            for (int r = 1; r < 12; r++) acc192_mad(A, s[r], PT(POSEIDON_FP_INIT)[11 * (r - 1) + (d - 1)]);
            o[d] = acc192_reduce(A);
        }
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = o[i];
    }
    for (int r = 0; r < 22; r++) {
        u64 x = gl_add(gl_canon(poseidon_sbox(s[0])), PT(POSEIDON_FP_RC)[r]);
        acc192 A = {0, 0, 0};
        acc192_mad(A, x, 25);                       // MDS0TO0 = circ[0] + diag[0]
#pragma unroll
        for (int i = 1; i < 12; i++) acc192_mad(A, s[i], PT(POSEIDON_FP_W_HATS)[11 * r + i - 1]);
#pragma unroll
        for (int i = 1; i < 12; i++) s[i] = gl_add(s[i], gl_mul(x, PT(POSEIDON_FP_VS)[11 * r + i - 1]));
        s[0] = acc192_reduce(A);
    }
}

// in/out canonical
GL_HD void poseidon_permute(u64 s[12]) {
#ifndef __CUDA_ARCH__
    // host (challenger, public-input hash, verifier): textbook rounds -- with the multiplier-free MDS a partial round costs one
    // S-box + ~200 adds, cheaper on a CPU than the 23 full products of the sparse form below (3.0 vs 7.3 us per permutation)
    for (int r = 0; r < 30; r++) {
        const bool full = (r < 4) || (r >= 26);
        const u64 *rc = PT(POSEIDON_RC) + 12 * r;
        for (int i = 0; i < 12; i++) {
            u64 c = rc[i], x = s[i] + c;
            if (x < c) x += GL_EPS;
            s[i] = (full || i == 0) ? poseidon_sbox(x) : x;
        }
        poseidon_mds(s);
    }
    for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
    return;
#endif
    poseidon_full_rounds(s, 0);
    poseidon_partial_rounds(s);
    poseidon_full_rounds(s, 12 * 26);
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
}

GL_HD void poseidon_two_to_one(const u64 l[4], const u64 r[4], u64 out[4]) {
    u64 s[12] = {l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0};
    poseidon_permute(s);
    out[0] = s[0]; out[1] = s[1]; out[2] = s[2]; out[3] = s[3];
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------------
// Compact device permutation.  ncu on the first version (state in registers, everything unrolled: ~15k SASS
// instructions per kernel) showed issue-slot utilisation of 35 % with "no_instruction" as the dominant stall: the
// instruction stream did not fit the SM's instruction caches.  This version keeps the 12-word state in SHARED memory
// (element-major: word i of thread t at st[i * stride], conflict-free) so that the S-box layer is a rolled loop and the
// 30 rounds share one copy of the (unrolled, shift/add-only) MDS code: ~1.3k instructions total.  Partial rounds use
// the textbook form (S-box on word 0, all 12 round constants, full MDS) because with the multiplier-free MDS that is
// cheaper than the "fast" sparse-matrix form (23 modmuls) and needs no extra tables.  Same permutation, bit-exact.
// ---------------------------------------------------------------------------------------------------------------
static __device__ __noinline__ void poseidon_permute_smem(u64 *st, const int stride) {
    for (int r = 0; r < 30; r++) {
        const bool full = (r < 4) || (r >= 26);
        const u64 *rc = d_POSEIDON_RC + 12 * r;
        u64 s[12];
        if (full) {
#pragma unroll 2
            for (int i = 0; i < 12; i++) {
                u64 c = rc[i];
                u64 x = st[i * stride] + c;
                if (x < c) x += GL_EPS;
                st[i * stride] = poseidon_sbox(x);
            }
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = st[i * stride];
        } else {
#pragma unroll
            for (int i = 0; i < 12; i++) {
                u64 c = rc[i];
                u64 x = st[i * stride] + c;
                if (x < c) x += GL_EPS;
                s[i] = x;
            }
            s[0] = poseidon_sbox(s[0]);
        }
        poseidon_mds(s);
#pragma unroll
        for (int i = 0; i < 12; i++) st[i * stride] = s[i];
    }
#pragma unroll
    for (int i = 0; i < 12; i++) st[i * stride] = gl_canon(st[i * stride]);
}
#endif

// host-side sponge helpers (used by the challenger / public-input hash)
inline void poseidon_hash_no_pad_host(const u64 *in, size_t n, u64 out[4]) {
    u64 s[12] = {0};
    for (size_t i = 0; i < n; i += 8) {
        size_t m = n - i < 8 ? n - i : 8;
        for (size_t j = 0; j < m; j++) s[j] = in[i + j];
        poseidon_permute(s);
    }
    for (int i = 0; i < 4; i++) out[i] = s[i];
}
