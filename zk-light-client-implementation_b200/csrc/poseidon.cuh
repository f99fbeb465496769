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

// MDS: out[r] = sum_i circ[i] * s[(i+r)%12] + 8*s[0]*[r==0]; circ = [17,15,41,16,2,28,13,13,39,18,34,20]
GL_HD void poseidon_mds(u64 s[12]) {
    const u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    u32 lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { lo[i] = (u32)s[i]; hi[i] = (u32)(s[i] >> 32); }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        u64 al = 0, ah = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            al += (u64)lo[(i + r) % 12] * C[i];
            ah += (u64)hi[(i + r) % 12] * C[i];
        }
        if (r == 0) { al += (u64)lo[0] * 8u; ah += (u64)hi[0] * 8u; }
        s[r] = poseidon_reduce96(al, ah);
    }
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
    u64 nl = A.lo + lo; u64 c = nl < lo;
    u64 nh = A.hi + hi; u32 c2 = nh < hi;
    u64 nh2 = nh + c; c2 += nh2 < c;
    A.lo = nl; A.hi = nh2; A.top += c2;
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
