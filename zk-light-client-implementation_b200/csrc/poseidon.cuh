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

// ---------------------------------------------------------------------------------------------------------------
// Limb-resident permutation (round 2).  ncu on the round-1 kernel: 33.2 k thread-instructions per permutation at 74 % of
// the issue slots -- only fewer instructions help.  Where they went: every MDS layer re-split its 12 inputs into 22-bit
// limbs and recombined + reduced its 12 outputs to 64 bits (~25 of ~50 instructions per word per layer), although in the
// 22 partial rounds eleven of the twelve words go from one linear layer straight into the next.  Here a word stays a limb
// triple (21 | 21 | 22 bits) between layers:
//   * partial rounds: words 1..11 are only carry-normalised (10 instructions: pl_normalise) -- the MDS grows a limb by
//     < 2^8.05, so limbs < 2^22 keep every intermediate of the shift/add transform < 2^31 (L1 norm of every intermediate
//     form <= 256, checked in tests/test_host_emulations.py) -- and their round constants are not added at all: they are
//     pushed forward through the linear layers into one scalar per round for word 0 and a correction of round 26's
//     constants (csrc/tools/gen_poseidon_limb_tables.py; the forward form of plonky2's FAST_PARTIAL_* constant pushing);
//   * word 0 of a partial round and all words of a full round are recombined to 64 bits for the S-box with a carry
//     chain that folds the top bits once (pl_recombine, 2^64 == 2^32 - 1), constants are added limb-wise before it;
//   * the 64x64 product is four independent IMAD.WIDE plus a six-instruction carry chain (gl_mul_nc4).
// Every carry step adds HB = 512 to the value it normalises (limb 0 += HB - h keeps limbs non-negative); the generated
// constants absorb it.  Same permutation, bit-exact (host emulation test + GPU parity tests).
// ---------------------------------------------------------------------------------------------------------------
#include "poseidon_limb_tables.cuh"
struct pl_c4 { u32 a, b, c, pad; };
static const pl_c4 h_POSEIDON_L_RC[96] = POSEIDON_L_RC_INIT;
static const pl_c4 h_POSEIDON_L_PK[22] = POSEIDON_L_PK_INIT;
#ifdef __CUDACC__
__constant__ pl_c4 d_POSEIDON_L_RC[96] = POSEIDON_L_RC_INIT;
__constant__ pl_c4 d_POSEIDON_L_PK[22] = POSEIDON_L_PK_INIT;
#endif
#ifndef PL_SBOX_UNROLL
#define PL_SBOX_UNROLL 2      /* S-boxes interleaved per iteration of the rolled S-box loop of a full round */
#endif
#define PL_PRAGMA_(x) _Pragma(#x)
#define PL_UNROLL(n) PL_PRAGMA_(unroll n)
#define PL_M21 0x1FFFFFu
#define PL_M22 0x3FFFFFu

// (a * b) mod p, any 64-bit inputs, result in [0, 2^64).  Device: 4 IMAD.WIDE.U32 without shifted addends (a shifted
// addend needs its 64-bit register pair rebuilt: two extra moves per product in the SASS of gl_mul_wide).
GL_HD u64 gl_mul_nc4(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u64 p00, p01, p10, p11;\n\t"
        ".reg .u32 x0, x1, x2, x3, a0, a1, b0, b1, l01, h01, l10, h10, l11, h11, h00, m0, m1, m2, t0l, t0h, bm, t1l, t1h, cm;\n\t"
        "mov.b64 {a0, a1}, %2;\n\t"
        "mov.b64 {b0, b1}, %3;\n\t"
        "mul.wide.u32 p00, a0, b0;\n\t"
        "mul.wide.u32 p01, a0, b1;\n\t"
        "mul.wide.u32 p10, a1, b0;\n\t"
        "mul.wide.u32 p11, a1, b1;\n\t"
        "mov.b64 {x0, h00}, p00;\n\t"
        "mov.b64 {l01, h01}, p01;\n\t"
        "mov.b64 {l10, h10}, p10;\n\t"
        "mov.b64 {l11, h11}, p11;\n\t"
        "add.cc.u32 m0, l01, l10;\n\t"          // mid = p01 + p10 (65 bits: m2:m1:m0)
        "addc.cc.u32 m1, h01, h10;\n\t"
        "addc.u32 m2, 0, 0;\n\t"
        "add.cc.u32 x1, h00, m0;\n\t"           // product = x3:x2:x1:x0
        "addc.cc.u32 x2, l11, m1;\n\t"
        "addc.u32 x3, h11, m2;\n\t"
        // reduce: x0 + x1*2^32 + x2*(2^32 - 1) - x3
        "sub.cc.u32 t0l, x0, x3;\n\t"
        "subc.cc.u32 t0h, x1, 0;\n\t"
        "subc.u32 bm, 0, 0;\n\t"                // borrow mask
        "sub.cc.u32 t0l, t0l, bm;\n\t"          // -= EPS on borrow
        "subc.u32 t0h, t0h, 0;\n\t"
        "sub.cc.u32 t1l, 0, x2;\n\t"            // t1 = (x2 << 32) - x2
        "subc.u32 t1h, x2, 0;\n\t"
        "add.cc.u32 %0, t0l, t1l;\n\t"
        "addc.cc.u32 %1, t0h, t1h;\n\t"
        "addc.u32 cm, 0, 0;\n\t"
        "neg.s32 cm, cm;\n\t"
        "add.cc.u32 %0, %0, cm;\n\t"            // += EPS on carry
        "addc.u32 %1, %1, 0;\n\t"
        "}" : "=&r"(r0), "=&r"(r1) : "l"(a), "l"(b));
    return ((u64)r1 << 32) | r0;
#else
    return gl_mul_nc(a, b);
#endif
}
// a^2 mod p with three wide multiplies (a0*a1 counted twice)
GL_HD u64 gl_sqr_nc4(u64 a) {
#ifdef __CUDA_ARCH__
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u64 p00, p01, p11;\n\t"
        ".reg .u32 x0, x1, x2, x3, a0, a1, l01, h01, l11, h11, h00, m0, m1, m2, t0l, t0h, bm, t1l, t1h, cm;\n\t"
        "mov.b64 {a0, a1}, %2;\n\t"
        "mul.wide.u32 p00, a0, a0;\n\t"
        "mul.wide.u32 p01, a0, a1;\n\t"
        "mul.wide.u32 p11, a1, a1;\n\t"
        "mov.b64 {x0, h00}, p00;\n\t"
        "mov.b64 {l01, h01}, p01;\n\t"
        "mov.b64 {l11, h11}, p11;\n\t"
        "shl.b32 m0, l01, 1;\n\t"                // mid = 2 * p01
        "shf.l.clamp.b32 m1, l01, h01, 1;\n\t"
        "shr.b32 m2, h01, 31;\n\t"
        "add.cc.u32 x1, h00, m0;\n\t"
        "addc.cc.u32 x2, l11, m1;\n\t"
        "addc.u32 x3, h11, m2;\n\t"
        "sub.cc.u32 t0l, x0, x3;\n\t"
        "subc.cc.u32 t0h, x1, 0;\n\t"
        "subc.u32 bm, 0, 0;\n\t"
        "sub.cc.u32 t0l, t0l, bm;\n\t"
        "subc.u32 t0h, t0h, 0;\n\t"
        "sub.cc.u32 t1l, 0, x2;\n\t"
        "subc.u32 t1h, x2, 0;\n\t"
        "add.cc.u32 %0, t0l, t1l;\n\t"
        "addc.cc.u32 %1, t0h, t1h;\n\t"
        "addc.u32 cm, 0, 0;\n\t"
        "neg.s32 cm, cm;\n\t"
        "add.cc.u32 %0, %0, cm;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "}" : "=&r"(r0), "=&r"(r1) : "l"(a));
    return ((u64)r1 << 32) | r0;
#else
    return gl_mul_nc(a, a);
#endif
}
GL_HD u64 poseidon_sbox4(u64 x) {
    u64 x2 = gl_sqr_nc4(x);
    u64 x4 = gl_sqr_nc4(x2);
    u64 x3 = gl_mul_nc4(x2, x);
    return gl_mul_nc4(x4, x3);
}

GL_HD void pl_split(u64 x, i32 &l0, i32 &l1, i32 &l2) {
    l0 = (i32)((u32)x & PL_M21); l1 = (i32)((u32)(x >> 21) & PL_M21); l2 = (i32)(u32)(x >> 42);
}
// limbs 0 <= y0, y1, y2 < 2^31 (y2 < 2^31 - 2^22)  ->  64-bit representative of y0 + y1*2^21 + y2*2^42 + HB  (mod p)
GL_HD u64 pl_recombine(u32 y0, u32 y1, u32 y2) {
    const u32 h = y2 >> 22;                       // weight 2^64 == 2^32 - 1
    y2 &= PL_M22; y1 += h << 11; y0 += POSEIDON_L_HB - h;
#ifdef __CUDA_ARCH__
    u32 r0, r1;
    asm("{\n\t"
        ".reg .u32 s0, s1, s2, cm;\n\t"
        "shl.b32 s0, %3, 21;\n\t"
        "shr.b32 s1, %3, 11;\n\t"
        "shl.b32 s2, %4, 10;\n\t"
        "add.cc.u32 %0, %2, s0;\n\t"
        "addc.u32 s1, s1, 0;\n\t"                 // y1 >> 11 < 2^20: no carry out
        "add.cc.u32 %1, s1, s2;\n\t"              // (y2 << 10) + ... may wrap 2^32 once
        "addc.u32 cm, 0, 0;\n\t"
        "neg.s32 cm, cm;\n\t"
        "add.cc.u32 %0, %0, cm;\n\t"              // += EPS on wrap (the wrapped high word is < 2^21: no second carry out)
        "addc.u32 %1, %1, 0;\n\t"
        "}" : "=&r"(r0), "=&r"(r1) : "r"(y0), "r"(y1), "r"(y2));
    return ((u64)r1 << 32) | r0;
#else
    u64 v = (u64)y0 + ((u64)y1 << 21);
    u64 t = (u64)y2 << 42;
    u64 r = v + t;
    if (r < t) r += GL_EPS;
    return r;
#endif
}
// carry-normalise an MDS output word in place: limbs back below (2^21 + HB, 2^21 + 2^20, 2^22); the value grows by HB
GL_HD void pl_normalise(i32 &y0, i32 &y1, i32 &y2) {
    const u32 t0 = (u32)y0, c0 = t0 >> 21;
    const u32 t1 = (u32)y1 + c0, c1 = t1 >> 21;
    const u32 t2 = (u32)y2 + c1, h = t2 >> 22;
    y2 = (i32)(t2 & PL_M22);
    y1 = (i32)((t1 & PL_M21) + (h << 11));
    y0 = (i32)((t0 & PL_M21) + POSEIDON_L_HB - h);
}
// MDS layer on limb triples, in place (diagonal term included)
GL_HD void pl_mds(i32 l0[12], i32 l1[12], i32 l2[12]) {
    i32 y0[12], y1[12], y2[12];
    poseidon_mds_limb(l0, y0); poseidon_mds_limb(l1, y1); poseidon_mds_limb(l2, y2);
    y0[0] += l0[0] << 3; y1[0] += l1[0] << 3; y2[0] += l2[0] << 3;
#pragma unroll
    for (int i = 0; i < 12; i++) { l0[i] = y0[i]; l1[i] = y1[i]; l2[i] = y2[i]; }
}

#ifdef __CUDA_ARCH__
#define PLT(name) d_##name
#else
#define PLT(name) h_##name
#endif
// The permutation on a strided state (device: shared memory, element-major; host: stride 1), in/out canonical.
GL_HD void poseidon_permute_limb(u64 *st, const int stride) {
    i32 l0[12], l1[12], l2[12];
#pragma unroll
    for (int i = 0; i < 12; i++) pl_split(st[i * stride], l0[i], l1[i], l2[i]);
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
            const pl_c4 *rc = PLT(POSEIDON_L_RC) + 12 * (4 * phase + r);
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const pl_c4 c = rc[i];
                st[i * stride] = pl_recombine((u32)l0[i] + c.a, (u32)l1[i] + c.b, (u32)l2[i] + c.c);
            }
PL_UNROLL(PL_SBOX_UNROLL)
            for (int i = 0; i < 12; i++) st[i * stride] = poseidon_sbox4(st[i * stride]);
#pragma unroll
            for (int i = 0; i < 12; i++) pl_split(st[i * stride], l0[i], l1[i], l2[i]);
            pl_mds(l0, l1, l2);
        }
        if (phase == 0) {
#pragma unroll
            for (int i = 1; i < 12; i++) pl_normalise(l0[i], l1[i], l2[i]);
#pragma unroll 1
            for (int r = 0; r < 22; r++) {
                const pl_c4 c = PLT(POSEIDON_L_PK)[r];
                const u64 x = poseidon_sbox4(pl_recombine((u32)l0[0] + c.a, (u32)l1[0] + c.b, (u32)l2[0] + c.c));
                pl_split(x, l0[0], l1[0], l2[0]);
                pl_mds(l0, l1, l2);
#pragma unroll
                for (int i = 1; i < 12; i++) pl_normalise(l0[i], l1[i], l2[i]);
            }
        }
    }
    const pl_c4 fin = POSEIDON_L_FIN_INIT;
#pragma unroll
    for (int i = 0; i < 12; i++) st[i * stride] = gl_canon(pl_recombine((u32)l0[i] + fin.a, (u32)l1[i] + fin.b, (u32)l2[i] + fin.c));
}

#ifdef __CUDACC__
// Device entry used by the Merkle / sponge kernels: the 12-word state of a thread lives in SHARED memory (element-major:
// word i of thread t at st[i * stride], conflict-free) so that the S-box layer of a full round is a rolled loop and the
// whole permutation stays ~1.5 k instructions (the round-1 register-resident, fully unrolled version was 15 k instructions
// and stalled on instruction fetch).  Out of line: one copy per kernel.
static __device__ __noinline__ void poseidon_permute_smem(u64 *st, const int stride) { poseidon_permute_limb(st, stride); }
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
