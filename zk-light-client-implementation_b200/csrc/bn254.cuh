// BN254 scalar field Fr (Montgomery, 4 x u64) and Poseidon-BN254 (t = 4, 8 full + 56 partial rounds, x^5, iden3
// optimised form) + the PoseidonBN128Hash sponge of PoseidonBN128GoldilocksConfig, host + device.
// Spec: crypto/plonky2_bn128/src/poseidon_bn128.rs:18-108 (permutation), config.rs:138-166 (hash_no_pad: 3 Goldilocks
// elements per Fr, little-endian bytes), :178-189 (hash_or_noop: <= 3 elements packed, no hash), :191-199 (two_to_one),
// :58-69 (to_vec: 32 LE bytes in 7-byte chunks for the challenger).
// Digests are stored exactly like Goldilocks digests: 4 u64 = the 32 little-endian bytes of the canonical value, so
// Merkle buffers, caps and the proof serializer are hasher-agnostic.
#pragma once
#include "gl.cuh"
#include "poseidon_bn254_tables.cuh"
#include "mont256.cuh"

struct fr_t { u64 v[4]; };

#define X(name, n, ...) static const u64 h_##name[n][4] = {__VA_ARGS__};
POSEIDON_BN254_TABLES(X)
#undef X
#ifdef __CUDACC__
#define X(name, n, ...) __constant__ u64 d_##name[n][4] = {__VA_ARGS__};
POSEIDON_BN254_TABLES(X)
#undef X
#endif
#ifdef __CUDA_ARCH__
#define BT(name) d_##name
#else
#define BT(name) h_##name
#endif

#define FR_M0 0x43e1f593f0000001ULL
#define FR_M1 0x2833e84879b97091ULL
#define FR_M2 0xb85045b68181585dULL
#define FR_M3 0x30644e72e131a029ULL
#define FR_INV 0xc2e1f593efffffffULL          // -r^{-1} mod 2^64

struct FrMod { static constexpr u64 M0 = FR_M0, M1 = FR_M1, M2 = FR_M2, M3 = FR_M3, INV = FR_INV; };
GL_HD u64 fr_mac(u64 a, u64 b, u64 c, u64 &carry) {      // a*b + c + carry -> low word, carry = high word
    u64 lo, hi; gl_mul_wide(a, b, lo, hi);
    lo += c; hi += (lo < c);
    lo += carry; hi += (lo < carry);
    carry = hi; return lo;
}
GL_HD bool fr_geq_mod(const u64 a[4]) {
    const u64 m[4] = {FR_M0, FR_M1, FR_M2, FR_M3};
    for (int i = 3; i >= 0; i--) { if (a[i] > m[i]) return true; if (a[i] < m[i]) return false; }
    return true;
}
GL_HD void fr_sub_mod(u64 a[4]) {
    const u64 m[4] = {FR_M0, FR_M1, FR_M2, FR_M3};
    u64 br = 0;
    for (int i = 0; i < 4; i++) { u64 d = a[i] - m[i]; u64 b1 = a[i] < m[i]; u64 d2 = d - br; u64 b2 = d < br; a[i] = d2; br = b1 | b2; }
}
GL_HD fr_t fr_add(const fr_t &a, const fr_t &b) {
#ifdef __CUDA_ARCH__
    fr_t rr; mont256_add<FrMod>(a.v, b.v, rr.v); return rr;
#endif
    fr_t r; u64 c = 0;
    for (int i = 0; i < 4; i++) { u64 s = a.v[i] + b.v[i]; u64 c1 = s < a.v[i]; u64 s2 = s + c; u64 c2 = s2 < s; r.v[i] = s2; c = c1 | c2; }
    if (c || fr_geq_mod(r.v)) fr_sub_mod(r.v);
    return r;
}
GL_HD fr_t fr_sub(const fr_t &a, const fr_t &b) {
#ifdef __CUDA_ARCH__
    fr_t rr; mont256_sub<FrMod>(a.v, b.v, rr.v); return rr;
#endif
    const u64 m[4] = {FR_M0, FR_M1, FR_M2, FR_M3};
    fr_t r; u64 br = 0;
    for (int i = 0; i < 4; i++) { u64 d = a.v[i] - b.v[i]; u64 b1 = a.v[i] < b.v[i]; u64 d2 = d - br; u64 b2 = d < br; r.v[i] = d2; br = b1 | b2; }
    if (br) { u64 c = 0; for (int i = 0; i < 4; i++) { u64 s = r.v[i] + m[i]; u64 c1 = s < r.v[i]; u64 s2 = s + c; u64 c2 = s2 < s; r.v[i] = s2; c = c1 | c2; } }
    return r;
}
// On the device fr_mul is ONE out-of-line function (the first build inlined it ~900 times: 25 k instructions per kernel,
// instruction-fetch bound); loops over it stay rolled.
#ifdef __CUDACC__
#define BN_FN static __host__ __device__ __noinline__
#define BN_ROLL _Pragma("unroll 1")
#else
#define BN_FN static inline
#define BN_ROLL
#endif
BN_FN fr_t fr_mul(const fr_t &a, const fr_t &b) {        // Montgomery product a * b / 2^256
#ifdef __CUDA_ARCH__
    fr_t rr; mont256_mul<FrMod>(a.v, b.v, rr.v); return rr;   // 32-bit-limb even/odd CIOS (mont256.cuh): 250 instructions, 136 wide multiplies
#endif
    const u64 m[4] = {FR_M0, FR_M1, FR_M2, FR_M3};
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u64 c = 0;
        for (int j = 0; j < 4; j++) t[j] = fr_mac(a.v[j], b.v[i], t[j], c);
        { u64 s = t[4] + c; t[5] = s < c; t[4] = s; }
        u64 q = t[0] * FR_INV;
        c = 0;
        (void)fr_mac(q, m[0], t[0], c);
        for (int j = 1; j < 4; j++) t[j - 1] = fr_mac(q, m[j], t[j], c);
        { u64 s = t[4] + c; u64 cc = s < c; t[3] = s; t[4] = t[5] + cc; }
    }
    fr_t r; for (int i = 0; i < 4; i++) r.v[i] = t[i];
    if (t[4] || fr_geq_mod(r.v)) fr_sub_mod(r.v);
    return r;
}
GL_HD fr_t fr_load(const u64 (*tab)[4], int i) { fr_t r; for (int k = 0; k < 4; k++) r.v[k] = tab[i][k]; return r; }
GL_HD fr_t fr_from_canonical(const u64 x[4]) {           // x * R^2 * R^-1 = x R
    const fr_t r2 = {{0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL}};
    fr_t a; for (int i = 0; i < 4; i++) a.v[i] = x[i];
    return fr_mul(a, r2);
}
GL_HD void fr_to_canonical(const fr_t &a, u64 out[4]) {
    const fr_t one = {{1, 0, 0, 0}};
    fr_t r = fr_mul(a, one);
    for (int i = 0; i < 4; i++) out[i] = r.v[i];
}
BN_FN fr_t fr_exp5(const fr_t &x) { fr_t x2 = fr_mul(x, x); return fr_mul(fr_mul(x2, x2), x); }

BN_FN void bn254_mix(fr_t s[4], const u64 (*m)[4]) {     // result[i] = sum_j m[j][i] * s[j]
    fr_t r[4];
    BN_ROLL
    for (int i = 0; i < 4; i++) {
        fr_t acc = fr_mul(fr_load(m, i), s[0]);
        BN_ROLL
        for (int j = 1; j < 4; j++) acc = fr_add(acc, fr_mul(fr_load(m, 4 * j + i), s[j]));
        r[i] = acc;
    }
    for (int i = 0; i < 4; i++) s[i] = r[i];
}
// state in Montgomery form
BN_FN void bn254_permute(fr_t s[4]) {
    for (int i = 0; i < 4; i++) s[i] = fr_add(s[i], fr_load(BT(BN254_C), i));
    BN_ROLL
    for (int r = 0; r < 3; r++) {
        BN_ROLL
        for (int i = 0; i < 4; i++) s[i] = fr_add(fr_exp5(s[i]), fr_load(BT(BN254_C), (r + 1) * 4 + i));
        bn254_mix(s, BT(BN254_M));
    }
    BN_ROLL
    for (int i = 0; i < 4; i++) s[i] = fr_add(fr_exp5(s[i]), fr_load(BT(BN254_C), 16 + i));
    bn254_mix(s, BT(BN254_P));
    BN_ROLL
    for (int r = 0; r < 56; r++) {
        s[0] = fr_add(fr_exp5(s[0]), fr_load(BT(BN254_C), 20 + r));
        fr_t n0 = fr_mul(fr_load(BT(BN254_S), 7 * r), s[0]);
        BN_ROLL
        for (int j = 1; j < 4; j++) n0 = fr_add(n0, fr_mul(fr_load(BT(BN254_S), 7 * r + j), s[j]));
        BN_ROLL
        for (int k = 1; k < 4; k++) s[k] = fr_add(s[k], fr_mul(s[0], fr_load(BT(BN254_S), 7 * r + 3 + k)));
        s[0] = n0;
    }
    BN_ROLL
    for (int r = 0; r < 3; r++) {
        BN_ROLL
        for (int i = 0; i < 4; i++) s[i] = fr_add(fr_exp5(s[i]), fr_load(BT(BN254_C), 76 + r * 4 + i));
        bn254_mix(s, BT(BN254_M));
    }
    BN_ROLL
    for (int i = 0; i < 4; i++) s[i] = fr_exp5(s[i]);
    bn254_mix(s, BT(BN254_M));
}
GL_HD void bn254_two_to_one(const u64 l[4], const u64 r[4], u64 out[4]) {
    fr_t s[4]; for (int i = 0; i < 4; i++) { s[0].v[i] = 0; s[1].v[i] = 0; }
    s[2] = fr_from_canonical(l); s[3] = fr_from_canonical(r);
    bn254_permute(s);
    fr_to_canonical(s[0], out);
}
// the 32 LE bytes of a digest in 7-byte chunks -> 5 Goldilocks elements (challenger observation, config.rs:58-69)
inline void bn254_to_vec_host(const u64 h[4], u64 out[5]) {
    unsigned char b[32];
    for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) b[8 * i + k] = (unsigned char)(h[i] >> (8 * k));
    for (int c = 0; c < 5; c++) { u64 v = 0; for (int j = 0; j < 7 && 7 * c + j < 32; j++) v |= (u64)b[7 * c + j] << (8 * j); out[c] = v; }
}
