/* C oracle: Poseidon over BN254 Fr (t = 4, 8 full + 56 partial rounds, x^5, iden3 optimised form) and the
 * PoseidonBN128Hash sponge used by PoseidonBN128GoldilocksConfig.  TEST INFRASTRUCTURE.
 * Restates crypto/plonky2_bn128/src/poseidon_bn128.rs:18-108 (permutation) and config.rs:138-199 (hash_no_pad,
 * hash_or_noop, two_to_one).  Digests are exchanged as 4 little-endian u64 limbs of the canonical value (= the 32
 * bytes the reference serializes). */
#include "gl.h"
#include "poseidon_bn254_constants.h"
#include <string.h>

typedef struct { u64 v[4]; } fr;
static const u64 FR_MOD[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const u64 FR_INV = 0xc2e1f593efffffffULL;     /* -r^{-1} mod 2^64 */
static const u64 FR_R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL, 0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};   /* 2^512 mod r */

static int fr_geq(const u64 a[4], const u64 b[4]) { for (int i = 3; i >= 0; i--) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; } return 1; }
static void fr_sub_mod(u64 a[4]) { u128 br = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)a[i] - FR_MOD[i] - br; a[i] = (u64)d; br = (d >> 64) & 1; } }
static fr fr_add(fr a, fr b) {
    fr r; u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a.v[i] + b.v[i]; r.v[i] = (u64)c; c >>= 64; }
    if (c || fr_geq(r.v, FR_MOD)) fr_sub_mod(r.v);
    return r;
}
static fr fr_mul(fr a, fr b) {      /* Montgomery CIOS */
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a.v[j] * b.v[i] + t[j]; t[j] = (u64)c; c >>= 64; }
        c += t[4]; t[4] = (u64)c; t[5] = (u64)(c >> 64);
        u64 m = t[0] * FR_INV;
        c = ((u128)m * FR_MOD[0] + t[0]) >> 64;
        for (int j = 1; j < 4; j++) { c += (u128)m * FR_MOD[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
        c += t[4]; t[3] = (u64)c; t[4] = t[5] + (u64)(c >> 64);
    }
    fr r; memcpy(r.v, t, 32);
    if (t[4] || fr_geq(r.v, FR_MOD)) fr_sub_mod(r.v);
    return r;
}
static fr fr_from_limbs(const u64 x[4]) { fr a, r2; memcpy(a.v, x, 32); memcpy(r2.v, FR_R2, 32); return fr_mul(a, r2); }   /* to Montgomery */
static void fr_to_limbs(fr a, u64 out[4]) { fr one = {{1, 0, 0, 0}}; fr r = fr_mul(a, one); memcpy(out, r.v, 32); }

static fr TC[88], TS[392], TM[16], TP[16]; static int tables_ready = 0;
static void init_tables(void) {
    #pragma omp critical
    {
        if (!tables_ready) {
            for (int i = 0; i < 88; i++) TC[i] = fr_from_limbs(PB_C[i]);
            for (int i = 0; i < 392; i++) TS[i] = fr_from_limbs(PB_S[i]);
            for (int i = 0; i < 16; i++) { TM[i] = fr_from_limbs(PB_M[i]); TP[i] = fr_from_limbs(PB_P[i]); }
            tables_ready = 1;
        }
    }
}
static fr exp5(fr x) { fr x2 = fr_mul(x, x); return fr_mul(fr_mul(x2, x2), x); }
static void mix(fr s[4], const fr *m) {      /* result[i] = sum_j m[j][i] * s[j] */
    fr r[4];
    for (int i = 0; i < 4; i++) { fr acc = fr_mul(m[0 * 4 + i], s[0]); for (int j = 1; j < 4; j++) acc = fr_add(acc, fr_mul(m[j * 4 + i], s[j])); r[i] = acc; }
    memcpy(s, r, sizeof r);
}
static void permute_mont(fr s[4]) {
    if (!tables_ready) init_tables();
    for (int i = 0; i < 4; i++) s[i] = fr_add(s[i], TC[i]);
    for (int r = 0; r < 3; r++) { for (int i = 0; i < 4; i++) s[i] = fr_add(exp5(s[i]), TC[(r + 1) * 4 + i]); mix(s, TM); }
    for (int i = 0; i < 4; i++) s[i] = fr_add(exp5(s[i]), TC[16 + i]);
    mix(s, TP);
    for (int r = 0; r < 56; r++) {
        s[0] = fr_add(exp5(s[0]), TC[20 + r]);
        fr n0 = fr_mul(TS[7 * r], s[0]);
        for (int j = 1; j < 4; j++) n0 = fr_add(n0, fr_mul(TS[7 * r + j], s[j]));
        for (int k = 1; k < 4; k++) s[k] = fr_add(s[k], fr_mul(s[0], TS[7 * r + 3 + k]));
        s[0] = n0;
    }
    for (int r = 0; r < 3; r++) { for (int i = 0; i < 4; i++) s[i] = fr_add(exp5(s[i]), TC[76 + r * 4 + i]); mix(s, TM); }
    for (int i = 0; i < 4; i++) s[i] = exp5(s[i]);
    mix(s, TM);
}
/* canonical limbs in / out (test entry point; G3 KAT) */
void orc_bn254_permute(u64 st[16]) {
    fr s[4]; for (int i = 0; i < 4; i++) s[i] = fr_from_limbs(st + 4 * i);
    permute_mont(s);
    for (int i = 0; i < 4; i++) fr_to_limbs(s[i], st + 4 * i);
}
/* hash_no_pad over Goldilocks elements read with a stride (column-major leaves): 3 elements per Fr, 3 Fr per permutation */
void orc_bn254_hash_strided(const u64 *in, size_t n, size_t stride, u64 out[4]) {
    fr s[4]; memset(s, 0, sizeof s);
    for (size_t i = 0; i < n; i += 9) {
        size_t m = n - i < 9 ? n - i : 9;
        for (size_t j = 0; j < m; j += 3) {
            u64 l[4] = {0, 0, 0, 0};
            for (size_t k = 0; k < 3 && j + k < m; k++) l[k] = in[(i + j + k) * stride];
            s[j / 3 + 1] = fr_from_limbs(l);
        }
        permute_mont(s);
    }
    fr_to_limbs(s[0], out);
}
void orc_bn254_hash_or_noop_strided(const u64 *in, size_t n, size_t stride, u64 out[4]) {
    if (n <= 3) { memset(out, 0, 32); for (size_t k = 0; k < n; k++) out[k] = in[k * stride]; }
    else orc_bn254_hash_strided(in, n, stride, out);
}
void orc_bn254_two_to_one(const u64 l[4], const u64 r[4], u64 out[4]) {
    fr s[4]; memset(s, 0, sizeof s);
    s[2] = fr_from_limbs(l); s[3] = fr_from_limbs(r);
    permute_mont(s);
    fr_to_limbs(s[0], out);
}
