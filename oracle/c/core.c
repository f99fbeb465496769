/* C oracle core: Poseidon-GL, sponge, Merkle caps, radix-2 NTT / coset LDE.  TEST INFRASTRUCTURE.
 * Follows:
 *   gnark-plonky2-verifier/poseidon/goldilocks.go:30-37,92-115,251-331   permutation (fast partial rounds)
 *   goldilocks.go:41-86                                                   hash_n_to_m_no_pad
 *   upstream plonky2 (wormhole-foundation/plonky2-near @ 2244a9d, not in tree): hash_or_noop, two_to_one,
 *   MerkleTree::new (binary tree, cap = level with 2^cap_height nodes), PolynomialBatch::from_values
 *   (ifft -> zero-pad x8 -> coset fft shift 7 -> bit-reversed rows)            [SURVEY.md 3.2 steps 3,6,9]
 */
#include "gl.h"
#include "poseidon_constants.h"
#include <stdlib.h>
#include <string.h>
#include <omp.h>

static inline u64 sbox7(u64 x) { u64 x2 = gl_mul(x, x), x4 = gl_mul(x2, x2), x3 = gl_mul(x2, x); return gl_mul(x4, x3); }

static void mds_layer(u64 s[12]) {
    u64 o[12];
    for (int r = 0; r < 12; r++) {
        u128 acc = 0;
        for (int i = 0; i < 12; i++) acc += (u128)s[(i + r) % 12] * PG_MDS_CIRC[i];
        acc += (u128)s[r] * PG_MDS_DIAG[r];
        o[r] = gl_red128(acc);
    }
    memcpy(s, o, sizeof o);
}

void orc_poseidon_permute(u64 s[12]) {
    int rc = 0;
    for (int r = 0; r < 4; r++, rc++) {
        for (int i = 0; i < 12; i++) s[i] = sbox7(gl_add(s[i], PG_RC[12 * rc + i] % GL_P));
        mds_layer(s);
    }
    for (int i = 0; i < 12; i++) s[i] = gl_add(s[i], PG_FP_FIRST_RC[i] % GL_P);
    {   /* mds_partial_layer_init */
        u64 o[12]; o[0] = s[0];
        for (int d = 1; d < 12; d++) {
            u64 acc = 0;
            for (int r = 1; r < 12; r++) acc = gl_add(acc, gl_mul(s[r], PG_FP_INIT[11 * (r - 1) + (d - 1)] % GL_P));
            o[d] = acc;
        }
        memcpy(s, o, sizeof o);
    }
    for (int r = 0; r < 22; r++) {
        s[0] = gl_add(sbox7(s[0]), PG_FP_RC[r] % GL_P);
        u64 d = gl_mul(s[0], 25);
        for (int i = 1; i < 12; i++) d = gl_add(d, gl_mul(s[i], PG_FP_W_HATS[11 * r + i - 1] % GL_P));
        for (int i = 1; i < 12; i++) s[i] = gl_add(s[i], gl_mul(s[0], PG_FP_VS[11 * r + i - 1] % GL_P));
        s[0] = d;
    }
    rc += 22;
    for (int r = 0; r < 4; r++, rc++) {
        for (int i = 0; i < 12; i++) s[i] = sbox7(gl_add(s[i], PG_RC[12 * rc + i] % GL_P));
        mds_layer(s);
    }
}

void orc_hash_no_pad(const u64 *in, size_t n, u64 out[4]) {
    u64 s[12] = {0};
    for (size_t i = 0; i < n; i += 8) {
        size_t m = n - i < 8 ? n - i : 8;
        for (size_t j = 0; j < m; j++) s[j] = in[i + j];
        orc_poseidon_permute(s);
    }
    memcpy(out, s, 4 * sizeof(u64));
}

void orc_hash_or_noop(const u64 *in, size_t n, u64 out[4]) {
    if (n <= 4) { memset(out, 0, 32); memcpy(out, in, n * 8); }
    else orc_hash_no_pad(in, n, out);
}

void orc_two_to_one(const u64 l[4], const u64 r[4], u64 out[4]) {
    u64 s[12] = { l[0], l[1], l[2], l[3], r[0], r[1], r[2], r[3], 0, 0, 0, 0 };
    orc_poseidon_permute(s);
    memcpy(out, s, 32);
}

/* Strided sponge: leaf `row` of a column-major matrix [cols][L] (element (c,row) at base[c*L+row]). */
static void hash_leaf_colmajor(const u64 *base, size_t cols, size_t L, size_t row, u64 out[4]) {
    if (cols <= 4) { memset(out, 0, 32); for (size_t c = 0; c < cols; c++) out[c] = base[c * L + row]; return; }
    u64 s[12] = {0};
    for (size_t c = 0; c < cols; c += 8) {
        size_t m = cols - c < 8 ? cols - c : 8;
        for (size_t j = 0; j < m; j++) s[j] = base[(c + j) * L + row];
        orc_poseidon_permute(s);
    }
    memcpy(out, s, 32);
}

/* Merkle tree over L = 2^log_L leaves given column-major data.  digests: level 0 (L x 4), level 1 (L/2 x 4), ...
 * down to the cap level (2^cap_height x 4); total 4*(2L - 2^cap_height) u64.  Returns pointer to cap inside digests. */
size_t orc_merkle_num_u64(int log_L, int cap_height) { return 4 * ((2ULL << log_L) - (1ULL << cap_height)); }

/* hasher: 0 = PoseidonHash (Goldilocks), 1 = PoseidonBN128Hash (oracle/c/bn254.c) */
void orc_bn254_hash_or_noop_strided(const u64 *in, size_t n, size_t stride, u64 out[4]);
void orc_bn254_two_to_one(const u64 l[4], const u64 r[4], u64 out[4]);

void orc_merkle_from_digests_h(u64 *digests, int log_L, int cap_height, int hasher) {
    u64 *cur = digests;
    for (int lv = log_L; lv > cap_height; lv--) {
        size_t n = 1ULL << lv;
        u64 *nxt = cur + 4 * n;
        #pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n / 2; i++) {
            if (hasher) orc_bn254_two_to_one(cur + 8 * i, cur + 8 * i + 4, nxt + 4 * i);
            else orc_two_to_one(cur + 8 * i, cur + 8 * i + 4, nxt + 4 * i);
        }
        cur = nxt;
    }
}
void orc_merkle_colmajor_h(const u64 *data, size_t cols, int log_L, int cap_height, u64 *digests, int hasher) {
    size_t L = 1ULL << log_L;
    #pragma omp parallel for schedule(static)
    for (size_t r = 0; r < L; r++) {
        if (hasher) orc_bn254_hash_or_noop_strided(data + r, cols, L, digests + 4 * r);
        else hash_leaf_colmajor(data, cols, L, r, digests + 4 * r);
    }
    orc_merkle_from_digests_h(digests, log_L, cap_height, hasher);
}
void orc_merkle_rowmajor_h(const u64 *data, size_t width, int log_L, int cap_height, u64 *digests, int hasher) {
    size_t L = 1ULL << log_L;
    #pragma omp parallel for schedule(static)
    for (size_t r = 0; r < L; r++) {
        if (hasher) orc_bn254_hash_or_noop_strided(data + r * width, width, 1, digests + 4 * r);
        else orc_hash_or_noop(data + r * width, width, digests + 4 * r);
    }
    orc_merkle_from_digests_h(digests, log_L, cap_height, hasher);
}

void orc_merkle_from_digests(u64 *digests, int log_L, int cap_height) {
    size_t L = 1ULL << log_L;
    u64 *cur = digests;
    for (int lv = log_L; lv > cap_height; lv--) {
        size_t n = 1ULL << lv;
        u64 *nxt = cur + 4 * n;
        #pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n / 2; i++) orc_two_to_one(cur + 8 * i, cur + 8 * i + 4, nxt + 4 * i);
        cur = nxt;
    }
    (void)L;
}

void orc_merkle_colmajor(const u64 *data, size_t cols, int log_L, int cap_height, u64 *digests) {
    size_t L = 1ULL << log_L;
    #pragma omp parallel for schedule(static)
    for (size_t r = 0; r < L; r++) hash_leaf_colmajor(data, cols, L, r, digests + 4 * r);
    orc_merkle_from_digests(digests, log_L, cap_height);
}

/* leaves given row-major: leaf r = data[r*width .. r*width+width) */
void orc_merkle_rowmajor(const u64 *data, size_t width, int log_L, int cap_height, u64 *digests) {
    size_t L = 1ULL << log_L;
    #pragma omp parallel for schedule(static)
    for (size_t r = 0; r < L; r++) orc_hash_or_noop(data + r * width, width, digests + 4 * r);
    orc_merkle_from_digests(digests, log_L, cap_height);
}

/* ---------- NTT: natural in -> natural out, in place ---------- */
void orc_ntt(u64 *a, int log_n, int inverse) {
    size_t n = 1ULL << log_n;
    /* a single large transform called outside a parallel region (the extension-field FRI polynomials) spreads its
     * butterflies over the team; inside the per-column batches the threads are already busy */
    const int par = !omp_in_parallel() && log_n >= 14;
    #pragma omp parallel for schedule(static) if(par)
    for (size_t i = 0; i < n; i++) { size_t j = bitrev(i, log_n); if (i < j) { u64 t = a[i]; a[i] = a[j]; a[j] = t; } }
    u64 w_n = gl_root(log_n);
    if (inverse) w_n = gl_inv(w_n);
    for (int s = 1; s <= log_n; s++) {
        size_t m = 1ULL << s, h = m >> 1;
        u64 wm = gl_pow(w_n, n >> s);
        u64 *tw = (u64 *)malloc(h * sizeof(u64));
        tw[0] = 1; for (size_t k = 1; k < h; k++) tw[k] = gl_mul(tw[k - 1], wm);
        #pragma omp parallel for schedule(static) if(par)
        for (size_t b = 0; b < n / 2; b++) {
            size_t j = b & (h - 1), k = (b >> (s - 1)) << s;
            u64 t = gl_mul(tw[j], a[k + j + h]), u = a[k + j];
            a[k + j] = gl_add(u, t); a[k + j + h] = gl_sub(u, t);
        }
        free(tw);
    }
    if (inverse) { u64 ninv = gl_inv(n % GL_P); for (size_t i = 0; i < n; i++) a[i] = gl_mul(a[i], ninv); }
}

/* batch over columns: data [cols][n] column-major, each column transformed in place */
void orc_ntt_batch(u64 *data, size_t cols, int log_n, int inverse) {
    size_t n = 1ULL << log_n;
    #pragma omp parallel for schedule(dynamic)
    for (size_t c = 0; c < cols; c++) orc_ntt(data + c * n, log_n, inverse);
}

/* coset LDE of one polynomial: coeffs (n) -> out (n << rate_bits) values at shift*w^i, stored at index bitrev(i). */
void orc_lde(const u64 *coeffs, int log_n, int rate_bits, u64 shift, u64 *out) {
    size_t n = 1ULL << log_n, N = n << rate_bits;
    u64 *tmp = (u64 *)calloc(N, sizeof(u64));
    u64 s = 1;
    for (size_t i = 0; i < n; i++) { tmp[i] = gl_mul(coeffs[i], s); s = gl_mul(s, shift); }
    orc_ntt(tmp, log_n + rate_bits, 0);
    for (size_t i = 0; i < N; i++) out[bitrev(i, log_n + rate_bits)] = tmp[i];
    free(tmp);
}

void orc_lde_batch(const u64 *coeffs, size_t cols, int log_n, int rate_bits, u64 shift, u64 *out) {
    size_t n = 1ULL << log_n, N = n << rate_bits;
    #pragma omp parallel for schedule(dynamic)
    for (size_t c = 0; c < cols; c++) orc_lde(coeffs + c * n, log_n, rate_bits, shift, out + c * N);
}

/* PolynomialBatch::from_values restated: values [cols][n] -> coeffs [cols][n], lde [cols][8n] (bit-reversed rows),
 * digests (all levels), cap = last 2^cap_height digests. */
void orc_commit_values(const u64 *values, size_t cols, int log_n, int rate_bits, int cap_height,
                       u64 *coeffs, u64 *lde, u64 *digests) {
    size_t n = 1ULL << log_n;
    memcpy(coeffs, values, cols * n * sizeof(u64));
    orc_ntt_batch(coeffs, cols, log_n, 1);
    orc_lde_batch(coeffs, cols, log_n, rate_bits, GL_GEN, lde);
    orc_merkle_colmajor(lde, cols, log_n + rate_bits, cap_height, digests);
}
