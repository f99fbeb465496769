/* C oracle: a complete plonky2 prover on the CPU (PoseidonGoldilocksConfig, D = 2, no zero-knowledge, no lookups).
 * TEST INFRASTRUCTURE -- the checker for the CUDA prover (byte-identical proofs expected) and bench.py's cpu_baseline.
 *
 * Restates upstream plonky2 (wormhole-foundation/plonky2-near @ 2244a9d, NOT in /root/reference):
 *   plonk/prover.rs::prove (steps 2-12), plonk/vanishing_poly.rs::eval_vanishing_poly_base_batch,
 *   plonk/plonk_common.rs (ZeroPolyOnCoset, reduce_with_powers), fri/oracle.rs::{from_values, prove_openings},
 *   fri/prover.rs::{fri_committed_trees, fri_proof_of_work (minimum nonce), fri_prover_query_rounds},
 *   util/serialization (proof bytes), iop/challenger.rs.
 * Everything a verifier can see is pinned through oracle/verifier.py (which reproduces the reference's golden
 * proofs); prover-private choices follow the upstream algorithm literally (coefficient-space FRI folding etc.),
 * deliberately differently from the CUDA implementation so that agreement is meaningful.
 */
#include "gl.h"
#include "gates.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <omp.h>
/* phase wall clock to stderr when ORC_TIMING is set (test infrastructure only) */
static double orc_t0_; static int orc_timing_ = -1;
#define ORC_TICK(name) do { if (orc_timing_ < 0) orc_timing_ = getenv("ORC_TIMING") != NULL; if (orc_timing_) { double t_ = omp_get_wtime(); fprintf(stderr, "[orc] %-22s %8.3f s\n", name, t_ - orc_t0_); orc_t0_ = t_; } } while (0)

void orc_poseidon_permute(u64 s[12]);
void orc_hash_no_pad(const u64 *in, size_t n, u64 out[4]);
void orc_ntt(u64 *a, int log_n, int inverse);
void orc_ntt_batch(u64 *data, size_t cols, int log_n, int inverse);
void orc_lde_batch(const u64 *coeffs, size_t cols, int log_n, int rate_bits, u64 shift, u64 *out);
size_t orc_merkle_num_u64(int log_L, int cap_height);
void orc_merkle_colmajor(const u64 *data, size_t cols, int log_L, int cap_height, u64 *digests);
void orc_merkle_rowmajor(const u64 *data, size_t width, int log_L, int cap_height, u64 *digests);
void orc_merkle_colmajor_h(const u64 *data, size_t cols, int log_L, int cap_height, u64 *digests, int hasher);
void orc_merkle_rowmajor_h(const u64 *data, size_t width, int log_L, int cap_height, u64 *digests, int hasher);

typedef struct {
    uint32_t num_wires, num_routed, num_constants, num_challenges, degree_bits, rate_bits, cap_height, pow_bits,
             num_queries, qdf, npp, num_gate_constraints, num_pi, num_selectors, n_arity;
    uint32_t arity_bits[16];
    uint32_t num_gates;
    const orc_gate *gates;
    const u64 *k_is;
    uint32_t hasher;      /* 0 = PoseidonGoldilocksConfig, 1 = PoseidonBN128GoldilocksConfig */
} orc_common;

/* ---------------- challenger (challenger.go:42-165) ---------------- */
typedef struct { u64 st[12]; u64 in[8]; int nin; u64 out[8]; int nout; } chal;
static void ch_init(chal *c) { memset(c, 0, sizeof *c); }
static void ch_duplex(chal *c) {
    for (int i = 0; i < c->nin; i++) c->st[i] = c->in[i];
    c->nin = 0;
    orc_poseidon_permute(c->st);
    for (int i = 0; i < 8; i++) c->out[i] = c->st[i];
    c->nout = 8;
}
static void ch_observe(chal *c, u64 x) { c->nout = 0; c->in[c->nin++] = x; if (c->nin == 8) ch_duplex(c); }
static void ch_observe_n(chal *c, const u64 *x, size_t n) { for (size_t i = 0; i < n; i++) ch_observe(c, x[i]); }
static u64 ch_get(chal *c) { if (c->nin || !c->nout) ch_duplex(c); return c->out[--c->nout]; }
static e2 ch_get_ext(chal *c) { e2 r; r.a = ch_get(c); r.b = ch_get(c); return r; }
/* observe a Merkle/sponge hash: GL = its 4 elements; BN128 = the 32 LE bytes in 7-byte chunks (config.rs:58-69) */
static void ch_observe_hash(chal *c, const u64 h[4], int hasher) {
    if (!hasher) { ch_observe_n(c, h, 4); return; }
    uint8_t b[32]; memcpy(b, h, 32);
    for (int i = 0; i < 32; i += 7) { u64 v = 0; for (int j = 0; j < 7 && i + j < 32; j++) v |= (u64)b[i + j] << (8 * j); ch_observe(c, v); }
}
static void ch_observe_cap(chal *c, const u64 *cap, size_t n_hashes, int hasher) { for (size_t i = 0; i < n_hashes; i++) ch_observe_hash(c, cap + 4 * i, hasher); }

/* ---------------- byte writer ---------------- */
typedef struct { uint8_t *p; size_t len, cap; } wbuf;
static void wb_put(wbuf *w, const void *d, size_t n) {
    if (w->len + n > w->cap) { w->cap = (w->len + n) * 2 + 1024; w->p = (uint8_t *)realloc(w->p, w->cap); }
    memcpy(w->p + w->len, d, n); w->len += n;
}
static void wb_u64(wbuf *w, u64 x) { wb_put(w, &x, 8); }
static void wb_u8(wbuf *w, uint8_t x) { wb_put(w, &x, 1); }

/* ---------------- commitment bundle ---------------- */
typedef struct { size_t cols; int log_n, rate_bits, cap_height, hasher; u64 *coeffs, *lde, *digests; } batch;
static const u64 *batch_cap(const batch *b) {
    return b->digests + orc_merkle_num_u64(b->log_n + b->rate_bits, b->cap_height) - (4ULL << b->cap_height);
}
static void batch_from_coeffs(batch *b, size_t cols, int log_n, int rate_bits, int cap_height, int hasher) {
    size_t n = 1ULL << log_n;
    b->hasher = hasher; b->cols = cols; b->log_n = log_n; b->rate_bits = rate_bits; b->cap_height = cap_height;
    b->lde = (u64 *)malloc(cols * (n << rate_bits) * 8);
    b->digests = (u64 *)malloc(orc_merkle_num_u64(log_n + rate_bits, cap_height) * 8);
    orc_lde_batch(b->coeffs, cols, log_n, rate_bits, GL_GEN, b->lde);
    orc_merkle_colmajor_h(b->lde, cols, log_n + rate_bits, cap_height, b->digests, hasher);
}
static void batch_from_values(batch *b, const u64 *values, size_t cols, int log_n, int rate_bits, int cap_height, int hasher) {
    size_t n = 1ULL << log_n;
    b->coeffs = (u64 *)malloc(cols * n * 8);
    memcpy(b->coeffs, values, cols * n * 8);
    orc_ntt_batch(b->coeffs, cols, log_n, 1);
    batch_from_coeffs(b, cols, log_n, rate_bits, cap_height, hasher);
}
static void batch_free(batch *b) { free(b->coeffs); free(b->lde); free(b->digests); memset(b, 0, sizeof *b); }

/* opaque circuit handle */
typedef struct { orc_common c; batch cs; u64 *sigma_values; u64 digest[4]; } orc_circuit;

orc_circuit *orc_circuit_new(const orc_common *c, const u64 *const_sigma_values, const u64 digest[4], u64 *cap_out) {
    orc_circuit *k = (orc_circuit *)calloc(1, sizeof *k);
    k->c = *c;
    size_t n = 1ULL << c->degree_bits;
    batch_from_values(&k->cs, const_sigma_values, c->num_constants + c->num_routed, (int)c->degree_bits, (int)c->rate_bits, (int)c->cap_height, (int)c->hasher);
    k->sigma_values = (u64 *)malloc((size_t)c->num_routed * n * 8);
    memcpy(k->sigma_values, const_sigma_values + (size_t)c->num_constants * n, (size_t)c->num_routed * n * 8);
    memcpy(k->digest, digest, 32);
    if (cap_out) memcpy(cap_out, batch_cap(&k->cs), 32ULL << c->cap_height);
    return k;
}
void orc_circuit_set_digest(orc_circuit *k, const u64 digest[4]) { memcpy(k->digest, digest, 32); }
void orc_circuit_free(orc_circuit *k) { if (!k) return; batch_free(&k->cs); free(k->sigma_values); free(k); }

/* Merkle proof extraction: siblings bottom-up */
static void write_merkle_proof(wbuf *w, const u64 *digests, int log_L, int cap_height, size_t index) {
    wb_u8(w, (uint8_t)(log_L - cap_height));
    const u64 *lvl = digests;
    for (int lv = log_L; lv > cap_height; lv--) {
        wb_put(w, lvl + 4 * (index ^ 1), 32);
        lvl += 4ULL << lv; index >>= 1;
    }
}

/* debug taps (may be NULL): intermediate values for GPU-vs-oracle triage */
typedef struct {
    u64 betas[4], gammas[4], alphas[4]; u64 zeta[2]; u64 fri_alpha[2]; u64 fri_betas[16][2];
    u64 *zs_pp_values;      /* [num_zs_pp][n]  (caller-allocated or NULL) */
    u64 *quotient_coeffs;   /* [num_challenges*qdf][n] natural order */
    u64 *fri_values0;       /* [8n][2] layer-0 values, bit-reversed order */
} orc_debug;

int orc_prove(const orc_circuit *k, const u64 *wires_values, const u64 *pis, u64 pow_hint,
              uint8_t **proof_out, size_t *proof_len, orc_debug *dbg) {
    const orc_common *c = &k->c;
    const int nlog = (int)c->degree_bits, rb = (int)c->rate_bits, ch_h = (int)c->cap_height;
    const size_t n = 1ULL << nlog, N = n << rb;
    const int Llog = nlog + rb;
    const uint32_t W = c->num_wires, R = c->num_routed, NC = c->num_constants, nch = c->num_challenges, qdf = c->qdf, npp = c->npp;
    const size_t capn = 4ULL << ch_h;
    if (qdf != (1u << rb)) return -5;

    orc_t0_ = omp_get_wtime();
    u64 pi_hash[4]; orc_hash_no_pad(pis, c->num_pi, pi_hash);
    const int HS = (int)c->hasher; const size_t ncaph = 1ULL << ch_h;
    batch wires; batch_from_values(&wires, wires_values, W, nlog, rb, ch_h, HS);

    ORC_TICK("wires commit");
    chal ch; ch_init(&ch);
    ch_observe_hash(&ch, k->digest, HS); ch_observe_n(&ch, pi_hash, 4); ch_observe_cap(&ch, batch_cap(&wires), ncaph, HS);
    u64 betas[4], gammas[4], alphas[4];
    for (uint32_t i = 0; i < nch; i++) betas[i] = ch_get(&ch);
    for (uint32_t i = 0; i < nch; i++) gammas[i] = ch_get(&ch);

    /* ---- partial products and Z (prover.rs::wires_permutation_partial_products_and_zs) ---- */
    const uint32_t nzp = nch * (1 + npp);
    u64 *zs_pp = (u64 *)malloc((size_t)nzp * n * 8);      /* order: Z_0..Z_{nch-1}, pp_0[0..npp), pp_1[...] */
    u64 w_n = gl_root(nlog);
    u64 *subgroup = (u64 *)malloc(n * 8);
    subgroup[0] = 1; for (size_t i = 1; i < n; i++) subgroup[i] = gl_mul(subgroup[i - 1], w_n);
    for (uint32_t cidx = 0; cidx < nch; cidx++) {
        u64 beta = betas[cidx], gamma = gammas[cidx];
        u64 *chunkprod = (u64 *)malloc((size_t)(npp + 1) * n * 8);   /* [row][npp+1] */
        #pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; i++) {
            u64 x = subgroup[i];
            for (uint32_t kk = 0; kk <= npp; kk++) {
                u64 num = 1, den = 1;
                for (uint32_t j = kk * qdf; j < (kk + 1) * qdf && j < R; j++) {
                    u64 wv = wires_values[(size_t)j * n + i];
                    u64 s_id = gl_mul(c->k_is[j], x);
                    num = gl_mul(num, gl_add(gl_add(wv, gl_mul(beta, s_id)), gamma));
                    den = gl_mul(den, gl_add(gl_add(wv, gl_mul(beta, k->sigma_values[(size_t)j * n + i])), gamma));
                }
                chunkprod[i * (npp + 1) + kk] = gl_mul(num, gl_inv(den));
            }
        }
        u64 z = 1;
        for (size_t i = 0; i < n; i++) {
            zs_pp[(size_t)cidx * n + i] = z;
            u64 acc = z;
            for (uint32_t kk = 0; kk <= npp; kk++) {
                acc = gl_mul(acc, chunkprod[i * (npp + 1) + kk]);
                if (kk < npp) zs_pp[((size_t)nch + (size_t)cidx * npp + kk) * n + i] = acc;
            }
            z = acc;
        }
        free(chunkprod);
        if (z != 1) { fprintf(stderr, "orc_prove: permutation product != 1 (copy constraints violated)\n"); }
    }
    if (dbg && dbg->zs_pp_values) memcpy(dbg->zs_pp_values, zs_pp, (size_t)nzp * n * 8);
    ORC_TICK("partial products");
    batch zsb; batch_from_values(&zsb, zs_pp, nzp, nlog, rb, ch_h, HS);
    ORC_TICK("zs commit");
    ch_observe_cap(&ch, batch_cap(&zsb), ncaph, HS);
    for (uint32_t i = 0; i < nch; i++) alphas[i] = ch_get(&ch);

    /* ---- quotient (prover.rs::compute_quotient_polys) ---- */
    u64 *qvals = (u64 *)malloc((size_t)nch * N * 8);       /* natural order */
    {
        u64 w_N = gl_root(Llog);
        u64 *points = (u64 *)malloc(N * 8);
        points[0] = 1; for (size_t i = 1; i < N; i++) points[i] = gl_mul(points[i - 1], w_N);
        u64 g_pow_n = gl_pow(GL_GEN, n);
        u64 zh_inv[64]; u64 w_rate = gl_root(rb);
        { u64 x = 1; for (uint32_t i = 0; i < (1u << rb); i++) { zh_inv[i] = gl_inv(gl_sub(gl_mul(g_pow_n, x), 1)); x = gl_mul(x, w_rate); } }
        u64 zh_ev[64]; { u64 x = 1; for (uint32_t i = 0; i < (1u << rb); i++) { zh_ev[i] = gl_sub(gl_mul(g_pow_n, x), 1); x = gl_mul(x, w_rate); } }
        const uint32_t nterms = nch + nch * (npp + 1) + c->num_gate_constraints;
        int bad = 0;
        #pragma omp parallel
        {
            u64 *consts = (u64 *)malloc((NC + R) * 8), *wv = (u64 *)malloc(W * 8), *zp = (u64 *)malloc(nzp * 8), *znext = (u64 *)malloc(nzp * 8);
            u64 *terms = (u64 *)malloc(nterms * 8), *gtot = (u64 *)malloc((c->num_gate_constraints + 1) * 8), *scr = (u64 *)malloc(1024 * 8);
            #pragma omp for schedule(static)
            for (size_t i = 0; i < N; i++) {
                size_t pos = bitrev(i, Llog), pos_next = bitrev((i + (1ULL << rb)) % N, Llog);
                u64 x = gl_mul(GL_GEN, points[i]);
                for (uint32_t j = 0; j < NC + R; j++) consts[j] = k->cs.lde[(size_t)j * N + pos];
                for (uint32_t j = 0; j < W; j++) wv[j] = wires.lde[(size_t)j * N + pos];
                for (uint32_t j = 0; j < nzp; j++) { zp[j] = zsb.lde[(size_t)j * N + pos]; }
                for (uint32_t j = 0; j < nch; j++) znext[j] = zsb.lde[(size_t)j * N + pos_next];
                orc_eval_filtered(c->gates, c->num_gates, c->num_selectors, c->num_gate_constraints, consts, wv, pi_hash, gtot, scr);
                u64 l0 = gl_mul(zh_ev[i % (1u << rb)], gl_inv(gl_mul(n % GL_P, gl_sub(x, 1))));
                uint32_t t = 0;
                for (uint32_t ci = 0; ci < nch; ci++) terms[t++] = gl_mul(l0, gl_sub(zp[ci], 1));
                for (uint32_t ci = 0; ci < nch; ci++) {
                    u64 prev = zp[ci];
                    for (uint32_t kk = 0; kk <= npp; kk++) {
                        u64 num = 1, den = 1;
                        for (uint32_t j = kk * qdf; j < (kk + 1) * qdf && j < R; j++) {
                            num = gl_mul(num, gl_add(gl_add(wv[j], gl_mul(betas[ci], gl_mul(c->k_is[j], x))), gammas[ci]));
                            den = gl_mul(den, gl_add(gl_add(wv[j], gl_mul(betas[ci], consts[NC + j])), gammas[ci]));
                        }
                        u64 next = kk < npp ? zp[nch + ci * npp + kk] : znext[ci];
                        terms[t++] = gl_sub(gl_mul(prev, num), gl_mul(next, den));
                        prev = next;
                    }
                }
                for (uint32_t j = 0; j < c->num_gate_constraints; j++) terms[t++] = gtot[j];
                for (uint32_t ci = 0; ci < nch; ci++) {
                    u64 acc = 0;
                    for (int j = (int)nterms - 1; j >= 0; j--) acc = gl_add(gl_mul(acc, alphas[ci]), terms[j]);
                    qvals[(size_t)ci * N + i] = gl_mul(acc, zh_inv[i % (1u << rb)]);
                }
            }
            free(consts); free(wv); free(zp); free(znext); free(terms); free(gtot); free(scr);
        }
        (void)bad;
        free(points);
    }
    ORC_TICK("quotient eval");
    /* coset_ifft(shift): ifft then divide coefficient k by shift^k */
    batch quot; quot.coeffs = (u64 *)malloc((size_t)nch * qdf * n * 8);
    {
        orc_ntt_batch(qvals, nch, Llog, 1);
        u64 ginv = gl_inv(GL_GEN);
        for (uint32_t ci = 0; ci < nch; ci++) {
            u64 s = 1;
            for (size_t kx = 0; kx < N; kx++) { qvals[(size_t)ci * N + kx] = gl_mul(qvals[(size_t)ci * N + kx], s); s = gl_mul(s, ginv); }
            memcpy(quot.coeffs + (size_t)ci * qdf * n, qvals + (size_t)ci * N, N * 8);    /* chunks of n are contiguous */
        }
    }
    free(qvals);
    if (dbg && dbg->quotient_coeffs) memcpy(dbg->quotient_coeffs, quot.coeffs, (size_t)nch * qdf * n * 8);
    ORC_TICK("quotient ifft");
    batch_from_coeffs(&quot, (size_t)nch * qdf, nlog, rb, ch_h, HS);
    ORC_TICK("quotient commit");
    ch_observe_cap(&ch, batch_cap(&quot), ncaph, HS);
    e2 zeta = ch_get_ext(&ch);

    /* ---- openings ---- */
    const batch *oracles[4] = { &k->cs, &wires, &zsb, &quot };
    size_t total_polys = 0; for (int o = 0; o < 4; o++) total_polys += oracles[o]->cols;
    e2 *open_zeta = (e2 *)malloc(total_polys * sizeof(e2));
    e2 *open_next = (e2 *)malloc(nch * sizeof(e2));
    e2 gzeta = e2_scale(zeta, w_n);
    {
        /* one Horner evaluation per polynomial (upstream: PolynomialBatch::eval via rayon over the polynomials) */
        #pragma omp parallel for schedule(dynamic, 4)
        for (size_t idx = 0; idx < total_polys; idx++) {
            size_t q = idx; int o = 0; while (q >= oracles[o]->cols) { q -= oracles[o]->cols; o++; }
            const u64 *cf = oracles[o]->coeffs + q * n;
            e2 acc = { 0, 0 };
            for (size_t kx = n; kx-- > 0;) { acc = e2_mul(acc, zeta); acc.a = gl_add(acc.a, cf[kx]); }
            open_zeta[idx] = acc;
        }
        for (uint32_t ci = 0; ci < nch; ci++) {
            const u64 *cf = zsb.coeffs + (size_t)ci * n;
            e2 acc = { 0, 0 };
            for (size_t kx = n; kx-- > 0;) { acc = e2_mul(acc, gzeta); acc.a = gl_add(acc.a, cf[kx]); }
            open_next[ci] = acc;
        }
    }
    ORC_TICK("openings");
    /* observe openings: constants, sigmas, wires, zs, partial products, quotient (== oracle order), then zs_next */
    for (size_t i = 0; i < total_polys; i++) { ch_observe(&ch, open_zeta[i].a); ch_observe(&ch, open_zeta[i].b); }
    for (uint32_t i = 0; i < nch; i++) { ch_observe(&ch, open_next[i].a); ch_observe(&ch, open_next[i].b); }

    /* ---- FRI (oracle.rs::prove_openings) ---- */
    e2 fri_alpha = ch_get_ext(&ch);
    e2 *final_poly = (e2 *)calloc(n, sizeof(e2));
    {
        /* batch 0: all polys at zeta ; batch 1: zs at g*zeta.  final = sum_i alpha^{shift} * quotient_i */
        for (int bi = 0; bi < 2; bi++) {
            size_t cnt = bi == 0 ? total_polys : nch;
            e2 *comp = (e2 *)calloc(n, sizeof(e2));
            /* reduce_polys_base: Horner from the last polynomial */
            for (size_t p = cnt; p-- > 0;) {
                const u64 *cf;
                if (bi == 0) { size_t q = p; int o = 0; while (q >= oracles[o]->cols) { q -= oracles[o]->cols; o++; } cf = oracles[o]->coeffs + q * n; }
                else cf = zsb.coeffs + p * n;
                #pragma omp parallel for schedule(static)
                for (size_t kx = 0; kx < n; kx++) { e2 t = e2_mul(comp[kx], fri_alpha); t.a = gl_add(t.a, cf[kx]); comp[kx] = t; }
            }
            /* divide_by_linear(point): synthetic division, drop remainder */
            e2 point = bi == 0 ? zeta : gzeta;
            e2 *quo = (e2 *)calloc(n, sizeof(e2));
            e2 carry = { 0, 0 };
            for (size_t kx = n; kx-- > 0;) {
                e2 cur = e2_add(comp[kx], e2_mul(carry, point));   /* b_{k} = a_k + z*b_{k+1}; quotient coeff k-1 = b_k */
                if (kx > 0) quo[kx - 1] = cur;
                carry = cur;
            }
            /* alpha.shift_poly(final): final *= alpha^cnt ; final += quo */
            e2 sh = e2_pow(fri_alpha, cnt);
            for (size_t kx = 0; kx < n; kx++) final_poly[kx] = e2_add(e2_mul(final_poly[kx], sh), quo[kx]);
            free(comp); free(quo);
        }
    }
    ORC_TICK("fri batch poly");
    /* lde + coset fft (shift g) of the extension-field polynomial, component-wise */
    size_t cur_len = N;
    e2 *coeffs = (e2 *)calloc(N, sizeof(e2));
    memcpy(coeffs, final_poly, n * sizeof(e2));
    free(final_poly);
    e2 *values = (e2 *)malloc(N * sizeof(e2));
    u64 shift = GL_GEN;
    #define COSET_FFT(LEN_LOG) do { \
        size_t len_ = 1ULL << (LEN_LOG); u64 *ta = (u64 *)malloc(len_ * 8), *tb = (u64 *)malloc(len_ * 8); u64 s_ = 1; \
        for (size_t i_ = 0; i_ < len_; i_++) { ta[i_] = gl_mul(coeffs[i_].a, s_); tb[i_] = gl_mul(coeffs[i_].b, s_); s_ = gl_mul(s_, shift); } \
        orc_ntt(ta, (LEN_LOG), 0); orc_ntt(tb, (LEN_LOG), 0); \
        for (size_t i_ = 0; i_ < len_; i_++) { values[i_].a = ta[i_]; values[i_].b = tb[i_]; } free(ta); free(tb); } while (0)
    int cur_log = Llog;
    COSET_FFT(cur_log);

    wbuf out = { 0, 0, 0 };
    wb_put(&out, batch_cap(&wires), capn * 8); wb_put(&out, batch_cap(&zsb), capn * 8); wb_put(&out, batch_cap(&quot), capn * 8);
    {   /* openings: constants, sigmas, wires, zs, zs_next, partial products, quotient */
        size_t o_const = 0, o_sig = NC, o_w = NC + R, o_z = o_w + W, o_pp = o_z + nch, o_q = o_z + nzp;
        for (size_t i = o_const; i < o_sig; i++) wb_put(&out, &open_zeta[i], 16);
        for (size_t i = o_sig; i < o_w; i++) wb_put(&out, &open_zeta[i], 16);
        for (size_t i = o_w; i < o_z; i++) wb_put(&out, &open_zeta[i], 16);
        for (size_t i = o_z; i < o_pp; i++) wb_put(&out, &open_zeta[i], 16);
        for (uint32_t i = 0; i < nch; i++) wb_put(&out, &open_next[i], 16);
        for (size_t i = o_pp; i < o_q; i++) wb_put(&out, &open_zeta[i], 16);
        for (size_t i = o_q; i < total_polys; i++) wb_put(&out, &open_zeta[i], 16);
    }
    /* commit phase (prover.rs::fri_committed_trees) */
    u64 *fri_leaves[16]; u64 *fri_digests[16]; int fri_logleaves[16];
    for (uint32_t li = 0; li < c->n_arity; li++) {
        int ab = (int)c->arity_bits[li]; size_t arity = 1ULL << ab;
        /* reverse_index_bits_in_place(values) ; leaves = chunks of `arity` ext values flattened */
        u64 *leaves = (u64 *)malloc(cur_len * 16);
        #pragma omp parallel for schedule(static)
        for (size_t i = 0; i < cur_len; i++) { size_t r = bitrev(i, cur_log); leaves[2 * r] = values[i].a; leaves[2 * r + 1] = values[i].b; }
        if (li == 0 && dbg && dbg->fri_values0) memcpy(dbg->fri_values0, leaves, cur_len * 16);
        int logleaves = cur_log - ab;
        u64 *dg = (u64 *)malloc(orc_merkle_num_u64(logleaves, ch_h) * 8);
        orc_merkle_rowmajor_h(leaves, 2 * arity, logleaves, ch_h, dg, HS);
        fri_leaves[li] = leaves; fri_digests[li] = dg; fri_logleaves[li] = logleaves;
        const u64 *cap = dg + orc_merkle_num_u64(logleaves, ch_h) - capn;
        wb_put(&out, cap, capn * 8);
        ch_observe_cap(&ch, cap, ncaph, HS);
        e2 beta = ch_get_ext(&ch);
        if (dbg) { dbg->fri_betas[li][0] = beta.a; dbg->fri_betas[li][1] = beta.b; }
        /* coeffs = chunks(arity).map(reduce_with_powers(chunk, beta)) */
        size_t new_len = cur_len >> ab;
        e2 *folded = (e2 *)malloc(new_len * sizeof(e2));
        #pragma omp parallel for schedule(static)
        for (size_t i = 0; i < new_len; i++) {
            e2 acc = { 0, 0 };
            for (size_t j = arity; j-- > 0;) acc = e2_add(e2_mul(acc, beta), coeffs[i * arity + j]);
            folded[i] = acc;
        }
        memcpy(coeffs, folded, new_len * sizeof(e2)); free(folded);
        shift = gl_pow(shift, arity);
        cur_len = new_len; cur_log -= ab;
        COSET_FFT(cur_log);
    }
    ORC_TICK("fri commit phase");
    /* final poly: truncate to len >> rate_bits (the removed coefficients must be zero) */
    size_t fin_len = cur_len >> rb;
    for (size_t i = fin_len; i < cur_len; i++) if (coeffs[i].a || coeffs[i].b) { fprintf(stderr, "orc_prove: final poly has high-degree terms (quotient not low degree)\n"); break; }
    for (size_t i = 0; i < fin_len; i++) { ch_observe(&ch, coeffs[i].a); ch_observe(&ch, coeffs[i].b); }

    /* PoW: minimum nonce (== reference with RAYON_NUM_THREADS=1), or the caller's hint */
    u64 pow_witness = 0;
    {
        chal base = ch;                      /* pending inputs stay buffered; candidate goes to slot nin */
        int found = 0;
        if (pow_hint) {
            chal t = base; ch_observe(&t, pow_hint);
            if ((ch_get(&t) >> (64 - c->pow_bits)) == 0) { pow_witness = pow_hint; found = 1; }
        }
        /* minimum nonce: scan windows of 2^14 candidates in parallel, keep the smallest hit of the first window that has one */
        for (u64 start = 0; !found; start += 1u << 14) {
            u64 best = ~0ULL;
            #pragma omp parallel for schedule(static) reduction(min : best)
            for (u64 cand = start; cand < start + (1u << 14); cand++) {
                chal t = base; ch_observe(&t, cand);
                if ((ch_get(&t) >> (64 - c->pow_bits)) == 0 && cand < best) best = cand;
            }
            if (best != ~0ULL) { pow_witness = best; found = 1; }
        }
        ch_observe(&ch, pow_witness);
        u64 resp = ch_get(&ch);
        if (resp >> (64 - c->pow_bits)) return -3;
    }
    ORC_TICK("pow");
    /* query rounds */
    wbuf qb = { 0, 0, 0 };
    for (uint32_t q = 0; q < c->num_queries; q++) {
        size_t x_index = (size_t)(ch_get(&ch) % N);
        for (int o = 0; o < 4; o++) {
            const batch *b = oracles[o];
            for (size_t p = 0; p < b->cols; p++) wb_u64(&qb, b->lde[p * N + x_index]);
            write_merkle_proof(&qb, b->digests, Llog, ch_h, x_index);
        }
        for (uint32_t li = 0; li < c->n_arity; li++) {
            int ab = (int)c->arity_bits[li]; size_t arity = 1ULL << ab;
            size_t leaf = x_index >> ab;
            wb_put(&qb, fri_leaves[li] + leaf * 2 * arity, 16 * arity);
            write_merkle_proof(&qb, fri_digests[li], fri_logleaves[li], ch_h, leaf);
            x_index = leaf;
        }
    }
    wb_put(&out, qb.p, qb.len); free(qb.p);
    for (size_t i = 0; i < fin_len; i++) wb_put(&out, &coeffs[i], 16);
    wb_u64(&out, pow_witness);
    wb_u64(&out, c->num_pi);
    for (uint32_t i = 0; i < c->num_pi; i++) wb_u64(&out, pis[i]);

    ORC_TICK("queries+serialize");
    if (dbg) {
        memcpy(dbg->betas, betas, sizeof betas); memcpy(dbg->gammas, gammas, sizeof gammas); memcpy(dbg->alphas, alphas, sizeof alphas);
        dbg->zeta[0] = zeta.a; dbg->zeta[1] = zeta.b; dbg->fri_alpha[0] = fri_alpha.a; dbg->fri_alpha[1] = fri_alpha.b;
    }
    for (uint32_t li = 0; li < c->n_arity; li++) { free(fri_leaves[li]); free(fri_digests[li]); }
    free(coeffs); free(values); free(open_zeta); free(open_next); free(zs_pp); free(subgroup);
    batch_free(&wires); batch_free(&zsb); batch_free(&quot);
    *proof_out = out.p; *proof_len = out.len;
    return 0;
}
void orc_free(void *p) { free(p); }
