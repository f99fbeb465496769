/* Base-field gate constraint evaluators of the C oracle prover.  TEST INFRASTRUCTURE.
 * Standard gates follow gnark-plonky2-verifier/plonk/gates/ (each *_gate.go) (EvalUnfiltered, evaluated over F_p instead of
 * F_p^2); custom gates follow crypto/plonky2_u32/src/gates/ (each .rs) (eval_unfiltered).  Constraint order is part of
 * the spec.  Cross-checked against oracle/gates.py (pinned by the golden proofs) in tests/test_gate_parity.py. */
#include "gates.h"
#include "poseidon_constants.h"
#include <string.h>

typedef struct { u64 a, b; } alg;   /* extension algebra element over F_p: a + b*X, X^2 = 7 */
static inline alg alg_add(alg x, alg y) { alg r = { gl_add(x.a, y.a), gl_add(x.b, y.b) }; return r; }
static inline alg alg_sub(alg x, alg y) { alg r = { gl_sub(x.a, y.a), gl_sub(x.b, y.b) }; return r; }
static inline alg alg_mul(alg x, alg y) {
    alg r = { gl_add(gl_mul(x.a, y.a), gl_mul(7, gl_mul(x.b, y.b))), gl_add(gl_mul(x.a, y.b), gl_mul(x.b, y.a)) }; return r; }
static inline alg alg_scalar(u64 s, alg x) { alg r = { gl_mul(s, x.a), gl_mul(s, x.b) }; return r; }
static inline alg W2(const u64 *w, uint32_t i) { alg r = { w[i], w[i + 1] }; return r; }

static inline u64 range_prod(u64 x, int n) { u64 p = 1; for (int k = 0; k < n; k++) p = gl_mul(p, gl_sub(x, (u64)k)); return p; }
static inline u64 sbox7(u64 x) { u64 x2 = gl_mul(x, x), x4 = gl_mul(x2, x2); return gl_mul(x4, gl_mul(x2, x)); }
static void mds(u64 s[12]) {
    u64 o[12];
    for (int r = 0; r < 12; r++) {
        u64 acc = 0;
        for (int i = 0; i < 12; i++) acc = gl_add(acc, gl_mul(s[(i + r) % 12], PG_MDS_CIRC[i]));
        acc = gl_add(acc, gl_mul(s[r], PG_MDS_DIAG[r]));
        o[r] = acc;
    }
    memcpy(s, o, sizeof o);
}

static int eval_poseidon(const u64 *w, u64 *out) {
    int k = 0;
    u64 swap = w[24];
    out[k++] = gl_mul(swap, gl_sub(swap, 1));
    for (int i = 0; i < 4; i++) out[k++] = gl_sub(gl_mul(swap, gl_sub(w[i + 4], w[i])), w[25 + i]);
    u64 st[12];
    for (int i = 0; i < 4; i++) { st[i] = gl_add(w[i], w[25 + i]); st[i + 4] = gl_sub(w[i + 4], w[25 + i]); }
    for (int i = 8; i < 12; i++) st[i] = w[i];
    int rc = 0;
    for (int r = 0; r < 4; r++, rc++) {
        for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], PG_RC[12 * rc + i] % GL_P);
        if (r != 0) for (int i = 0; i < 12; i++) { u64 sin = w[29 + 12 * (r - 1) + i]; out[k++] = gl_sub(st[i], sin); st[i] = sin; }
        for (int i = 0; i < 12; i++) st[i] = sbox7(st[i]);
        mds(st);
    }
    for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], PG_FP_FIRST_RC[i] % GL_P);
    { u64 o[12]; o[0] = st[0];
      for (int d = 1; d < 12; d++) { u64 acc = 0; for (int r = 1; r < 12; r++) acc = gl_add(acc, gl_mul(st[r], PG_FP_INIT[11 * (r - 1) + d - 1] % GL_P)); o[d] = acc; }
      memcpy(st, o, sizeof o); }
    for (int r = 0; r < 22; r++) {
        u64 sin = w[65 + r];
        out[k++] = gl_sub(st[0], sin);
        st[0] = sbox7(sin);
        if (r != 21) st[0] = gl_add(st[0], PG_FP_RC[r] % GL_P);
        u64 d = gl_mul(st[0], 25);
        for (int i = 1; i < 12; i++) d = gl_add(d, gl_mul(st[i], PG_FP_W_HATS[11 * r + i - 1] % GL_P));
        for (int i = 1; i < 12; i++) st[i] = gl_add(st[i], gl_mul(st[0], PG_FP_VS[11 * r + i - 1] % GL_P));
        st[0] = d;
    }
    rc += 22;
    for (int r = 0; r < 4; r++, rc++) {
        for (int i = 0; i < 12; i++) st[i] = gl_add(st[i], PG_RC[12 * rc + i] % GL_P);
        for (int i = 0; i < 12; i++) { u64 sin = w[87 + 12 * r + i]; out[k++] = gl_sub(st[i], sin); st[i] = sin; }
        for (int i = 0; i < 12; i++) st[i] = sbox7(st[i]);
        mds(st);
    }
    for (int i = 0; i < 12; i++) out[k++] = gl_sub(st[i], w[12 + i]);
    return k;
}

static void partial_interp(const u64 *dom, const alg *vals, const u64 *ws, int n, alg point, alg *ev, alg *prod) {
    for (int i = 0; i < n; i++) {
        alg x = { dom[i], 0 };
        alg term = alg_sub(point, x);
        alg wv = alg_scalar(ws[i], vals[i]);
        *ev = alg_add(alg_mul(*ev, term), alg_mul(wv, *prod));
        *prod = alg_mul(*prod, term);
    }
}

int orc_eval_gate(const orc_gate *g, const u64 *c, const u64 *w, const u64 pi[4], u64 *out) {
    int k = 0;
    switch (g->kind) {
    case G_NOOP: return 0;
    case G_CONSTANT: for (uint32_t i = 0; i < g->p[0]; i++) out[k++] = gl_sub(c[i], w[i]); return k;
    case G_PUBLIC_INPUT: for (int i = 0; i < 4; i++) out[k++] = gl_sub(w[i], pi[i]); return k;
    case G_ARITHMETIC:
        for (uint32_t i = 0; i < g->p[0]; i++)
            out[k++] = gl_sub(w[4 * i + 3], gl_add(gl_mul(gl_mul(w[4 * i], w[4 * i + 1]), c[0]), gl_mul(w[4 * i + 2], c[1])));
        return k;
    case G_ARITHMETIC_EXT:
        for (uint32_t i = 0; i < g->p[0]; i++) {
            alg comp = alg_add(alg_scalar(c[1], W2(w, 8 * i + 4)), alg_scalar(c[0], alg_mul(W2(w, 8 * i), W2(w, 8 * i + 2))));
            alg d = alg_sub(W2(w, 8 * i + 6), comp); out[k++] = d.a; out[k++] = d.b;
        }
        return k;
    case G_MUL_EXT:
        for (uint32_t i = 0; i < g->p[0]; i++) {
            alg d = alg_sub(W2(w, 6 * i + 4), alg_scalar(c[0], alg_mul(W2(w, 6 * i), W2(w, 6 * i + 2)))); out[k++] = d.a; out[k++] = d.b;
        }
        return k;
    case G_BASE_SUM: {
        uint32_t n = g->p[0], base = g->p[1];
        u64 acc = 0;
        for (int i = (int)n - 1; i >= 0; i--) acc = gl_add(gl_mul(acc, base), w[1 + i]);
        out[k++] = gl_sub(acc, w[0]);
        for (uint32_t i = 0; i < n; i++) out[k++] = range_prod(w[1 + i], (int)base);
        return k; }
    case G_REDUCING: case G_REDUCING_EXT: {
        uint32_t n = g->p[0]; int ext = g->kind == G_REDUCING_EXT;
        uint32_t start_accs = 6 + (ext ? 2 * n : n);
        alg alpha = W2(w, 2), acc = W2(w, 4);
        for (uint32_t i = 0; i < n; i++) {
            alg coeff; if (ext) coeff = W2(w, 6 + 2 * i); else { coeff.a = w[6 + i]; coeff.b = 0; }
            alg ai = (i == n - 1) ? W2(w, 0) : W2(w, start_accs + 2 * i);
            alg t = alg_sub(alg_add(alg_mul(acc, alpha), coeff), ai);
            out[k++] = t.a; out[k++] = t.b; acc = ai;
        }
        return k; }
    case G_RANDOM_ACCESS: {
        uint32_t bits = g->p[0], copies = g->p[1], extra = g->p[2], vec = 1u << bits;
        uint32_t routed = (2 + vec) * copies + extra;
        for (uint32_t cp = 0; cp < copies; cp++) {
            uint32_t base = (2 + vec) * cp;
            u64 items[64];
            for (uint32_t i = 0; i < vec; i++) items[i] = w[base + 2 + i];
            const u64 *b = w + routed + cp * bits;
            for (uint32_t i = 0; i < bits; i++) out[k++] = gl_sub(gl_mul(b[i], b[i]), b[i]);
            u64 rec = 0;
            for (int i = (int)bits - 1; i >= 0; i--) rec = gl_add(gl_mul(rec, 2), b[i]);
            out[k++] = gl_sub(rec, w[base]);
            uint32_t len = vec;
            for (uint32_t bi = 0; bi < bits; bi++) {
                for (uint32_t i = 0; i < len; i += 2) items[i / 2] = gl_add(items[i], gl_mul(b[bi], gl_sub(items[i + 1], items[i])));
                len /= 2;
            }
            out[k++] = gl_sub(items[0], w[base + 1]);
        }
        for (uint32_t i = 0; i < extra; i++) out[k++] = gl_sub(c[i], w[(2 + vec) * copies + i]);
        return k; }
    case G_POSEIDON_MDS:
        for (int r = 0; r < 12; r++) {
            alg acc = { 0, 0 };
            for (int i = 0; i < 12; i++) acc = alg_add(acc, alg_scalar(PG_MDS_CIRC[i], W2(w, 2 * ((i + r) % 12))));
            acc = alg_add(acc, alg_scalar(PG_MDS_DIAG[r], W2(w, 2 * r)));
            alg d = alg_sub(W2(w, 24 + 2 * r), acc); out[k++] = d.a; out[k++] = d.b;
        }
        return k;
    case G_POSEIDON: return eval_poseidon(w, out);
    case G_COSET_INTERPOLATION: {
        uint32_t sb = g->p[0], degree = g->p[1], np_ = 1u << sb, n_inter = (np_ - 2) / (degree - 1);
        uint32_t start_inter = 1 + 2 * np_ + 4;
        u64 shift = w[0];
        alg ep = W2(w, 1 + 2 * np_), sep = W2(w, start_inter + 4 * n_inter);
        alg t = alg_add(alg_scalar(gl_neg(shift), sep), ep); out[k++] = t.a; out[k++] = t.b;
        u64 dom[64]; alg vals[64];
        u64 gen = gl_root((int)sb); dom[0] = 1;
        for (uint32_t i = 1; i < np_; i++) dom[i] = gl_mul(dom[i - 1], gen);
        for (uint32_t i = 0; i < np_; i++) vals[i] = W2(w, 1 + 2 * i);
        alg ev = { 0, 0 }, prod = { 1, 0 };
        partial_interp(dom, vals, g->extra, (int)degree, sep, &ev, &prod);
        for (uint32_t i = 0; i < n_inter; i++) {
            alg ie = W2(w, start_inter + 2 * i), ip = W2(w, start_inter + 2 * (n_inter + i));
            alg d = alg_sub(ie, ev); out[k++] = d.a; out[k++] = d.b;
            d = alg_sub(ip, prod); out[k++] = d.a; out[k++] = d.b;
            uint32_t s = 1 + (degree - 1) * (i + 1), e = s + degree - 1; if (e > np_) e = np_;
            ev = ie; prod = ip;
            partial_interp(dom + s, vals + s, g->extra + s, (int)(e - s), sep, &ev, &prod);
        }
        alg d = alg_sub(W2(w, 1 + 2 * np_ + 2), ev); out[k++] = d.a; out[k++] = d.b;
        return k; }
    case G_U32_ARITHMETIC: {
        uint32_t n = g->p[0];
        for (uint32_t i = 0; i < n; i++) {
            const u64 *o = w + 6 * i;
            u64 comp = gl_add(gl_mul(o[0], o[1]), o[2]);
            out[k++] = gl_mul(gl_sub(gl_mul(o[5], gl_sub(0xFFFFFFFFULL, o[4])), 1), o[3]);
            out[k++] = gl_sub(gl_add(gl_mul(o[4], 1ULL << 32), o[3]), comp);
            u64 cl = 0, ch = 0;
            for (int j = 31; j >= 0; j--) {
                u64 limb = w[6 * n + 32 * i + j];
                out[k++] = range_prod(limb, 4);
                if (j < 16) cl = gl_add(gl_mul(cl, 4), limb); else ch = gl_add(gl_mul(ch, 4), limb);
            }
            out[k++] = gl_sub(cl, o[3]); out[k++] = gl_sub(ch, o[4]);
        }
        return k; }
    case G_U32_ADD_MANY: {
        uint32_t na = g->p[0], n = g->p[1], r = na + 3;
        for (uint32_t i = 0; i < n; i++) {
            u64 comp = 0;
            for (uint32_t j = 0; j < na; j++) comp = gl_add(comp, w[r * i + j]);
            comp = gl_add(comp, w[r * i + na]);
            u64 res = w[r * i + na + 1], carry = w[r * i + na + 2];
            out[k++] = gl_sub(gl_add(gl_mul(carry, 1ULL << 32), res), comp);
            u64 cr = 0, cc = 0;
            for (int j = 17; j >= 0; j--) {
                u64 limb = w[r * n + 18 * i + j];
                out[k++] = range_prod(limb, 4);
                if (j < 16) cr = gl_add(gl_mul(cr, 4), limb); else cc = gl_add(gl_mul(cc, 4), limb);
            }
            out[k++] = gl_sub(cr, res); out[k++] = gl_sub(cc, carry);
        }
        return k; }
    case G_U32_SUBTRACTION: {
        uint32_t n = g->p[0];
        for (uint32_t i = 0; i < n; i++) {
            const u64 *o = w + 5 * i;
            u64 init = gl_sub(gl_sub(o[0], o[1]), o[2]);
            out[k++] = gl_sub(o[3], gl_add(init, gl_mul(1ULL << 32, o[4])));
            u64 cl = 0;
            for (int j = 15; j >= 0; j--) { u64 limb = w[5 * n + 16 * i + j]; out[k++] = range_prod(limb, 4); cl = gl_add(gl_mul(cl, 4), limb); }
            out[k++] = gl_sub(cl, o[3]);
            out[k++] = gl_mul(o[4], gl_sub(1, o[4]));
        }
        return k; }
    case G_U32_RANGE_CHECK: {
        uint32_t n = g->p[0];
        for (uint32_t i = 0; i < n; i++) {
            u64 acc = 0;
            for (int j = 15; j >= 0; j--) acc = gl_add(gl_mul(acc, 4), w[n + 16 * i + j]);
            out[k++] = gl_sub(acc, w[i]);
            for (int j = 0; j < 16; j++) out[k++] = range_prod(w[n + 16 * i + j], 4);
        }
        return k; }
    case G_U32_INTERLEAVE: {
        uint32_t n = g->p[0];
        for (uint32_t i = 0; i < n; i++) {
            const u64 *b = w + 2 * n + 32 * i;
            u64 cx = 0, ci = 0;
            for (int j = 0; j < 32; j++) { cx = gl_add(gl_mul(cx, 2), b[j]); ci = gl_add(gl_mul(ci, 4), b[j]); }
            out[k++] = gl_sub(cx, w[2 * i]); out[k++] = gl_sub(ci, w[2 * i + 1]);
            for (int j = 0; j < 32; j++) out[k++] = range_prod(b[j], 2);
        }
        return k; }
    case G_UNINTERLEAVE_U32: case G_UNINTERLEAVE_B32: {
        uint32_t n = g->p[0]; int b32 = g->kind == G_UNINTERLEAVE_B32;
        for (uint32_t i = 0; i < n; i++) {
            const u64 *b = w + 3 * n + 64 * i;
            u64 ci = 0;
            for (int j = 0; j < 64; j++) ci = gl_add(gl_mul(ci, 2), b[j]);
            out[k++] = gl_sub(ci, w[3 * i]);
            u64 ce = 0, co = 0;
            for (int j = 0; j < 32; j++) {
                u64 coeff = b32 ? (1ULL << (2 * (31 - j))) : (1ULL << (31 - j));
                ce = gl_add(ce, gl_mul(coeff, b[2 * j])); co = gl_add(co, gl_mul(coeff, b[2 * j + 1]));
            }
            out[k++] = gl_sub(ce, w[3 * i + 1]); out[k++] = gl_sub(co, w[3 * i + 2]);
            for (int j = 0; j < 64; j++) out[k++] = range_prod(b[j], 2);
        }
        return k; }
    case G_COMPARISON: {
        uint32_t nb = g->p[0], nc = g->p[1], cb = (nb + nc - 1) / nc;
        u64 fc = 0, sc = 0;
        for (int i = (int)nc - 1; i >= 0; i--) { fc = gl_add(gl_mul(fc, 1ULL << cb), w[4 + i]); sc = gl_add(gl_mul(sc, 1ULL << cb), w[4 + nc + i]); }
        out[k++] = gl_sub(fc, w[0]); out[k++] = gl_sub(sc, w[1]);
        u64 msd = 0;
        for (uint32_t i = 0; i < nc; i++) {
            u64 f = w[4 + i], s = w[4 + nc + i];
            out[k++] = range_prod(f, 1 << cb); out[k++] = range_prod(s, 1 << cb);
            u64 diff = gl_sub(s, f), dummy = w[4 + 2 * nc + i], eq = w[4 + 3 * nc + i];
            out[k++] = gl_sub(gl_mul(diff, dummy), gl_sub(1, eq));
            out[k++] = gl_mul(eq, diff);
            u64 inter = w[4 + 4 * nc + i];
            out[k++] = gl_sub(inter, gl_mul(eq, msd));
            msd = gl_add(inter, gl_mul(gl_sub(1, eq), diff));
        }
        out[k++] = gl_sub(w[3], msd);
        const u64 *bits = w + 4 + 5 * nc;
        for (uint32_t i = 0; i <= cb; i++) out[k++] = gl_mul(bits[i], gl_sub(1, bits[i]));
        u64 comb = 0;
        for (int i = (int)cb; i >= 0; i--) comb = gl_add(gl_mul(comb, 2), bits[i]);
        out[k++] = gl_sub(gl_add(1ULL << cb, w[3]), comb);
        out[k++] = gl_sub(w[2], bits[cb]);
        return k; }
    }
    return -1;
}

void orc_eval_filtered(const orc_gate *gates, uint32_t num_gates, uint32_t num_selectors, uint32_t ngc,
                       const u64 *consts, const u64 *wires, const u64 pi[4], u64 *total, u64 *scratch) {
    for (uint32_t i = 0; i < ngc; i++) total[i] = 0;
    for (uint32_t gi = 0; gi < num_gates; gi++) {
        const orc_gate *g = &gates[gi];
        u64 sel = consts[g->selector_index], filt = 1;
        for (uint32_t i = g->group_start; i < g->group_end; i++) if (i != g->row) filt = gl_mul(filt, gl_sub(i, sel));
        if (num_selectors > 1) filt = gl_mul(filt, gl_sub(0xFFFFFFFFULL, sel));
        int k = orc_eval_gate(g, consts + num_selectors, wires, pi, scratch);
        for (int i = 0; i < k; i++) total[i] = gl_add(total[i], gl_mul(scratch[i], filt));
    }
}
