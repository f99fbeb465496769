// Gate constraint evaluators generic over the field type T (operator-overloaded), host side.  Used by b200_verify at
// T = F_p^2 (evaluation at zeta).  Formulas and constraint order: gnark-plonky2-verifier/plonk/gates/ (13 standard
// gates, each *_gate.go EvalUnfiltered) and crypto/plonky2_u32/src/gates/ (8 custom gates, eval_unfiltered) -- the same
// specification gates.cuh implements for the prover at T = F_p on the device.
#pragma once
#include "gates.cuh"
#include <vector>

template <class T> struct Alg { T a, b; };      // extension-algebra element a + b*X, X^2 = 7, over T
template <class T> static inline Alg<T> alg_add(const Alg<T> &x, const Alg<T> &y) { return {x.a + y.a, x.b + y.b}; }
template <class T> static inline Alg<T> alg_sub(const Alg<T> &x, const Alg<T> &y) { return {x.a - y.a, x.b - y.b}; }
template <class T> static inline Alg<T> alg_mul(const Alg<T> &x, const Alg<T> &y) { return {x.a * y.a + (x.b * y.b) * T::from(7), x.a * y.b + x.b * y.a}; }
template <class T> static inline Alg<T> alg_scalar(const T &s, const Alg<T> &x) { return {s * x.a, s * x.b}; }

template <class T> static T range_prod(const T &x, int n) { T p = T::from(1); for (int k = 0; k < n; k++) p = p * (x - T::from((u64)k)); return p; }
template <class T> static T sbox7(const T &x) { T x2 = x * x, x4 = x2 * x2; return x4 * (x2 * x); }
template <class T> static void mds12(T s[12]) {
    static const u64 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    T o[12];
    for (int r = 0; r < 12; r++) { T acc = T::from(0); for (int i = 0; i < 12; i++) acc = acc + s[(i + r) % 12] * T::from(C[i]); if (r == 0) acc = acc + s[0] * T::from(8); o[r] = acc; }
    for (int i = 0; i < 12; i++) s[i] = o[i];
}

// returns the constraints of gate g on (constants after the selector prefix, wires)
template <class T>
static bool eval_gate_generic(const GateDesc &g, const T *c, const T *w, const u64 pi_hash[4], const u64 *extra, std::vector<T> &out) {
    auto W2 = [&](u32 i) { return Alg<T>{w[i], w[i + 1]}; };
    auto push2 = [&](const Alg<T> &d) { out.push_back(d.a); out.push_back(d.b); };
    const T ONE = T::from(1), ZERO = T::from(0);
    switch (g.kind) {
    case G_NOOP: return true;
    case G_CONSTANT: for (u32 i = 0; i < g.p[0]; i++) out.push_back(c[i] - w[i]); return true;
    case G_PUBLIC_INPUT: for (int i = 0; i < 4; i++) out.push_back(w[i] - T::from(pi_hash[i])); return true;
    case G_ARITHMETIC: for (u32 i = 0; i < g.p[0]; i++) out.push_back(w[4 * i + 3] - (w[4 * i] * w[4 * i + 1] * c[0] + w[4 * i + 2] * c[1])); return true;
    case G_ARITHMETIC_EXT:
        for (u32 i = 0; i < g.p[0]; i++) push2(alg_sub(W2(8 * i + 6), alg_add(alg_scalar(c[1], W2(8 * i + 4)), alg_scalar(c[0], alg_mul(W2(8 * i), W2(8 * i + 2))))));
        return true;
    case G_MUL_EXT: for (u32 i = 0; i < g.p[0]; i++) push2(alg_sub(W2(6 * i + 4), alg_scalar(c[0], alg_mul(W2(6 * i), W2(6 * i + 2))))); return true;
    case G_BASE_SUM: {
        u32 n = g.p[0], base = g.p[1]; T acc = ZERO;
        for (int i = (int)n - 1; i >= 0; i--) acc = acc * T::from(base) + w[1 + i];
        out.push_back(acc - w[0]);
        for (u32 i = 0; i < n; i++) out.push_back(range_prod(w[1 + i], (int)base));
        return true; }
    case G_REDUCING: case G_REDUCING_EXT: {
        u32 n = g.p[0]; bool ext = g.kind == G_REDUCING_EXT; u32 start_accs = 6 + (ext ? 2 * n : n);
        Alg<T> alpha = W2(2), acc = W2(4);
        for (u32 i = 0; i < n; i++) {
            Alg<T> coeff = ext ? W2(6 + 2 * i) : Alg<T>{w[6 + i], ZERO};
            Alg<T> ai = (i == n - 1) ? W2(0) : W2(start_accs + 2 * i);
            push2(alg_sub(alg_add(alg_mul(acc, alpha), coeff), ai)); acc = ai;
        }
        return true; }
    case G_RANDOM_ACCESS: {
        u32 bits = g.p[0], copies = g.p[1], extra_c = g.p[2], vec = 1u << bits, routed = (2 + vec) * copies + extra_c;
        for (u32 cp = 0; cp < copies; cp++) {
            u32 base = (2 + vec) * cp;
            std::vector<T> items(w + base + 2, w + base + 2 + vec);
            const T *b = w + routed + cp * bits;
            for (u32 i = 0; i < bits; i++) out.push_back(b[i] * b[i] - b[i]);
            T rec = ZERO; for (int i = (int)bits - 1; i >= 0; i--) rec = rec * T::from(2) + b[i];
            out.push_back(rec - w[base]);
            u32 len = vec;
            for (u32 bi = 0; bi < bits; bi++) { for (u32 i = 0; i < len; i += 2) items[i / 2] = items[i] + b[bi] * (items[i + 1] - items[i]); len >>= 1; }
            out.push_back(items[0] - w[base + 1]);
        }
        for (u32 i = 0; i < extra_c; i++) out.push_back(c[i] - w[(2 + vec) * copies + i]);
        return true; }
    case G_POSEIDON_MDS: {
        static const u64 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
        for (int r = 0; r < 12; r++) {
            Alg<T> acc{ZERO, ZERO};
            for (int i = 0; i < 12; i++) acc = alg_add(acc, alg_scalar(T::from(C[i]), W2(2 * ((i + r) % 12))));
            if (r == 0) acc = alg_add(acc, alg_scalar(T::from(8), W2(0)));
            push2(alg_sub(W2(24 + 2 * r), acc));
        }
        return true; }
    case G_POSEIDON: {
        T swap = w[24];
        out.push_back(swap * (swap - ONE));
        T st[12];
        for (int i = 0; i < 4; i++) { out.push_back(swap * (w[i + 4] - w[i]) - w[25 + i]); st[i] = w[i] + w[25 + i]; st[i + 4] = w[i + 4] - w[25 + i]; }
        for (int i = 8; i < 12; i++) st[i] = w[i];
        for (int r = 0; r < 30; r++) {
            bool full = r < 4 || r >= 26;
            for (int i = 0; i < 12; i++) st[i] = st[i] + T::from(h_POSEIDON_RC[12 * r + i]);
            if (full) {
                int wbase = r < 4 ? 29 + 12 * (r - 1) : 87 + 12 * (r - 26);
                for (int i = 0; i < 12; i++) { if (r != 0) { out.push_back(st[i] - w[wbase + i]); st[i] = w[wbase + i]; } st[i] = sbox7(st[i]); }
            } else { out.push_back(st[0] - w[65 + (r - 4)]); st[0] = sbox7(w[65 + (r - 4)]); }
            mds12(st);
        }
        for (int i = 0; i < 12; i++) out.push_back(st[i] - w[12 + i]);
        return true; }
    case G_COSET_INTERPOLATION: {
        u32 sb = g.p[0], degree = g.p[1], np = 1u << sb, n_inter = (np - 2) / (degree - 1), start_inter = 1 + 2 * np + 4;
        const u64 *weights = extra + g.extra_off;
        Alg<T> ep = W2(1 + 2 * np), sep = W2(start_inter + 4 * n_inter);
        push2(alg_add(alg_scalar(ZERO - w[0], sep), ep));
        u64 gen = gl_root_of_unity((int)sb), x = 1;
        Alg<T> ev{ZERO, ZERO}, prod{ONE, ZERO};
        u32 next_break = degree, k = 0;
        for (u32 i = 0; i < np; i++) {
            if (i == next_break && k < n_inter) {
                Alg<T> ie = W2(start_inter + 2 * k), ip = W2(start_inter + 2 * (n_inter + k));
                push2(alg_sub(ie, ev)); push2(alg_sub(ip, prod)); ev = ie; prod = ip; k++; next_break += degree - 1;
            }
            Alg<T> term = alg_sub(sep, Alg<T>{T::from(x), ZERO});
            Alg<T> wv = alg_scalar(T::from(weights[i]), W2(1 + 2 * i));
            ev = alg_add(alg_mul(ev, term), alg_mul(wv, prod)); prod = alg_mul(prod, term);
            x = gl_mul(x, gen);
        }
        push2(alg_sub(W2(1 + 2 * np + 2), ev));
        return true; }
    case G_U32_ARITHMETIC: {
        u32 n = g.p[0];
        for (u32 i = 0; i < n; i++) {
            const T *o = w + 6 * i;
            out.push_back((o[5] * (T::from(0xFFFFFFFFULL) - o[4]) - ONE) * o[3]);
            out.push_back(o[4] * T::from(1ULL << 32) + o[3] - (o[0] * o[1] + o[2]));
            T cl = ZERO, ch = ZERO;
            for (int j = 31; j >= 0; j--) { const T &limb = w[6 * n + 32 * i + j]; out.push_back(range_prod(limb, 4)); if (j < 16) cl = cl * T::from(4) + limb; else ch = ch * T::from(4) + limb; }
            out.push_back(cl - o[3]); out.push_back(ch - o[4]);
        }
        return true; }
    case G_U32_ADD_MANY: {
        u32 na = g.p[0], n = g.p[1], r = na + 3;
        for (u32 i = 0; i < n; i++) {
            T comp = ZERO; for (u32 j = 0; j <= na; j++) comp = comp + w[r * i + j];
            const T &res = w[r * i + na + 1], &carry = w[r * i + na + 2];
            out.push_back(carry * T::from(1ULL << 32) + res - comp);
            T cr = ZERO, cc = ZERO;
            for (int j = 17; j >= 0; j--) { const T &limb = w[r * n + 18 * i + j]; out.push_back(range_prod(limb, 4)); if (j < 16) cr = cr * T::from(4) + limb; else cc = cc * T::from(4) + limb; }
            out.push_back(cr - res); out.push_back(cc - carry);
        }
        return true; }
    case G_U32_SUBTRACTION: {
        u32 n = g.p[0];
        for (u32 i = 0; i < n; i++) {
            const T *o = w + 5 * i;
            out.push_back(o[3] - (o[0] - o[1] - o[2] + T::from(1ULL << 32) * o[4]));
            T cl = ZERO;
            for (int j = 15; j >= 0; j--) { const T &limb = w[5 * n + 16 * i + j]; out.push_back(range_prod(limb, 4)); cl = cl * T::from(4) + limb; }
            out.push_back(cl - o[3]); out.push_back(o[4] * (ONE - o[4]));
        }
        return true; }
    case G_U32_RANGE_CHECK: {
        u32 n = g.p[0];
        for (u32 i = 0; i < n; i++) {
            T acc = ZERO; for (int j = 15; j >= 0; j--) acc = acc * T::from(4) + w[n + 16 * i + j];
            out.push_back(acc - w[i]);
            for (int j = 0; j < 16; j++) out.push_back(range_prod(w[n + 16 * i + j], 4));
        }
        return true; }
    case G_U32_INTERLEAVE: {
        u32 n = g.p[0];
        for (u32 i = 0; i < n; i++) {
            const T *b = w + 2 * n + 32 * i; T cx = ZERO, ci = ZERO;
            for (int j = 0; j < 32; j++) { cx = cx * T::from(2) + b[j]; ci = ci * T::from(4) + b[j]; }
            out.push_back(cx - w[2 * i]); out.push_back(ci - w[2 * i + 1]);
            for (int j = 0; j < 32; j++) out.push_back(range_prod(b[j], 2));
        }
        return true; }
    case G_UNINTERLEAVE_U32: case G_UNINTERLEAVE_B32: {
        u32 n = g.p[0]; bool b32 = g.kind == G_UNINTERLEAVE_B32;
        for (u32 i = 0; i < n; i++) {
            const T *b = w + 3 * n + 64 * i; T ci = ZERO, ce = ZERO, co = ZERO;
            for (int j = 0; j < 64; j++) ci = ci * T::from(2) + b[j];
            for (int j = 0; j < 32; j++) { T coeff = T::from(b32 ? (1ULL << (2 * (31 - j))) : (1ULL << (31 - j))); ce = ce + coeff * b[2 * j]; co = co + coeff * b[2 * j + 1]; }
            out.push_back(ci - w[3 * i]); out.push_back(ce - w[3 * i + 1]); out.push_back(co - w[3 * i + 2]);
            for (int j = 0; j < 64; j++) out.push_back(range_prod(b[j], 2));
        }
        return true; }
    case G_COMPARISON: {
        u32 nb = g.p[0], nc = g.p[1], cb = (nb + nc - 1) / nc;
        T fc = ZERO, sc = ZERO;
        for (int i = (int)nc - 1; i >= 0; i--) { fc = fc * T::from(1ULL << cb) + w[4 + i]; sc = sc * T::from(1ULL << cb) + w[4 + nc + i]; }
        out.push_back(fc - w[0]); out.push_back(sc - w[1]);
        T msd = ZERO;
        for (u32 i = 0; i < nc; i++) {
            const T &f = w[4 + i], &s = w[4 + nc + i];
            out.push_back(range_prod(f, 1 << cb)); out.push_back(range_prod(s, 1 << cb));
            T diff = s - f; const T &dummy = w[4 + 2 * nc + i], &eq = w[4 + 3 * nc + i], &inter = w[4 + 4 * nc + i];
            out.push_back(diff * dummy - (ONE - eq)); out.push_back(eq * diff); out.push_back(inter - eq * msd);
            msd = inter + (ONE - eq) * diff;
        }
        out.push_back(w[3] - msd);
        const T *bits = w + 4 + 5 * nc;
        for (u32 i = 0; i <= cb; i++) out.push_back(bits[i] * (ONE - bits[i]));
        T comb = ZERO; for (int i = (int)cb; i >= 0; i--) comb = comb * T::from(2) + bits[i];
        out.push_back(T::from(1ULL << cb) + w[3] - comb);
        out.push_back(w[2] - bits[cb]);
        return true; }
    }
    return false;
}
