// Per-point gate constraint evaluators for the quotient kernel (loop C of SURVEY.md section 8 / a8).
// Standard gates: formulas of upstream plonky2 gates as mirrored in gnark-plonky2-verifier/plonk/gates/
// (arithmetic_gate.go:60, arithmetic_extension_gate.go:59, multiplication_extension_gate.go:55, base_sum_gate.go:66,
// constant_gate.go:57, public_input_gate.go:32, reducing_gate.go:77, reducing_extension_gate.go:77,
// random_access_gate.go:131, poseidon_gate.go:92, poseidon_mds_gate.go:76, coset_interpolation_gate.go:151).
// Custom gates: crypto/plonky2_u32/src/gates/{arithmetic_u32.rs:108-167, add_many_u32.rs:109-147,
// subtraction_u32.rs:102-137, range_check_u32.rs:71-93, interleave_u32.rs:101-138, uninterleave_to_u32.rs:110-156,
// uninterleave_to_b32.rs:114-162, comparison.rs:118-195}.  Constraint order follows the reference exactly.
//
// Evaluation model: one thread = one point of the 8n coset.  A gate never materialises its constraint vector:
// each constraint c_i is folded on the fly into  S_k += c_i * alpha_k^i  (k = challenge index) with 192-bit lazy
// accumulators; the caller multiplies S_k by the gate's selector filter once.  Mathematically identical to
// reduce_with_powers over the summed filtered constraints (all arithmetic is exact in F_p).
#pragma once
#include "gl.cuh"
#include "poseidon.cuh"

enum GateKind {
    G_NOOP = 0, G_CONSTANT, G_PUBLIC_INPUT, G_ARITHMETIC, G_ARITHMETIC_EXT, G_MUL_EXT, G_BASE_SUM, G_REDUCING,
    G_REDUCING_EXT, G_RANDOM_ACCESS, G_POSEIDON_MDS, G_POSEIDON, G_COSET_INTERPOLATION,
    G_U32_ARITHMETIC, G_U32_ADD_MANY, G_U32_SUBTRACTION, G_U32_RANGE_CHECK, G_U32_INTERLEAVE,
    G_UNINTERLEAVE_U32, G_UNINTERLEAVE_B32, G_COMPARISON, G_NUM_KINDS
};

struct GateDesc {
    u32 kind;
    u32 p[4];
    u32 selector_index, group_start, group_end, row;
    u32 num_constraints;
    u32 extra_off;          // offset into the circuit's extra-constants array (coset interpolation weights)
};

#define MAX_CHALLENGES 2

#ifdef __CUDACC__
struct Wires {              // column-major matrix accessor at one LDE position
    const u64 *base; size_t stride, pos;
    __device__ __forceinline__ u64 operator[](u32 j) const { return __ldg(base + (size_t)j * stride + pos); }
};

struct ConstraintSink {     // S_k += c * alpha_k^idx
    const u64 *apow;        // [MAX_CHALLENGES][apow_stride] powers of alpha
    u32 apow_stride;
    acc192 acc[MAX_CHALLENGES];
    u32 idx;
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int k = 0; k < MAX_CHALLENGES; k++) { acc[k].lo = 0; acc[k].hi = 0; acc[k].top = 0; }
        idx = 0;
    }
    __device__ __forceinline__ void push(u64 c) {
#pragma unroll
        for (int k = 0; k < MAX_CHALLENGES; k++) acc192_mad(acc[k], c, __ldg(apow + k * apow_stride + idx));
        idx++;
    }
    __device__ __forceinline__ u64 result(int k) const { return acc192_reduce(acc[k]); }
};

struct alg2 { u64 a, b; };   // extension-algebra element over the base field: a + b X, X^2 = 7
__device__ __forceinline__ alg2 A2(u64 a, u64 b) { alg2 r; r.a = a; r.b = b; return r; }
__device__ __forceinline__ alg2 alg_add(alg2 x, alg2 y) { return A2(gl_add(x.a, y.a), gl_add(x.b, y.b)); }
__device__ __forceinline__ alg2 alg_sub(alg2 x, alg2 y) { return A2(gl_sub(x.a, y.a), gl_sub(x.b, y.b)); }
__device__ __forceinline__ alg2 alg_mul(alg2 x, alg2 y) {
    return A2(gl_add(gl_mul(x.a, y.a), gl_mul(7, gl_mul(x.b, y.b))), gl_add(gl_mul(x.a, y.b), gl_mul(x.b, y.a)));
}
__device__ __forceinline__ alg2 alg_scalar(u64 s, alg2 x) { return A2(gl_mul(s, x.a), gl_mul(s, x.b)); }
__device__ __forceinline__ alg2 W2(const Wires &w, u32 i) { return A2(w[i], w[i + 1]); }

// x(x-1)(x-2)(x-3) = t(t+2) with t = x(x-3): two multiplications
__device__ __forceinline__ u64 range_prod4(u64 x) {
    u64 t = gl_mul(x, gl_sub(x, 3));
    return gl_mul(t, gl_add(t, 2));
}
__device__ __forceinline__ u64 gl_dbl(u64 x) { return gl_add(x, x); }
__device__ __forceinline__ u64 gl_x4(u64 x) { return gl_dbl(gl_dbl(x)); }
__device__ __forceinline__ u64 range_prod2(u64 x) { return gl_mul(x, gl_sub(x, 1)); }
__device__ __forceinline__ u64 range_prod_n(u64 x, int n) { u64 p = 1; for (int k = 0; k < n; k++) p = gl_mul(p, gl_sub(x, (u64)k)); return p; }
__device__ __forceinline__ u64 sbox7c(u64 x) { return gl_canon(poseidon_sbox(x)); }

__device__ __forceinline__ void mds_canon(u64 s[12]) {
    poseidon_mds(s);
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
}

// PoseidonGate (poseidon_gate.go:92-181) evaluated with the compact permutation structure: the state lives in shared
// memory, the S-box layer is a rolled loop, all 30 rounds share one MDS.  Partial rounds use the textbook form: the
// S-box input of round r is (state[0] + RC[12r]) in both the textbook and the "fast" formulation the reference gate
// is written in (checked in tests/test_gate_parity.py), so every constraint value is identical.
static __device__ __noinline__ void eval_poseidon_gate_smem(const Wires &w, ConstraintSink &sink, u64 *st, const int stride) {
    {
        u64 swap = w[24];
        sink.push(gl_mul(swap, gl_sub(swap, 1)));
#pragma unroll
        for (int i = 0; i < 4; i++) {
            u64 lhs = w[i], rhs = w[i + 4], d = w[25 + i];
            sink.push(gl_sub(gl_mul(swap, gl_sub(rhs, lhs)), d));
            st[i * stride] = gl_add(lhs, d); st[(i + 4) * stride] = gl_sub(rhs, d);
        }
#pragma unroll
        for (int i = 8; i < 12; i++) st[i * stride] = w[i];
    }
    for (int r = 0; r < 30; r++) {
        const bool full = (r < 4) || (r >= 26);
        const u64 *rc = d_POSEIDON_RC + 12 * r;
        u64 s[12];
        if (full) {
            const int wbase = (r < 4) ? 29 + 12 * (r - 1) : 87 + 12 * (r - 26);
            for (int i = 0; i < 12; i++) {
                u64 x = gl_add(gl_canon(st[i * stride]), rc[i]);
                if (r != 0) { u64 sin = w[wbase + i]; sink.push(gl_sub(x, sin)); x = sin; }
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
            u64 sin = w[65 + (r - 4)];
            sink.push(gl_sub(gl_canon(s[0]), sin));
            s[0] = poseidon_sbox(sin);
        }
        poseidon_mds(s);
#pragma unroll
        for (int i = 0; i < 12; i++) st[i * stride] = s[i];
    }
    for (int i = 0; i < 12; i++) sink.push(gl_sub(gl_canon(st[i * stride]), w[12 + i]));
}

static __device__ __noinline__ void eval_coset_interpolation(const GateDesc &g, const Wires &w, const u64 *weights, ConstraintSink &sink) {
    const u32 sb = g.p[0], degree = g.p[1], np = 1u << sb, n_inter = (np - 2) / (degree - 1);
    const u32 start_inter = 1 + 2 * np + 4;
    u64 shift = w[0];
    alg2 ep = W2(w, 1 + 2 * np), sep = W2(w, start_inter + 4 * n_inter);
    alg2 t = alg_add(alg_scalar(gl_neg(shift), sep), ep);
    sink.push(t.a); sink.push(t.b);
    const u64 gen = gl_root_of_unity((int)sb);
    alg2 ev = A2(0, 0), prod = A2(1, 0);
    u64 x = 1;
    u32 next_break = degree, inter_i = 0;
    for (u32 i = 0; i < np; i++) {
        if (i == next_break && inter_i < n_inter) {
            alg2 ie = W2(w, start_inter + 2 * inter_i), ip = W2(w, start_inter + 2 * (n_inter + inter_i));
            alg2 d = alg_sub(ie, ev); sink.push(d.a); sink.push(d.b);
            d = alg_sub(ip, prod); sink.push(d.a); sink.push(d.b);
            ev = ie; prod = ip; inter_i++; next_break += degree - 1;
        }
        alg2 term = alg_sub(sep, A2(x, 0));
        alg2 wv = alg_scalar(__ldg(weights + i), W2(w, 1 + 2 * i));
        ev = alg_add(alg_mul(ev, term), alg_mul(wv, prod));
        prod = alg_mul(prod, term);
        x = gl_mul(x, gen);
    }
    alg2 d = alg_sub(W2(w, 1 + 2 * np + 2), ev); sink.push(d.a); sink.push(d.b);
}

// consts: constant columns AFTER the selector prefix.  KIND is a compile-time constant: every gate kind gets its own
// small kernel (k_gate<KIND> in plonk.cu) -- a monolithic all-gates kernel (26 k instructions) was instruction-fetch bound.
template <int KIND>
__device__ __forceinline__ void eval_gate(const GateDesc &g, const Wires &c, const Wires &w, const u64 *pi_hash,
                                          const u64 *extra, ConstraintSink &sink, u64 *smem_state, int smem_stride) {
    switch (KIND) {
    case G_NOOP: break;
    case G_CONSTANT: for (u32 i = 0; i < g.p[0]; i++) sink.push(gl_sub(c[i], w[i])); break;
    case G_PUBLIC_INPUT: for (u32 i = 0; i < 4; i++) sink.push(gl_sub(w[i], pi_hash[i])); break;
    case G_ARITHMETIC: {
        u64 c0 = c[0], c1 = c[1];
        for (u32 i = 0; i < g.p[0]; i++)
            sink.push(gl_sub(w[4 * i + 3], gl_add(gl_mul(gl_mul(w[4 * i], w[4 * i + 1]), c0), gl_mul(w[4 * i + 2], c1))));
        break; }
    case G_ARITHMETIC_EXT: {
        u64 c0 = c[0], c1 = c[1];
        for (u32 i = 0; i < g.p[0]; i++) {
            alg2 comp = alg_add(alg_scalar(c1, W2(w, 8 * i + 4)), alg_scalar(c0, alg_mul(W2(w, 8 * i), W2(w, 8 * i + 2))));
            alg2 d = alg_sub(W2(w, 8 * i + 6), comp); sink.push(d.a); sink.push(d.b);
        }
        break; }
    case G_MUL_EXT: {
        u64 c0 = c[0];
        for (u32 i = 0; i < g.p[0]; i++) {
            alg2 d = alg_sub(W2(w, 6 * i + 4), alg_scalar(c0, alg_mul(W2(w, 6 * i), W2(w, 6 * i + 2)))); sink.push(d.a); sink.push(d.b);
        }
        break; }
    case G_BASE_SUM: {
        const u32 n = g.p[0], base = g.p[1];
        u64 acc = 0;
        for (int i = (int)n - 1; i >= 0; i--) acc = gl_add(gl_mul(acc, base), w[1 + i]);
        sink.push(gl_sub(acc, w[0]));
        for (u32 i = 0; i < n; i++) sink.push(base == 2 ? range_prod2(w[1 + i]) : range_prod_n(w[1 + i], (int)base));
        break; }
    case G_REDUCING: case G_REDUCING_EXT: {
        const u32 n = g.p[0]; const bool ext = KIND == G_REDUCING_EXT;
        const u32 start_accs = 6 + (ext ? 2 * n : n);
        alg2 alpha = W2(w, 2), acc = W2(w, 4);
        for (u32 i = 0; i < n; i++) {
            alg2 coeff = ext ? W2(w, 6 + 2 * i) : A2(w[6 + i], 0);
            alg2 ai = (i == n - 1) ? W2(w, 0) : W2(w, start_accs + 2 * i);
            alg2 t = alg_sub(alg_add(alg_mul(acc, alpha), coeff), ai);
            sink.push(t.a); sink.push(t.b); acc = ai;
        }
        break; }
    case G_RANDOM_ACCESS: {
        const u32 bits = g.p[0], copies = g.p[1], extra_c = g.p[2], vec = 1u << bits;
        const u32 routed = (2 + vec) * copies + extra_c;
        for (u32 cp = 0; cp < copies; cp++) {
            const u32 base = (2 + vec) * cp;
            u64 b[6];
            for (u32 i = 0; i < bits && i < 6; i++) { b[i] = w[routed + cp * bits + i]; sink.push(gl_sub(gl_mul(b[i], b[i]), b[i])); }
            u64 rec = 0;
            for (int i = (int)bits - 1; i >= 0; i--) rec = gl_add(gl_add(rec, rec), b[i]);
            sink.push(gl_sub(rec, w[base]));
            // fold the list: selected = sum over items of item * prod(bit matches); do it level by level on the fly
            // (vec <= 64): recursive halving with a small local array
            u64 items[64];
            for (u32 i = 0; i < vec; i++) items[i] = w[base + 2 + i];
            u32 len = vec;
            for (u32 bi = 0; bi < bits; bi++) {
                for (u32 i = 0; i < len; i += 2) items[i / 2] = gl_add(items[i], gl_mul(b[bi], gl_sub(items[i + 1], items[i])));
                len >>= 1;
            }
            sink.push(gl_sub(items[0], w[base + 1]));
        }
        for (u32 i = 0; i < extra_c; i++) sink.push(gl_sub(c[i], w[(2 + vec) * copies + i]));
        break; }
    case G_POSEIDON_MDS: {
        const u32 C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
        for (int r = 0; r < 12; r++) {
            alg2 acc = A2(0, 0);
            for (int i = 0; i < 12; i++) acc = alg_add(acc, alg_scalar(C[i], W2(w, 2 * ((i + r) % 12))));
            if (r == 0) acc = alg_add(acc, alg_scalar(8, W2(w, 0)));
            alg2 d = alg_sub(W2(w, 24 + 2 * r), acc); sink.push(d.a); sink.push(d.b);
        }
        break; }
    case G_POSEIDON: eval_poseidon_gate_smem(w, sink, smem_state, smem_stride); break;
    case G_COSET_INTERPOLATION: eval_coset_interpolation(g, w, extra + g.extra_off, sink); break;
    case G_U32_ARITHMETIC: {
        const u32 n = g.p[0];
        for (u32 i = 0; i < n; i++) {
            u64 m0 = w[6 * i], m1 = w[6 * i + 1], ad = w[6 * i + 2], lo = w[6 * i + 3], hi = w[6 * i + 4], inv = w[6 * i + 5];
            u64 comp = gl_add(gl_mul(m0, m1), ad);
            sink.push(gl_mul(gl_sub(gl_mul(inv, gl_sub(0xFFFFFFFFULL, hi)), 1), lo));
            sink.push(gl_sub(gl_add(gl_mul(hi, 1ULL << 32), lo), comp));
            u64 cl = 0, ch = 0;
            for (int j = 31; j >= 0; j--) {
                u64 limb = w[6 * n + 32 * i + j];
                sink.push(range_prod4(limb));
                if (j < 16) cl = gl_add(gl_x4(cl), limb); else ch = gl_add(gl_x4(ch), limb);
            }
            sink.push(gl_sub(cl, lo)); sink.push(gl_sub(ch, hi));
        }
        break; }
    case G_U32_ADD_MANY: {
        const u32 na = g.p[0], n = g.p[1], r = na + 3;
        for (u32 i = 0; i < n; i++) {
            u64 comp = 0;
            for (u32 j = 0; j <= na; j++) comp = gl_add(comp, w[r * i + j]);
            u64 res = w[r * i + na + 1], carry = w[r * i + na + 2];
            sink.push(gl_sub(gl_add(gl_mul(carry, 1ULL << 32), res), comp));
            u64 cr = 0, cc = 0;
            for (int j = 17; j >= 0; j--) {
                u64 limb = w[r * n + 18 * i + j];
                sink.push(range_prod4(limb));
                if (j < 16) cr = gl_add(gl_x4(cr), limb); else cc = gl_add(gl_x4(cc), limb);
            }
            sink.push(gl_sub(cr, res)); sink.push(gl_sub(cc, carry));
        }
        break; }
    case G_U32_SUBTRACTION: {
        const u32 n = g.p[0];
        for (u32 i = 0; i < n; i++) {
            u64 x = w[5 * i], y = w[5 * i + 1], bi = w[5 * i + 2], res = w[5 * i + 3], bo = w[5 * i + 4];
            sink.push(gl_sub(res, gl_add(gl_sub(gl_sub(x, y), bi), gl_mul(1ULL << 32, bo))));
            u64 cl = 0;
            for (int j = 15; j >= 0; j--) { u64 limb = w[5 * n + 16 * i + j]; sink.push(range_prod4(limb)); cl = gl_add(gl_x4(cl), limb); }
            sink.push(gl_sub(cl, res));
            sink.push(gl_mul(bo, gl_sub(1, bo)));
        }
        break; }
    case G_U32_RANGE_CHECK: {
        const u32 n = g.p[0];
        for (u32 i = 0; i < n; i++) {
            u64 acc = 0;
            for (int j = 15; j >= 0; j--) acc = gl_add(gl_x4(acc), w[n + 16 * i + j]);
            sink.push(gl_sub(acc, w[i]));
            for (int j = 0; j < 16; j++) sink.push(range_prod4(w[n + 16 * i + j]));
        }
        break; }
    case G_U32_INTERLEAVE: {
        const u32 n = g.p[0];
        for (u32 i = 0; i < n; i++) {
            u64 cx = 0, ci = 0;
            for (int j = 0; j < 32; j++) { u64 b = w[2 * n + 32 * i + j]; cx = gl_add(gl_dbl(cx), b); ci = gl_add(gl_x4(ci), b); }
            sink.push(gl_sub(cx, w[2 * i])); sink.push(gl_sub(ci, w[2 * i + 1]));
            for (int j = 0; j < 32; j++) sink.push(range_prod2(w[2 * n + 32 * i + j]));
        }
        break; }
    case G_UNINTERLEAVE_U32: case G_UNINTERLEAVE_B32: {
        const u32 n = g.p[0]; const bool b32 = KIND == G_UNINTERLEAVE_B32;
        for (u32 i = 0; i < n; i++) {
            u64 ci = 0, ce = 0, co = 0;
            for (int j = 0; j < 64; j++) {
                u64 b = w[3 * n + 64 * i + j];
                ci = gl_add(gl_add(ci, ci), b);
                int jj = j >> 1;
                u64 coeff = b32 ? (1ULL << (2 * (31 - jj))) : (1ULL << (31 - jj));
                if (j & 1) co = gl_add(co, gl_mul(coeff, b)); else ce = gl_add(ce, gl_mul(coeff, b));
            }
            sink.push(gl_sub(ci, w[3 * i])); sink.push(gl_sub(ce, w[3 * i + 1])); sink.push(gl_sub(co, w[3 * i + 2]));
            for (int j = 0; j < 64; j++) sink.push(range_prod2(w[3 * n + 64 * i + j]));
        }
        break; }
    case G_COMPARISON: {
        const u32 nb = g.p[0], nc = g.p[1], cb = (nb + nc - 1) / nc;
        u64 fc = 0, sc = 0;
        for (int i = (int)nc - 1; i >= 0; i--) { fc = gl_add(gl_mul(fc, 1ULL << cb), w[4 + i]); sc = gl_add(gl_mul(sc, 1ULL << cb), w[4 + nc + i]); }
        sink.push(gl_sub(fc, w[0])); sink.push(gl_sub(sc, w[1]));
        u64 msd = 0;
        for (u32 i = 0; i < nc; i++) {
            u64 f = w[4 + i], s = w[4 + nc + i];
            sink.push(range_prod_n(f, 1 << cb)); sink.push(range_prod_n(s, 1 << cb));
            u64 diff = gl_sub(s, f), dummy = w[4 + 2 * nc + i], eq = w[4 + 3 * nc + i];
            sink.push(gl_sub(gl_mul(diff, dummy), gl_sub(1, eq)));
            sink.push(gl_mul(eq, diff));
            u64 inter = w[4 + 4 * nc + i];
            sink.push(gl_sub(inter, gl_mul(eq, msd)));
            msd = gl_add(inter, gl_mul(gl_sub(1, eq), diff));
        }
        u64 msdw = w[3];
        sink.push(gl_sub(msdw, msd));
        u64 comb = 0;
        for (int i = (int)cb; i >= 0; i--) comb = gl_add(gl_add(comb, comb), w[4 + 5 * nc + i]);
        for (u32 i = 0; i <= cb; i++) { u64 b = w[4 + 5 * nc + i]; sink.push(gl_mul(b, gl_sub(1, b))); }
        sink.push(gl_sub(gl_add(1ULL << cb, msdw), comb));
        sink.push(gl_sub(w[2], w[4 + 5 * nc + cb]));
        break; }
    default: break;
    }
}
#endif
