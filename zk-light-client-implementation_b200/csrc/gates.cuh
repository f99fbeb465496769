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

// Degree of a gate's constraint polynomials in the wire/constant columns (upstream Gate::degree(); the u32 gates:
// crypto/plonky2_u32/src/gates/*.rs `fn degree`).  Used by the quotient to evaluate low-degree gates on a sub-coset only.
static inline u32 gate_degree(const GateDesc &g) {
    switch (g.kind) {
    case G_NOOP: return 0;
    case G_CONSTANT: case G_PUBLIC_INPUT: case G_POSEIDON_MDS: return 1;
    case G_ARITHMETIC: case G_ARITHMETIC_EXT: case G_MUL_EXT: return 3;
    case G_BASE_SUM: return g.p[1];
    case G_REDUCING: case G_REDUCING_EXT: return 2;
    case G_RANDOM_ACCESS: return g.p[0] + 1;
    case G_POSEIDON: return 7;
    case G_COSET_INTERPOLATION: return g.p[1];
    case G_U32_ARITHMETIC: case G_U32_ADD_MANY: case G_U32_SUBTRACTION: case G_U32_RANGE_CHECK: return 4;
    case G_U32_INTERLEAVE: case G_UNINTERLEAVE_U32: case G_UNINTERLEAVE_B32: return 2;
    case G_COMPARISON: { u32 d = 1u << ((g.p[0] + g.p[1] - 1) / g.p[1]); return d < 3 ? 3 : d; }     // eq * (inter + (1 - eq) * diff) is cubic
    default: return 1u << 30;
    }
}

// Everything below compiles for the device (the quotient kernels of plonk.cu) AND for the host: tests/gate_eval_emulation.cu
// runs these evaluators on the CPU against the generic evaluators of verify_gates.h (which b200_verify uses and the
// reference's golden proofs pin), so the arithmetic shortcuts below are checked without a GPU.
#if defined(__CUDACC__) || defined(B200_GATES_HOST)
#ifdef __CUDA_ARCH__
#define GATE_LD(p) __ldg(p)
#else
#define GATE_LD(p) (*(p))
#endif
#ifdef __CUDACC__
#define GATE_NOINLINE static __host__ __device__ __noinline__
#else
#define GATE_NOINLINE static
#endif

struct Wires {              // column-major matrix accessor at one LDE position
    const u64 *base; size_t stride, pos;
    GL_HD u64 operator[](u32 j) const { return GATE_LD(base + (size_t)j * stride + pos); }
    // B consecutive columns at once: the loads are independent, so B of them are in flight per thread (the round-1 kernels
    // walked the columns one dependent load at a time: 33-53 % long_scoreboard stalls, profiles/ncu_gates_r01_v9.jsonl)
    template <int B> GL_HD void load(u32 j0, u64 (&out)[B]) const {
        const u64 *p = base + (size_t)j0 * stride + pos;
#pragma unroll
        for (int k = 0; k < B; k++) out[k] = GATE_LD(p + (size_t)k * stride);
    }
};

struct ConstraintSink {     // S_k += c * alpha_k^idx
    const u64 *apow;        // [MAX_CHALLENGES][apow_stride] powers of alpha
    u32 apow_stride;
    acc192 acc[MAX_CHALLENGES];
    u32 idx;
    GL_HD void reset() {
#pragma unroll
        for (int k = 0; k < MAX_CHALLENGES; k++) { acc[k].lo = 0; acc[k].hi = 0; acc[k].top = 0; }
        idx = 0;
    }
    // c: ANY 64-bit representative of the constraint value (the product is exact, the accumulator is reduced at the end)
    GL_HD void push(u64 c) {
#pragma unroll
        for (int k = 0; k < MAX_CHALLENGES; k++) acc192_mad(acc[k], c, GATE_LD(apow + k * apow_stride + idx));
        idx++;
    }
    // same, at an explicit constraint index (lets a gate produce its constraints in one pass over the wires in any order)
    GL_HD void push_at(u64 c, u32 at) {
#pragma unroll
        for (int k = 0; k < MAX_CHALLENGES; k++) acc192_mad(acc[k], c, GATE_LD(apow + k * apow_stride + at));
    }
    GL_HD u64 result(int k) const { return acc192_reduce(acc[k]); }
};

struct alg2 { u64 a, b; };   // extension-algebra element over the base field: a + b X, X^2 = 7
GL_HD alg2 A2(u64 a, u64 b) { alg2 r; r.a = a; r.b = b; return r; }
GL_HD alg2 alg_add(alg2 x, alg2 y) { return A2(gl_add(x.a, y.a), gl_add(x.b, y.b)); }
GL_HD alg2 alg_sub(alg2 x, alg2 y) { return A2(gl_sub(x.a, y.a), gl_sub(x.b, y.b)); }
GL_HD alg2 alg_mul(alg2 x, alg2 y) {
    return A2(gl_add(gl_mul(x.a, y.a), gl_mul(7, gl_mul(x.b, y.b))), gl_add(gl_mul(x.a, y.b), gl_mul(x.b, y.a)));
}
GL_HD alg2 alg_scalar(u64 s, alg2 x) { return A2(gl_mul(s, x.a), gl_mul(s, x.b)); }
GL_HD alg2 W2(const Wires &w, u32 i) { return A2(w[i], w[i + 1]); }

// ---- lazy helpers: results are arbitrary 64-bit representatives (fine for ConstraintSink::push and as multiplicands) ----
// x + k for a small constant k, x any 64-bit representative
GL_HD u64 gl_add_small_nc(u64 x, u32 k) { u64 s = x + k; if (s < k) s += GL_EPS; return s; }
// x(x-1)(x-2)(x-3) = t(t+2) with t = x(x-3): two multiplications; x canonical
GL_HD u64 range_prod4(u64 x) {
    u64 t = gl_mul_nc(x, gl_sub(x, 3));
    return gl_mul_nc(t, gl_add_small_nc(t, 2));
}
GL_HD u64 range_prod2(u64 x) { return gl_mul_nc(x, gl_sub(x, 1)); }
GL_HD u64 range_prod_n(u64 x, int n) { u64 p = 1; for (int k = 0; k < n; k++) p = gl_mul(p, gl_sub(x, (u64)k)); return p; }
GL_HD u64 sbox7c(u64 x) { return gl_canon(poseidon_sbox(x)); }

// Exact integer accumulator for recompositions sum_j limb_j * B^j (limbs are arbitrary field elements on the coset, < 2^64):
// Horner steps are a 128-bit shift-and-add (6 instructions) instead of two modular doublings and a modular add (24);
// one reduction at the end.  The caller keeps the running value below 2^128 (bounds at each use).
struct acc128 { u64 lo, hi; };
GL_HD void acc128_shl_add(acc128 &A, int sh, u64 x) {            // A = (A << sh) + x
    A.hi = (A.hi << sh) | (A.lo >> (64 - sh)); A.lo <<= sh;
    u64 l = A.lo + x; A.hi += l < x; A.lo = l;
}
GL_HD void acc128_add_shifted(acc128 &A, u64 x, int sh) {       // A += x << sh   (0 <= sh < 64)
    u64 xl = x << sh, xh = sh ? x >> (64 - sh) : 0;
    u64 l = A.lo + xl; A.hi += xh + (l < xl); A.lo = l;
}
GL_HD u64 acc128_reduce(const acc128 &A) { return gl_canon(gl_reduce128_nc(A.lo, A.hi)); }

// PoseidonGate (poseidon_gate.go:92-181) evaluated with the structure of the Merkle kernels' permutation: the state lives in
// shared memory, the S-box layer is a rolled loop, the linear layers run on limb triples.  The S-box input of a partial
// round is (state[0] + RC[12r]) in the textbook form, in the "fast" formulation the reference gate is written in (checked in
// tests/test_gate_parity.py) and in the pushed-constant limb form used here, so every constraint value is identical.
GATE_NOINLINE void eval_poseidon_gate_smem(const Wires &w, ConstraintSink &sink, u64 *st, const int stride) {
    {
        u64 swap = w[24];
        sink.push(gl_mul_nc(swap, gl_sub(swap, 1)));
#pragma unroll
        for (int i = 0; i < 4; i++) {
            u64 lhs = w[i], rhs = w[i + 4], d = w[25 + i];
            sink.push(gl_sub(gl_mul(swap, gl_sub(rhs, lhs)), d));
            st[i * stride] = gl_add(lhs, d); st[(i + 4) * stride] = gl_sub(rhs, d);
        }
#pragma unroll
        for (int i = 8; i < 12; i++) st[i * stride] = w[i];
    }
    // The rounds follow poseidon_permute_limb (poseidon.cuh) step for step -- limb-resident linear layers, constants of the
    // partial rounds pushed forward -- with the gate's hooks: every S-box input after round 0 is constrained against its
    // wire and the WIRE goes through the S-box.  The constant pushing is an identity of the linear layers for arbitrary
    // S-box outputs, so each constraint is the same field element as in the textbook form (host check:
    // tests/gate_eval_emulation.cu on random rows, where no constraint vanishes).
    i32 l0[12], l1[12], l2[12];
#pragma unroll
    for (int i = 0; i < 12; i++) pl_split(st[i * stride], l0[i], l1[i], l2[i]);
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
            const pl_c4 *rc = PLT(POSEIDON_L_RC) + 12 * (4 * phase + r);
            const bool constrained = phase == 1 || r != 0;
            const int wbase = phase == 0 ? 29 + 12 * (r - 1) : 87 + 12 * r;
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const pl_c4 c = rc[i];
                u64 x = pl_recombine((u32)l0[i] + c.a, (u32)l1[i] + c.b, (u32)l2[i] + c.c);
                if (constrained) { const u64 sin = w[wbase + i]; sink.push(gl_sub(gl_canon(x), sin)); x = sin; }
                st[i * stride] = x;
            }
#pragma unroll 2
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
                const u64 sin = w[65 + r];
                sink.push(gl_sub(gl_canon(pl_recombine((u32)l0[0] + c.a, (u32)l1[0] + c.b, (u32)l2[0] + c.c)), sin));
                pl_split(poseidon_sbox4(sin), l0[0], l1[0], l2[0]);
                pl_mds(l0, l1, l2);
#pragma unroll
                for (int i = 1; i < 12; i++) pl_normalise(l0[i], l1[i], l2[i]);
            }
        }
    }
    const pl_c4 fin = POSEIDON_L_FIN_INIT;
#pragma unroll
    for (int i = 0; i < 12; i++)
        sink.push(gl_sub(gl_canon(pl_recombine((u32)l0[i] + fin.a, (u32)l1[i] + fin.b, (u32)l2[i] + fin.c)), w[12 + i]));
}

GATE_NOINLINE void eval_coset_interpolation(const GateDesc &g, const Wires &w, const u64 *weights, ConstraintSink &sink) {
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
        alg2 wv = alg_scalar(GATE_LD(weights + i), W2(w, 1 + 2 * i));
        ev = alg_add(alg_mul(ev, term), alg_mul(wv, prod));
        prod = alg_mul(prod, term);
        x = gl_mul(x, gen);
    }
    alg2 d = alg_sub(W2(w, 1 + 2 * np + 2), ev); sink.push(d.a); sink.push(d.b);
}

// 2-bit-limb blocks shared by the u32 gates: NB batches of 8 limb wires, columns [col, col + 8 NB), processed highest
// limb first as the reference iterates (`.rev()`): pushes the range-check product of every limb and Horner-accumulates
// sum_j limb_j 4^j exactly (< 2^64 * 4^16 / 3 < 2^96 per 16 limbs).  The first SPLIT batches go to `hi`, the rest to `lo`.
// Loads are software-pipelined: the next batch of 8 columns is in flight while the current one is evaluated (the kernels
// are otherwise latency-bound on their column walks: ncu long_scoreboard 40 %, profiles/ncu_gates_r02_v2.jsonl); the batch
// loop stays rolled (two batches per iteration) so that the kernel body fits the instruction cache.
GL_HD void limbs4_batch(const u64 (&L)[8], ConstraintSink &sink, acc128 &A) {
#pragma unroll
    for (int k = 7; k >= 0; k--) { sink.push(range_prod4(L[k])); acc128_shl_add(A, 2, L[k]); }
}
template <int NB, int SPLIT>
GL_HD void limbs4_blocks(const Wires &w, u32 col, ConstraintSink &sink, acc128 &hi, acc128 &lo) {
    static_assert(NB % 2 == 0 && SPLIT % 2 == 0, "batches are processed in pairs");
    u64 A[8], B[8];
    acc128 acc = {0, 0};
    w.load<8>(col + 8 * (NB - 1), A);
#pragma unroll 1
    for (int b = NB - 1; b >= 1; b -= 2) {
        w.load<8>(col + 8 * (b - 1), B);
        limbs4_batch(A, sink, acc);
        if (b >= 2) w.load<8>(col + 8 * (b - 2), A);
        limbs4_batch(B, sink, acc);
        if (SPLIT && b - 1 == NB - SPLIT) { hi = acc; acc.lo = 0; acc.hi = 0; }
    }
    lo = acc;
}

// consts: constant columns AFTER the selector prefix.  KIND is a compile-time constant: every gate kind gets its own
// small kernel (k_gate<KIND> in plonk.cu) -- a monolithic all-gates kernel (26 k instructions) was instruction-fetch bound.
template <int KIND>
GL_HD void eval_gate(const GateDesc &g, const Wires &c, const Wires &w, const u64 *pi_hash,
                     const u64 *extra, ConstraintSink &sink, u64 *smem_state, int smem_stride) {
    switch (KIND) {
    case G_NOOP: break;
    case G_CONSTANT: for (u32 i = 0; i < g.p[0]; i++) sink.push(gl_sub(c[i], w[i])); break;
    case G_PUBLIC_INPUT: for (u32 i = 0; i < 4; i++) sink.push(gl_sub(w[i], pi_hash[i])); break;
    case G_ARITHMETIC: {
        u64 c0 = c[0], c1 = c[1];
        for (u32 i = 0; i < g.p[0]; i++) {
            u64 x[4]; w.load<4>(4 * i, x);
            sink.push(gl_sub(x[3], gl_add(gl_mul(gl_mul(x[0], x[1]), c0), gl_mul(x[2], c1))));
        }
        break; }
    case G_ARITHMETIC_EXT: {
        u64 c0 = c[0], c1 = c[1];
        for (u32 i = 0; i < g.p[0]; i++) {
            u64 x[8]; w.load<8>(8 * i, x);
            alg2 comp = alg_add(alg_scalar(c1, A2(x[4], x[5])), alg_scalar(c0, alg_mul(A2(x[0], x[1]), A2(x[2], x[3]))));
            alg2 d = alg_sub(A2(x[6], x[7]), comp); sink.push(d.a); sink.push(d.b);
        }
        break; }
    case G_MUL_EXT: {
        u64 c0 = c[0];
        for (u32 i = 0; i < g.p[0]; i++) {
            u64 x[6]; w.load<6>(6 * i, x);
            alg2 d = alg_sub(A2(x[4], x[5]), alg_scalar(c0, alg_mul(A2(x[0], x[1]), A2(x[2], x[3])))); sink.push(d.a); sink.push(d.b);
        }
        break; }
    case G_BASE_SUM: {
        const u32 n = g.p[0], base = g.p[1];
        if (base == 2 && n <= 64) {
            // sum_i limb_i 2^i as an exact integer (< 2^64 * 2^n: split at 32 limbs so that each half stays < 2^96)
            acc128 hi = {0, 0}, lo = {0, 0};
            for (int i = (int)n - 1; i >= 32; i--) acc128_shl_add(hi, 1, w[1 + i]);
            for (int i = (int)(n < 32 ? n : 32) - 1; i >= 0; i--) acc128_shl_add(lo, 1, w[1 + i]);
            u64 acc = gl_add(gl_mul(acc128_reduce(hi), 1ULL << 32), acc128_reduce(lo));
            sink.push(gl_sub(acc, w[0]));
            for (u32 i = 0; i < n; i++) sink.push(range_prod2(w[1 + i]));
        } else {
            u64 acc = 0;
            for (int i = (int)n - 1; i >= 0; i--) acc = gl_add(gl_mul(acc, base), w[1 + i]);
            sink.push(gl_sub(acc, w[0]));
            for (u32 i = 0; i < n; i++) sink.push(base == 2 ? range_prod2(w[1 + i]) : range_prod_n(w[1 + i], (int)base));
        }
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
            for (u32 i = 0; i < bits && i < 6; i++) { b[i] = w[routed + cp * bits + i]; sink.push(range_prod2(b[i])); }
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
            u64 r[6]; w.load<6>(6 * i, r);              // m0, m1, addend, out_lo, out_hi, inverse
            u64 comp = gl_add(gl_mul(r[0], r[1]), r[2]);
            sink.push(gl_mul_nc(gl_sub(gl_mul(r[5], gl_sub(0xFFFFFFFFULL, r[4])), 1), r[3]));
            sink.push(gl_sub(gl_add(gl_mul(r[4], 1ULL << 32), r[3]), comp));
            acc128 ch = {0, 0}, cl = {0, 0};
            limbs4_blocks<4, 2>(w, 6 * n + 32 * i, sink, ch, cl);
            sink.push(gl_sub(acc128_reduce(cl), r[3])); sink.push(gl_sub(acc128_reduce(ch), r[4]));
        }
        break; }
    case G_U32_ADD_MANY: {
        const u32 na = g.p[0], n = g.p[1], r = na + 3;
        for (u32 i = 0; i < n; i++) {
            acc128 sum = {0, 0};                         // <= 17 addends < 2^64 each
            for (u32 j = 0; j <= na; j++) acc128_add_shifted(sum, w[r * i + j], 0);
            u64 res = w[r * i + na + 1], carry = w[r * i + na + 2];
            sink.push(gl_sub(gl_add(gl_mul(carry, 1ULL << 32), res), acc128_reduce(sum)));
            acc128 cc = {0, 0}, cr = {0, 0};
            const u32 lb = r * n + 18 * i;
            { u64 L[2]; w.load<2>(lb + 16, L);
              sink.push(range_prod4(L[1])); acc128_shl_add(cc, 2, L[1]); sink.push(range_prod4(L[0])); acc128_shl_add(cc, 2, L[0]); }
            { acc128 unused = {0, 0}; limbs4_blocks<2, 0>(w, lb, sink, unused, cr); }
            sink.push(gl_sub(acc128_reduce(cr), res)); sink.push(gl_sub(acc128_reduce(cc), carry));
        }
        break; }
    case G_U32_SUBTRACTION: {
        const u32 n = g.p[0];
        for (u32 i = 0; i < n; i++) {
            u64 r[5]; w.load<5>(5 * i, r);              // x, y, borrow_in, out, borrow_out
            sink.push(gl_sub(r[3], gl_add(gl_sub(gl_sub(r[0], r[1]), r[2]), gl_mul(1ULL << 32, r[4]))));
            acc128 cl = {0, 0}, unused = {0, 0};
            limbs4_blocks<2, 0>(w, 5 * n + 16 * i, sink, unused, cl);
            sink.push(gl_sub(acc128_reduce(cl), r[3]));
            sink.push(gl_mul_nc(r[4], gl_sub(1, r[4])));
        }
        break; }
    case G_U32_RANGE_CHECK: {
        const u32 n = g.p[0];
        for (u32 i = 0; i < n; i++) {
            u64 L[16]; w.load<16>(n + 16 * i, L);
            acc128 acc = {0, 0};
#pragma unroll
            for (int j = 15; j >= 0; j--) acc128_shl_add(acc, 2, L[j]);
            sink.push(gl_sub(acc128_reduce(acc), w[i]));
#pragma unroll
            for (int j = 0; j < 16; j++) sink.push(range_prod4(L[j]));
        }
        break; }
    case G_U32_INTERLEAVE: {
        const u32 n = g.p[0];
        for (u32 i = 0; i < n; i++) {
            // one pass over the 32 bit wires: the two recomposition constraints come first in the reference's order, so the
            // range checks are pushed at their explicit indices and the recompositions afterwards
            acc128 cx = {0, 0}, ci = {0, 0};             // cx < 2^64 * 2^32, ci < 2^64 * 4^32 / 3 < 2^127
            const u32 lb = 2 * n + 32 * i, at0 = sink.idx;
            u64 A[8], B[8];
            w.load<8>(lb, A);
#pragma unroll 1
            for (int jb = 0; jb < 32; jb += 16) {
                w.load<8>(lb + jb + 8, B);
#pragma unroll
                for (int k = 0; k < 8; k++) { acc128_shl_add(cx, 1, A[k]); acc128_shl_add(ci, 2, A[k]); sink.push_at(range_prod2(A[k]), at0 + 2 + jb + k); }
                if (jb == 0) w.load<8>(lb + 16, A);
#pragma unroll
                for (int k = 0; k < 8; k++) { acc128_shl_add(cx, 1, B[k]); acc128_shl_add(ci, 2, B[k]); sink.push_at(range_prod2(B[k]), at0 + 10 + jb + k); }
            }
            sink.push_at(gl_sub(acc128_reduce(cx), w[2 * i]), at0); sink.push_at(gl_sub(acc128_reduce(ci), w[2 * i + 1]), at0 + 1);
            sink.idx = at0 + 34;
        }
        break; }
    case G_UNINTERLEAVE_U32: case G_UNINTERLEAVE_B32: {
        const u32 n = g.p[0]; const bool b32 = KIND == G_UNINTERLEAVE_B32;
        for (u32 i = 0; i < n; i++) {
            // x_interleaved = sum_j b_j 2^(63-j): two exact halves of 32 bits each (< 2^96), joined modularly;
            // evens / odds: sum over 32 bits with weights 2^(31-jj) (u32) or 4^(31-jj) (b32): < 2^64 * 2^63 exact.
            // One pass over the 64 bit wires, range checks pushed at their explicit indices (after the three recompositions).
            acc128 ih = {0, 0}, il = {0, 0}, ce = {0, 0}, co = {0, 0};
            const u32 lb = 3 * n + 64 * i, at0 = sink.idx;
            u64 A[8], B[8];
            w.load<8>(lb, A);
#pragma unroll 1
            for (int jb = 0; jb < 64; jb += 16) {
                w.load<8>(lb + jb + 8, B);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (jb < 32) acc128_shl_add(ih, 1, A[k]); else acc128_shl_add(il, 1, A[k]);
                    if (k & 1) acc128_shl_add(co, b32 ? 2 : 1, A[k]); else acc128_shl_add(ce, b32 ? 2 : 1, A[k]);
                    sink.push_at(range_prod2(A[k]), at0 + 3 + jb + k);
                }
                if (jb + 16 < 64) w.load<8>(lb + jb + 16, A);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    if (jb < 32) acc128_shl_add(ih, 1, B[k]); else acc128_shl_add(il, 1, B[k]);
                    if (k & 1) acc128_shl_add(co, b32 ? 2 : 1, B[k]); else acc128_shl_add(ce, b32 ? 2 : 1, B[k]);
                    sink.push_at(range_prod2(B[k]), at0 + 11 + jb + k);
                }
            }
            u64 cint = gl_add(gl_mul(acc128_reduce(ih), 1ULL << 32), acc128_reduce(il));
            sink.push_at(gl_sub(cint, w[3 * i]), at0); sink.push_at(gl_sub(acc128_reduce(ce), w[3 * i + 1]), at0 + 1);
            sink.push_at(gl_sub(acc128_reduce(co), w[3 * i + 2]), at0 + 2);
            sink.idx = at0 + 67;
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
            sink.push(gl_mul_nc(eq, diff));
            u64 inter = w[4 + 4 * nc + i];
            sink.push(gl_sub(inter, gl_mul(eq, msd)));
            msd = gl_add(inter, gl_mul(gl_sub(1, eq), diff));
        }
        u64 msdw = w[3];
        sink.push(gl_sub(msdw, msd));
        u64 comb = 0;
        for (int i = (int)cb; i >= 0; i--) comb = gl_add(gl_add(comb, comb), w[4 + 5 * nc + i]);
        for (u32 i = 0; i <= cb; i++) { u64 b = w[4 + 5 * nc + i]; sink.push(gl_mul_nc(b, gl_sub(1, b))); }
        sink.push(gl_sub(gl_add(1ULL << cb, msdw), comb));
        sink.push(gl_sub(w[2], w[4 + 5 * nc + cb]));
        break; }
    default: break;
    }
}
#endif
