// Host run of the device gate evaluators of gates.cuh (eval_gate<KIND>: what the quotient kernels execute, with their lazy
// 128-bit recompositions and non-canonical products) against the generic evaluators of verify_gates.h (what b200_verify
// executes; pinned by the reference's golden proofs): for every gate kind, S_k = sum_i c_i alpha_k^i must agree on random
// and on small-valued rows.  Built and run by tests/test_host_emulations.py (nvcc host compile, no device needed).
#include "host_common.h"
#include "verify_gates.h"
#include <cstdio>
#include <cstdlib>
#include <string>
struct FpT {
    u64 v;
    static FpT from(u64 x) { FpT r; r.v = x % GL_P; return r; }
    FpT operator+(const FpT &o) const { FpT r; r.v = gl_add(v, o.v); return r; }
    FpT operator-(const FpT &o) const { FpT r; r.v = gl_sub(v, o.v); return r; }
    FpT operator*(const FpT &o) const { FpT r; r.v = gl_mul(v, o.v); return r; }
};
static u64 rnd() { return (((u64)rand() << 42) ^ ((u64)rand() << 21) ^ (u64)rand()) % GL_P; }
#define PH "_phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField>"
static const char *IDS[] = {
    "ConstantGate { num_consts: 2 }", "PublicInputGate", "ArithmeticGate { num_ops: 20 }", "ArithmeticExtensionGate { num_ops: 10 }",
    "MulExtensionGate { num_ops: 13 }", "BaseSumGate { num_limbs: 63 } + Base: 2", "BaseSumGate { num_limbs: 20 } + Base: 3", "ReducingGate { num_coeffs: 43 }",
    "ReducingExtensionGate { num_coeffs: 32 }", "RandomAccessGate { bits: 4, num_copies: 4, num_extra_constants: 2, " PH " }<D=2>",
    "PoseidonMdsGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>",
    "PoseidonGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>",
    "CosetInterpolationGate { subgroup_bits: 4, degree: 6, barycentric_weights: [17293822565076172801, 18374686475376656385, 18446744069413535745, 281474976645120, "
    "17592186044416, 256, 18446744000695107601, 18446744065119617025, 1152921504338411520, 72057594037927936, 1048576, 18446462594437939201, 18446726477228539905, "
    "18446744069414584065, 68719476720, 4294967296], " PH " }<D=2>",
    "U32ArithmeticGate { num_ops: 6, " PH " }", "U32ArithmeticGate { num_ops: 3, " PH " }", "U32AddManyGate { num_addends: 5, num_ops: 9, " PH " }",
    "U32AddManyGate { num_addends: 2, num_ops: 10, " PH " }", "U32AddManyGate { num_addends: 16, num_ops: 4, " PH " }", "U32SubtractionGate { num_ops: 11, " PH " }",
    "U32RangeCheckGate { num_input_limbs: 8, " PH " }", "U32InterleaveGate { num_ops: 6 }", "UninterleaveToU32Gate { num_ops: 3 }", "UninterleaveToB32Gate { num_ops: 3 }",
    "ComparisonGate { num_bits: 32, num_chunks: 16, " PH " }<D=2>",
};
template <int K> static void run_kind(const GateDesc &g, const Wires &c, const Wires &w, const u64 *pi, const u64 *extra, ConstraintSink &sink, u64 *st) {
    eval_gate<K>(g, c, w, pi, extra, sink, st, 1);
}
// fixture mode: argv[1] = text file, per case two lines: the gate id, then "expect_zero n w0 .. w(n-1)".  Every case is evaluated
// by the device evaluator (S_k under two random alphas must be 0 exactly when all constraints vanish) and by the generic one.
static int run_cases(const char *path) {
    FILE *f = fopen(path, "r"); if (!f) { printf("cannot open %s\n", path); return 1; }
    static char line[1 << 16];
    int ran = 0, skipped = 0;
    while (fgets(line, sizeof line, f)) {
        std::string id(line); while (!id.empty() && (id.back() == '\n' || id.back() == '\r')) id.pop_back();
        int expect = 0, n = 0;
        if (fscanf(f, "%d %d", &expect, &n) != 2) { printf("bad case header\n"); return 1; }
        std::vector<u64> wires(240, 0);
        for (int i = 0; i < n; i++) { unsigned long long v; if (fscanf(f, "%llu", &v) != 1) return 1; if (i < 240) wires[i] = (u64)v; }
        fgets(line, sizeof line, f);                            // rest of the line
        GateDesc g; std::vector<u64> extra;
        if (!parse_gate(id.c_str(), g, extra)) { skipped++; continue; }      // parameterisation outside the product's range (e.g. 8-bit comparison chunks)
        if (extra.empty()) extra.push_back(0);
        std::vector<u64> consts(4, 0), apow(2 * 600); u64 pi[4] = {0, 0, 0, 0};
        u64 alpha[2] = {rnd(), rnd()};
        for (int k = 0; k < 2; k++) { u64 p = 1; for (int i = 0; i < 600; i++) { apow[600 * k + i] = p; p = gl_mul(p, alpha[k]); } }
        std::vector<FpT> cw(240), cc(4), cons;
        for (int i = 0; i < 240; i++) cw[i] = FpT::from(wires[i]);
        for (int i = 0; i < 4; i++) cc[i] = FpT::from(0);
        if (!eval_gate_generic<FpT>(g, cc.data(), cw.data(), pi, extra.data(), cons)) return 1;
        bool gen_zero = true; for (auto &c : cons) if (c.v) gen_zero = false;
        Wires W{wires.data(), 1, 0}, C{consts.data(), 1, 0};
        ConstraintSink sink; sink.apow = apow.data(); sink.apow_stride = 600; sink.reset();
        u64 st[12];
        switch (g.kind) {
#define GK(K) case K: run_kind<K>(g, C, W, pi, extra.data(), sink, st); break;
        GK(G_U32_ARITHMETIC) GK(G_U32_ADD_MANY) GK(G_U32_SUBTRACTION) GK(G_U32_RANGE_CHECK) GK(G_U32_INTERLEAVE) GK(G_UNINTERLEAVE_U32)
        GK(G_UNINTERLEAVE_B32) GK(G_COMPARISON) GK(G_ARITHMETIC) GK(G_BASE_SUM)
#undef GK
        default: printf("unhandled kind in fixture\n"); return 1;
        }
        bool dev_zero = sink.result(0) == 0 && sink.result(1) == 0;
        if (gen_zero != (expect != 0) || dev_zero != (expect != 0)) { printf("CASE MISMATCH: %s expect_zero=%d generic=%d device=%d\n", id.c_str(), expect, gen_zero, dev_zero); return 1; }
        ran++;
    }
    fclose(f);
    printf("reference gate cases: %d evaluated as the reference expects by the device and generic evaluators (%d outside the supported parameter range)\n", ran, skipped);
    return 0;
}
// gate_degree() (gates.cuh) is what lets the quotient evaluate a gate on a sub-coset: every constraint, restricted to a random
// line through the space of wire / constant values, must be a polynomial of degree <= gate_degree in the line parameter
// ((D+1)-th finite difference == 0), and the bound should be tight for at least one constraint (D-th difference != 0).
static int check_degrees() {
    const int NW = 240, NCST = 4;
    int checked = 0;
    for (const char *id : IDS) {
        GateDesc g; std::vector<u64> extra;
        if (!parse_gate(id, g, extra)) { printf("parse failed: %s\n", id); return 1; }
        if (extra.empty()) extra.push_back(0);
        const int D = (int)gate_degree(g);
        if (D < 1 || D > 8) { printf("implausible degree %d for %s\n", D, id); return 1; }
        std::vector<u64> wa(NW), wb(NW), ca(NCST), cb(NCST); u64 pi[4];
        for (auto &x : wa) x = rnd(); for (auto &x : wb) x = rnd(); for (auto &x : ca) x = rnd(); for (auto &x : cb) x = rnd();
        for (auto &x : pi) x = rnd();
        std::vector<std::vector<FpT>> vals;                    // vals[t][i] = constraint i at parameter t
        for (int t = 0; t <= D + 1; t++) {
            std::vector<FpT> cw(NW), cc(NCST), cons;
            for (int i = 0; i < NW; i++) cw[i] = FpT::from(gl_add(wa[i], gl_mul(wb[i], (u64)t)));
            for (int i = 0; i < NCST; i++) cc[i] = FpT::from(gl_add(ca[i], gl_mul(cb[i], (u64)t)));
            if (!eval_gate_generic<FpT>(g, cc.data(), cw.data(), pi, extra.data(), cons)) { printf("generic eval failed: %s\n", id); return 1; }
            vals.push_back(cons);
        }
        auto diff = [&](int m, size_t i) {                      // m-th forward difference at 0
            u64 acc = 0, binom = 1;
            for (int j = 0; j <= m; j++) {
                u64 term = gl_mul(binom % GL_P, vals[j][i].v);
                acc = ((m - j) & 1) ? gl_sub(acc, term) : gl_add(acc, term);
                binom = binom * (u64)(m - j) / (u64)(j + 1);
            }
            return acc;
        };
        bool tight = false;
        for (size_t i = 0; i < vals[0].size(); i++) {
            if (diff(D + 1, i) != 0) { printf("DEGREE VIOLATION: constraint %zu of %s exceeds gate_degree = %d\n", i, id, D); return 1; }
            if (diff(D, i) != 0) tight = true;
        }
        if (!tight) { printf("gate_degree = %d is not tight for %s\n", D, id); return 1; }
        checked++;
    }
    printf("gate_degree() bounds every constraint and is tight for %d gate parameterisations\n", checked);
    return 0;
}
int main(int argc, char **argv) {
    srand(11);
    if (argc > 1) return run_cases(argv[1]);
    if (check_degrees()) return 1;
    const int NW = 240, NCST = 4;
    int checked = 0;
    for (const char *id : IDS) {
        GateDesc g; std::vector<u64> extra;
        if (!parse_gate(id, g, extra)) { printf("parse failed: %s\n", id); return 1; }
        if (extra.empty()) extra.push_back(0);
        for (int trial = 0; trial < 6; trial++) {
            std::vector<u64> wires(NW), consts(NCST), apow(2 * 600); u64 pi[4];
            for (auto &x : wires) x = trial == 1 ? (u64)(rand() % 4) : (trial == 2 ? GL_P - 1 - (u64)(rand() % 3) : (trial == 3 ? (u64)0xFFFFFFFF00000000ULL : rnd()));
            for (auto &x : consts) x = rnd();
            for (auto &x : pi) x = rnd();
            u64 alpha[2] = {rnd(), rnd()};
            for (int k = 0; k < 2; k++) { u64 p = 1; for (int i = 0; i < 600; i++) { apow[600 * k + i] = p; p = gl_mul(p, alpha[k]); } }
            // generic evaluation (the verifier's)
            std::vector<FpT> cw(NW), cc(NCST), cons;
            for (int i = 0; i < NW; i++) cw[i] = FpT::from(wires[i]);
            for (int i = 0; i < NCST; i++) cc[i] = FpT::from(consts[i]);
            if (!eval_gate_generic<FpT>(g, cc.data(), cw.data(), pi, extra.data(), cons) || cons.size() != g.num_constraints) { printf("generic eval failed: %s\n", id); return 1; }
            u64 want[2] = {0, 0};
            for (int k = 0; k < 2; k++) for (size_t i = 0; i < cons.size(); i++) want[k] = gl_add(want[k], gl_mul(cons[i].v, apow[600 * k + i]));
            // device evaluator on the host: one "point" (stride irrelevant, pos 0, one column per wire)
            Wires W{wires.data(), 1, 0}, C{consts.data(), 1, 0};
            ConstraintSink sink; sink.apow = apow.data(); sink.apow_stride = 600; sink.reset();
            u64 st[12];
            switch (g.kind) {
#define GK(K) case K: run_kind<K>(g, C, W, pi, extra.data(), sink, st); break;
            GK(G_CONSTANT) GK(G_PUBLIC_INPUT) GK(G_ARITHMETIC) GK(G_ARITHMETIC_EXT) GK(G_MUL_EXT) GK(G_BASE_SUM) GK(G_REDUCING)
            GK(G_REDUCING_EXT) GK(G_RANDOM_ACCESS) GK(G_POSEIDON_MDS) GK(G_POSEIDON) GK(G_COSET_INTERPOLATION) GK(G_U32_ARITHMETIC)
            GK(G_U32_ADD_MANY) GK(G_U32_SUBTRACTION) GK(G_U32_RANGE_CHECK) GK(G_U32_INTERLEAVE) GK(G_UNINTERLEAVE_U32)
            GK(G_UNINTERLEAVE_B32) GK(G_COMPARISON)
#undef GK
            default: printf("unhandled kind\n"); return 1;
            }
            if (sink.idx != g.num_constraints) { printf("constraint count mismatch for %s: %u vs %u\n", id, sink.idx, g.num_constraints); return 1; }
            for (int k = 0; k < 2; k++) if (sink.result(k) != want[k]) { printf("MISMATCH %s trial %d challenge %d\n", id, trial, k); return 1; }
            checked++;
        }
    }
    printf("device gate evaluators == generic evaluators on %d (gate, row) cases\n", checked);
}
