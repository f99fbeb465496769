// The vanishing-polynomial identity at zeta, host side:  sum_t alpha^t * term_t(zeta) == Z_H(zeta) * sum_k zeta^{nk} q_k(zeta)
// for every challenge, with the terms in plonky2's order (L_0 * (Z - 1), the partial-product checks, the filtered gate
// constraints).  Shared by b200_verify (verifier.cu) and by b200_prove (prover.cu), which checks its own openings before it
// serialises a proof: a witness that violates a GATE constraint (but no copy constraint) still yields low-degree committed
// polynomials, so only this identity can tell.  Mirrors gnark-plonky2-verifier/plonk/plonk.go:121-250 and
// plonk/gates/evaluate_gates.go:33-75.
#pragma once
#include "verify_gates.h"

struct X2 {            // F_p^2 with operators, for the generic gate evaluators
    gl2 v;
    static X2 from(u64 x) { X2 r; r.v = gl2_make(x % GL_P, 0); return r; }
    X2 operator+(const X2 &o) const { X2 r; r.v = gl2_add(v, o.v); return r; }
    X2 operator-(const X2 &o) const { X2 r; r.v = gl2_sub(v, o.v); return r; }
    X2 operator*(const X2 &o) const { X2 r; r.v = gl2_mul(v, o.v); return r; }
};
static X2 mk(gl2 v) { X2 r; r.v = v; return r; }


struct VanishingInput {
    u32 num_wires, num_routed, num_constants, num_challenges, qdf, npp, num_gate_constraints, num_selector_groups; int degree_bits;
    const u64 *k_is; const std::vector<GateDesc> *gates; const u64 *extra; const u64 *pi_hash;
    const u64 *betas, *gammas, *alphas; gl2 zeta;
    const gl2 *o_const, *o_sig, *o_wires, *o_zs, *o_zs_next, *o_pp, *o_quot;
};
// nullptr if the identity holds, else a static reason string
static inline const char *vanishing_identity_at_zeta(const VanishingInput &in) {
    const u32 W = in.num_wires, R = in.num_routed, NC = in.num_constants, nch = in.num_challenges, qdf = in.qdf, npp = in.npp, nsel = in.num_selector_groups;
    const size_t n = size_t(1) << in.degree_bits;
    const gl2 zeta = in.zeta;
    gl2 zeta_pow_n = zeta; for (int i = 0; i < in.degree_bits; i++) zeta_pow_n = gl2_mul(zeta_pow_n, zeta_pow_n);
    std::vector<X2> consts(NC), wires(W);
    for (u32 i = 0; i < NC; i++) consts[i] = mk(in.o_const[i]);
    for (u32 i = 0; i < W; i++) wires[i] = mk(in.o_wires[i]);
    std::vector<X2> gate_terms(in.num_gate_constraints, X2::from(0));
    for (auto &g : *in.gates) {
        X2 sel = consts[g.selector_index], filt = X2::from(1);
        for (u32 r = g.group_start; r < g.group_end; r++) if (r != g.row) filt = filt * (X2::from(r) - sel);
        if (nsel > 1) filt = filt * (X2::from(0xFFFFFFFFULL) - sel);
        std::vector<X2> cons;
        if (!eval_gate_generic<X2>(g, consts.data() + nsel, wires.data(), in.pi_hash, in.extra, cons)) return "gate evaluation failed";
        if (cons.size() != g.num_constraints || cons.size() > in.num_gate_constraints) return "gate constraint count mismatch";
        for (size_t i = 0; i < cons.size(); i++) gate_terms[i] = gate_terms[i] + cons[i] * filt;
    }
    gl2 one = gl2_make(1, 0);
    gl2 l0 = gl2_mul(gl2_sub(zeta_pow_n, one), gl2_inv(gl2_scale(gl2_sub(zeta, one), (u64)n % GL_P)));
    std::vector<gl2> terms;
    for (u32 i = 0; i < nch; i++) terms.push_back(gl2_mul(l0, gl2_sub(in.o_zs[i], one)));
    for (u32 i = 0; i < nch; i++) {
        std::vector<gl2> accs; accs.push_back(in.o_zs[i]); for (u32 k = 0; k < npp; k++) accs.push_back(in.o_pp[i * npp + k]); accs.push_back(in.o_zs_next[i]);
        for (u32 k = 0; k <= npp; k++) {
            gl2 num = one, den = one;
            for (u32 j = k * qdf; j < (k + 1) * qdf && j < R; j++) {
                gl2 wg = gl2_add_base(in.o_wires[j], in.gammas[i]);
                num = gl2_mul(num, gl2_add(wg, gl2_scale(zeta, gl_mul(in.k_is[j], in.betas[i]))));
                den = gl2_mul(den, gl2_add(wg, gl2_scale(in.o_sig[j], in.betas[i])));
            }
            terms.push_back(gl2_sub(gl2_mul(accs[k], num), gl2_mul(accs[k + 1], den)));
        }
    }
    for (auto &t : gate_terms) terms.push_back(t.v);
    gl2 zh = gl2_sub(zeta_pow_n, one);
    for (u32 i = 0; i < nch; i++) {
        gl2 acc = gl2_make(0, 0);
        for (size_t t = terms.size(); t-- > 0;) acc = gl2_add(gl2_scale(acc, in.alphas[i]), terms[t]);
        gl2 q = gl2_make(0, 0);
        for (u32 k = qdf; k-- > 0;) q = gl2_add(gl2_mul(q, zeta_pow_n), in.o_quot[i * qdf + k]);
        if (!gl2_eq(acc, gl2_mul(zh, q))) return "vanishing polynomial identity failed";
    }
    return nullptr;
}
