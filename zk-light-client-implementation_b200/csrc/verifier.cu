// b200_verify: plonky2 proof verification on the host (CircuitData::verify at near_bft_finality/src/prove_block_data/
// header_bphash.rs:94, primitives.rs:110,160).  Pure host code (no device needed).  Mirrors upstream plonky2's verifier
// as restated by gnark-plonky2-verifier/verifier/verifier.go:45-170 (challenges), plonk/plonk.go:57-250 (vanishing
// identity at zeta), fri/fri.go:63-548 (PoW, Merkle openings, batching, folding, final polynomial).  Both hashers:
// PoseidonGoldilocksConfig and PoseidonBN128GoldilocksConfig -- the latter is checked against the reference's golden
// wrapper proofs in tests/test_product_verifier.py.
#include "../../include/b200_plonky2.h"
#include "host_common.h"
#include "verify_gates.h"
#include "vanishing.h"
#include <string>

static thread_local std::string g_verify_err;
struct VerifyFail { std::string msg; };
#define VFAIL(m) throw VerifyFail{m}

struct Reader {
    const uint8_t *p; size_t len, off;
    u64 u64v() { if (off + 8 > len) VFAIL("proof truncated"); u64 v = 0; for (int i = 0; i < 8; i++) v |= (u64)p[off + i] << (8 * i); off += 8; return v; }
    u64 fe() { u64 v = u64v(); if (v >= GL_P) VFAIL("non-canonical field element in proof"); return v; }
    gl2 ext() { gl2 r; r.a = fe(); r.b = fe(); return r; }
    uint8_t u8v() { if (off + 1 > len) VFAIL("proof truncated"); return p[off++]; }
    void hash(u64 h[4]) { for (int i = 0; i < 4; i++) h[i] = u64v(); }
};
struct H4 { u64 w[4]; };
struct QueryStep { std::vector<gl2> evals; std::vector<H4> sib; };
struct QueryRound { std::vector<std::vector<u64>> leaf; std::vector<std::vector<H4>> sib; std::vector<QueryStep> steps; };

static void verify_merkle(const u64 *leaf, size_t n, size_t index, const std::vector<H4> &sib, const u64 *cap, size_t cap_len, int hasher) {
    u64 d[4]; hash_or_noop(leaf, n, hasher, d);
    for (auto &s : sib) { u64 o[4]; if (index & 1) two_to_one(s.w, d, hasher, o); else two_to_one(d, s.w, hasher, o); memcpy(d, o, 32); index >>= 1; }
    if (index >= cap_len || memcmp(cap + 4 * index, d, 32) != 0) VFAIL("merkle proof does not match cap");
}

static gl2 ext_pow(gl2 a, u64 e) { return gl2_pow(a, e); }

static int verify_impl(const b200_common *c, const b200_verifier_only *vo, const uint8_t *proof, size_t len) {
    if (!c || !vo || !proof || !vo->constants_sigmas_cap) VFAIL("null argument");
    if (c->hasher > 1 || c->num_challenges == 0 || c->num_challenges > 4) VFAIL("unsupported config");
    if (c->degree_bits == 0 || c->degree_bits > 32 || c->rate_bits == 0 || c->rate_bits > 8 || c->cap_height > c->degree_bits + c->rate_bits ||
        c->proof_of_work_bits == 0 || c->proof_of_work_bits > 32 || c->num_reduction_arity_bits > 16 || c->num_query_rounds > 4096 ||
        c->quotient_degree_factor == 0 || c->num_routed_wires > c->num_wires || c->num_gates > 4096)
        VFAIL("common data out of range");
    for (u32 i = 0; i < c->num_reduction_arity_bits; i++) if (c->reduction_arity_bits[i] == 0 || c->reduction_arity_bits[i] > 8) VFAIL("reduction arity bits out of range");
    const int HS = (int)c->hasher;
    const u32 W = c->num_wires, R = c->num_routed_wires, NC = c->num_constants, nch = c->num_challenges, qdf = c->quotient_degree_factor, npp = c->num_partial_products;
    const int nlog = (int)c->degree_bits, rb = (int)c->rate_bits, chh = (int)c->cap_height, Llog = nlog + rb;
    const size_t ncap = size_t(1) << chh;
    const u32 widths[4] = {NC + R, W, nch * (1 + npp), nch * qdf};
    std::vector<GateDesc> gates; std::vector<u64> extra;
    for (u32 i = 0; i < c->num_gates; i++) {
        GateDesc g; if (!parse_gate(c->gate_ids[i], g, extra)) VFAIL(std::string("unsupported gate id: ") + c->gate_ids[i]);
        u32 si = c->selector_indices[i]; if (si >= c->num_selector_groups) VFAIL("selector index out of range");
        g.selector_index = si; g.group_start = c->group_starts[si]; g.group_end = c->group_ends[si]; g.row = i; gates.push_back(g);
    }
    // ---- parse ----
    Reader rd{proof, len, 0};
    std::vector<u64> caps[3]; for (auto &cp : caps) { cp.resize(4 * ncap); for (size_t i = 0; i < ncap; i++) rd.hash(&cp[4 * i]); }
    auto exts = [&](size_t n) { std::vector<gl2> v(n); for (auto &e : v) e = rd.ext(); return v; };
    auto o_const = exts(NC), o_sig = exts(R), o_wires = exts(W), o_zs = exts(nch), o_zs_next = exts(nch), o_pp = exts((size_t)nch * npp), o_quot = exts((size_t)nch * qdf);
    std::vector<std::vector<u64>> fri_caps(c->num_reduction_arity_bits);
    for (auto &cp : fri_caps) { cp.resize(4 * ncap); for (size_t i = 0; i < ncap; i++) rd.hash(&cp[4 * i]); }
    std::vector<QueryRound> rounds(c->num_query_rounds);
    for (auto &q : rounds) {
        q.leaf.resize(4); q.sib.resize(4);
        for (int o = 0; o < 4; o++) {
            q.leaf[o].resize(widths[o]); for (auto &x : q.leaf[o]) x = rd.fe();
            size_t ns = rd.u8v(); if ((int)ns + chh != Llog) VFAIL("bad initial merkle proof length");
            q.sib[o].resize(ns); for (auto &h : q.sib[o]) rd.hash(h.w);
        }
        int bits = Llog;
        q.steps.resize(c->num_reduction_arity_bits);
        for (u32 li = 0; li < c->num_reduction_arity_bits; li++) {
            u32 ab = c->reduction_arity_bits[li]; bits -= (int)ab;
            q.steps[li].evals = exts(size_t(1) << ab);
            size_t ns = rd.u8v(); if ((int)ns + chh != bits) VFAIL("bad step merkle proof length");
            q.steps[li].sib.resize(ns); for (auto &h : q.steps[li].sib) rd.hash(h.w);
        }
    }
    u32 total_ab = 0; for (u32 li = 0; li < c->num_reduction_arity_bits; li++) total_ab += c->reduction_arity_bits[li];
    if ((int)total_ab > nlog) VFAIL("bad reduction arity bits");
    auto final_poly = exts(size_t(1) << (nlog - total_ab));
    u64 pow_witness = rd.fe();
    u64 npi = rd.u64v(); if (npi != c->num_public_inputs) VFAIL("wrong number of public inputs");
    std::vector<u64> pis(npi); for (auto &x : pis) x = rd.fe();
    if (rd.off != len) VFAIL("trailing bytes after proof");
    // ---- challenges (verifier.go:45-82) ----
    u64 pi_hash[4]; poseidon_hash_no_pad_host(pis.data(), pis.size(), pi_hash);
    Challenger ch;
    ch.observe_hash(vo->circuit_digest, HS); ch.observe_n(pi_hash, 4); ch.observe_cap(caps[0].data(), ncap, HS);
    std::vector<u64> betas(nch), gammas(nch), alphas(nch);
    for (auto &b : betas) b = ch.get(); for (auto &g : gammas) g = ch.get();
    ch.observe_cap(caps[1].data(), ncap, HS);
    for (auto &a : alphas) a = ch.get();
    ch.observe_cap(caps[2].data(), ncap, HS);
    gl2 zeta = ch.get_ext();
    for (auto *v : {&o_const, &o_sig, &o_wires, &o_zs, &o_pp, &o_quot}) for (auto &e : *v) ch.observe_ext(e);
    for (auto &e : o_zs_next) ch.observe_ext(e);
    gl2 fri_alpha = ch.get_ext();
    std::vector<gl2> fri_betas;
    for (auto &cp : fri_caps) { ch.observe_cap(cp.data(), ncap, HS); fri_betas.push_back(ch.get_ext()); }
    for (auto &e : final_poly) ch.observe_ext(e);
    ch.observe(pow_witness);
    u64 pow_response = ch.get();
    std::vector<u64> qidx(c->num_query_rounds); for (auto &q : qidx) q = ch.get();
    // ---- vanishing polynomial identity at zeta (plonk.go:121-250) ----
    {
        VanishingInput vi{c->num_wires, c->num_routed_wires, c->num_constants, nch, qdf, npp, c->num_gate_constraints, c->num_selector_groups, nlog,
                          c->k_is, &gates, extra.data(), pi_hash, betas.data(), gammas.data(), alphas.data(), zeta,
                          o_const.data(), o_sig.data(), o_wires.data(), o_zs.data(), o_zs_next.data(), o_pp.data(), o_quot.data()};
        const char *why = vanishing_identity_at_zeta(vi);
        if (why) VFAIL(why);
    }
    // ---- FRI (fri.go:63-548) ----
    if (pow_response >> (64 - c->proof_of_work_bits)) VFAIL("proof of work failed");
    std::vector<gl2> batch0; for (auto *v : {&o_const, &o_sig, &o_wires, &o_zs, &o_pp, &o_quot}) batch0.insert(batch0.end(), v->begin(), v->end());
    auto reduce = [&](const std::vector<gl2> &v) { gl2 acc = gl2_make(0, 0); for (size_t i = v.size(); i-- > 0;) acc = gl2_add(gl2_mul(acc, fri_alpha), v[i]); return acc; };
    gl2 y0 = reduce(batch0), y1 = reduce(o_zs_next);
    gl2 gzeta = gl2_scale(zeta, gl_root_of_unity(nlog));
    const u64 *init_caps[4] = {vo->constants_sigmas_cap, caps[0].data(), caps[1].data(), caps[2].data()};
    u64 w_lde = gl_root_of_unity(Llog);
    for (u32 qi = 0; qi < c->num_query_rounds; qi++) {
        const QueryRound &q = rounds[qi];
        size_t x_index = (size_t)(qidx[qi] & ((u64(1) << Llog) - 1));
        for (int o = 0; o < 4; o++) verify_merkle(q.leaf[o].data(), q.leaf[o].size(), x_index, q.sib[o], init_caps[o], ncap, HS);
        u64 subgroup_x = gl_mul(GL_GEN, gl_pow(w_lde, bitrev32((u32)x_index, Llog)));
        // fri_combine_initial
        gl2 f0 = gl2_make(0, 0), f1 = gl2_make(0, 0);
        {
            std::vector<u64> all; for (int o = 0; o < 4; o++) all.insert(all.end(), q.leaf[o].begin(), q.leaf[o].end());
            for (size_t i = all.size(); i-- > 0;) f0 = gl2_add_base(gl2_mul(f0, fri_alpha), all[i]);
            for (u32 i = nch; i-- > 0;) f1 = gl2_add_base(gl2_mul(f1, fri_alpha), q.leaf[2][i]);
        }
        gl2 xe = gl2_make(subgroup_x, 0);
        gl2 sum = gl2_mul(gl2_sub(f0, y0), gl2_inv(gl2_sub(xe, zeta)));
        sum = gl2_add(gl2_mul(sum, ext_pow(fri_alpha, nch)), gl2_mul(gl2_sub(f1, y1), gl2_inv(gl2_sub(xe, gzeta))));
        gl2 old_eval = sum;
        size_t idx = x_index;
        for (u32 li = 0; li < c->num_reduction_arity_bits; li++) {
            u32 ab = c->reduction_arity_bits[li]; size_t arity = size_t(1) << ab;
            const QueryStep &st = q.steps[li];
            size_t within = idx & (arity - 1), coset = idx >> ab;
            if (!gl2_eq(st.evals[within], old_eval)) VFAIL("fri fold consistency failed");
            // compute_evaluation: interpolate the coset values and evaluate at beta
            u64 g = gl_root_of_unity((int)ab), ginv = gl_inv(g);
            u64 start = gl_mul(subgroup_x, gl_pow(ginv, bitrev32((u32)within, (int)ab)));
            std::vector<u64> xs(arity); std::vector<gl2> ys(arity);
            for (size_t i = 0; i < arity; i++) { xs[i] = gl_mul(start, gl_pow(g, i)); ys[bitrev32((u32)i, (int)ab)] = st.evals[i]; }
            gl2 beta = fri_betas[li], lx = gl2_make(1, 0), s = gl2_make(0, 0); bool hit = false; gl2 hitv = gl2_make(0, 0);
            for (size_t i = 0; i < arity; i++) {
                gl2 d = gl2_sub(beta, gl2_make(xs[i], 0));
                if (d.a == 0 && d.b == 0) { hit = true; hitv = ys[i]; break; }
                lx = gl2_mul(lx, d);
                u64 wgt = 1; for (size_t j = 0; j < arity; j++) if (j != i) wgt = gl_mul(wgt, gl_sub(xs[i], xs[j]));
                s = gl2_add(s, gl2_mul(ys[i], gl2_scale(gl2_inv(d), gl_inv(wgt))));
            }
            old_eval = hit ? hitv : gl2_mul(lx, s);
            std::vector<u64> flat; for (auto &e : st.evals) { flat.push_back(e.a); flat.push_back(e.b); }
            verify_merkle(flat.data(), flat.size(), coset, st.sib, fri_caps[li].data(), ncap, HS);
            subgroup_x = gl_pow(subgroup_x, (u64)arity);
            idx = coset;
        }
        gl2 fe = gl2_make(0, 0), xq = gl2_make(subgroup_x, 0);
        for (size_t i = final_poly.size(); i-- > 0;) fe = gl2_add(gl2_mul(fe, xq), final_poly[i]);
        if (!gl2_eq(fe, old_eval)) VFAIL("final polynomial evaluation mismatch");
    }
    return B200_OK;
}

extern "C" {
int b200_verify(const b200_common *common, const b200_verifier_only *verifier_only, const uint8_t *proof, size_t len) {
    try { g_verify_err.clear(); return verify_impl(common, verifier_only, proof, len); }
    catch (const VerifyFail &e) { g_verify_err = e.msg; return B200_ERR_INVALID_PROOF; }
    catch (const std::exception &e) { g_verify_err = e.what(); return B200_ERR_BAD_ARG; }
}
const char *b200_verify_error(void) { return g_verify_err.c_str(); }
}
