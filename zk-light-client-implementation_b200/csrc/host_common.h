// Host-side pieces shared by the prover (prover.cu) and the verifier (verifier.cu): gate-id parsing and the
// Fiat-Shamir challenger.
#pragma once
#include "gates.cuh"
#include "poseidon.cuh"
#include "bn254.cuh"
#include "b200_internal.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

// ------------------------------------------------------------------------------------------------
// gate id parsing (plonky2 Gate::id() strings as they appear in common_data.json)
// ------------------------------------------------------------------------------------------------
static inline bool starts_with(const char *s, const char *p) { return strncmp(s, p, strlen(p)) == 0; }

static inline bool parse_gate(const char *id, GateDesc &g, std::vector<u64> &extra) {
    memset(&g, 0, sizeof g);
    unsigned a = 0, b = 0, c = 0;
    if (strcmp(id, "NoopGate") == 0) { g.kind = G_NOOP; g.num_constraints = 0; return true; }
    if (sscanf(id, "ConstantGate { num_consts: %u }", &a) == 1) { g.kind = G_CONSTANT; g.p[0] = a; g.num_constraints = a; return true; }
    if (strcmp(id, "PublicInputGate") == 0) { g.kind = G_PUBLIC_INPUT; g.num_constraints = 4; return true; }
    if (sscanf(id, "ArithmeticGate { num_ops: %u }", &a) == 1) { g.kind = G_ARITHMETIC; g.p[0] = a; g.num_constraints = a; return true; }
    if (sscanf(id, "ArithmeticExtensionGate { num_ops: %u }", &a) == 1) { g.kind = G_ARITHMETIC_EXT; g.p[0] = a; g.num_constraints = 2 * a; return true; }
    if (sscanf(id, "MulExtensionGate { num_ops: %u }", &a) == 1) { g.kind = G_MUL_EXT; g.p[0] = a; g.num_constraints = 2 * a; return true; }
    if (sscanf(id, "BaseSumGate { num_limbs: %u } + Base: %u", &a, &b) == 2) { g.kind = G_BASE_SUM; g.p[0] = a; g.p[1] = b; g.num_constraints = 1 + a; return true; }
    if (sscanf(id, "ReducingGate { num_coeffs: %u }", &a) == 1) { g.kind = G_REDUCING; g.p[0] = a; g.num_constraints = 2 * a; return true; }
    if (sscanf(id, "ReducingExtensionGate { num_coeffs: %u }", &a) == 1) { g.kind = G_REDUCING_EXT; g.p[0] = a; g.num_constraints = 2 * a; return true; }
    if (sscanf(id, "RandomAccessGate { bits: %u, num_copies: %u, num_extra_constants: %u", &a, &b, &c) == 3) {
        if (a > 6) return false;
        g.kind = G_RANDOM_ACCESS; g.p[0] = a; g.p[1] = b; g.p[2] = c; g.num_constraints = b * (a + 2) + c; return true; }
    if (starts_with(id, "PoseidonMdsGate")) { g.kind = G_POSEIDON_MDS; g.num_constraints = 24; return true; }
    if (starts_with(id, "PoseidonGate")) { g.kind = G_POSEIDON; g.num_constraints = 123; return true; }
    if (sscanf(id, "CosetInterpolationGate { subgroup_bits: %u, degree: %u, barycentric_weights: [", &a, &b) == 2) {
        const char *p = strchr(id, '['); if (!p) return false; p++;
        g.kind = G_COSET_INTERPOLATION; g.p[0] = a; g.p[1] = b; g.extra_off = (u32)extra.size();
        unsigned np = 1u << a;
        for (unsigned i = 0; i < np; i++) {
            char *end; unsigned long long v = strtoull(p, &end, 10); if (end == p) return false;
            extra.push_back((u64)v); p = end; while (*p == ',' || *p == ' ') p++;
        }
        if (b < 2 || a > 6) return false;
        unsigned n_inter = (np - 2) / (b - 1);
        g.num_constraints = 2 + 4 * n_inter + 2; return true; }
    if (sscanf(id, "U32ArithmeticGate { num_ops: %u", &a) == 1) { g.kind = G_U32_ARITHMETIC; g.p[0] = a; g.num_constraints = 36 * a; return true; }
    if (sscanf(id, "U32AddManyGate { num_addends: %u, num_ops: %u", &a, &b) == 2) { g.kind = G_U32_ADD_MANY; g.p[0] = a; g.p[1] = b; g.num_constraints = 21 * b; return true; }
    if (sscanf(id, "U32SubtractionGate { num_ops: %u", &a) == 1) { g.kind = G_U32_SUBTRACTION; g.p[0] = a; g.num_constraints = 19 * a; return true; }
    if (sscanf(id, "U32RangeCheckGate { num_input_limbs: %u", &a) == 1) { g.kind = G_U32_RANGE_CHECK; g.p[0] = a; g.num_constraints = 17 * a; return true; }
    if (sscanf(id, "U32InterleaveGate { num_ops: %u", &a) == 1) { g.kind = G_U32_INTERLEAVE; g.p[0] = a; g.num_constraints = 34 * a; return true; }
    if (sscanf(id, "UninterleaveToU32Gate { num_ops: %u", &a) == 1) { g.kind = G_UNINTERLEAVE_U32; g.p[0] = a; g.num_constraints = 67 * a; return true; }
    if (sscanf(id, "UninterleaveToB32Gate { num_ops: %u", &a) == 1) { g.kind = G_UNINTERLEAVE_B32; g.p[0] = a; g.num_constraints = 67 * a; return true; }
    if (sscanf(id, "ComparisonGate { num_bits: %u, num_chunks: %u", &a, &b) == 2) {
        if (b == 0) return false;
        unsigned cb = (a + b - 1) / b; if (cb > 6) return false;
        g.kind = G_COMPARISON; g.p[0] = a; g.p[1] = b; g.num_constraints = 6 + 5 * b + cb; return true; }
    return false;
}

// ------------------------------------------------------------------------------------------------
// Fiat-Shamir challenger (host).  Semantics: gnark-plonky2-verifier/challenger/challenger.go:42-165.
// ------------------------------------------------------------------------------------------------
struct Challenger {
    u64 st[12]; u64 in[8]; int nin; u64 out[8]; int nout;
    Challenger() { memset(this, 0, sizeof *this); }
    void duplex() {
        for (int i = 0; i < nin; i++) st[i] = in[i];
        nin = 0; poseidon_permute(st);
        for (int i = 0; i < 8; i++) out[i] = st[i];
        nout = 8;
    }
    void observe(u64 x) { nout = 0; in[nin++] = x; if (nin == 8) duplex(); }
    void observe_n(const u64 *x, size_t n) { for (size_t i = 0; i < n; i++) observe(x[i]); }
    void observe_ext(gl2 e) { observe(e.a); observe(e.b); }
    // a Merkle/sponge hash of C::Hasher: 4 elements (Poseidon-GL) or 5 x 56-bit limbs (Poseidon-BN128)
    void observe_hash(const u64 h[4], int hasher) {
        if (hasher == HASHER_BN254) { u64 v[5]; bn254_to_vec_host(h, v); observe_n(v, 5); }
        else observe_n(h, 4);
    }
    void observe_cap(const u64 *cap, size_t n_hashes, int hasher) { for (size_t i = 0; i < n_hashes; i++) observe_hash(cap + 4 * i, hasher); }
    u64 get() { if (nin || !nout) duplex(); return out[--nout]; }
    gl2 get_ext() { gl2 r; r.a = get(); r.b = get(); return r; }
};


// ------------------------------------------------------------------------------------------------
// C::Hasher on the host (both hashers): hash_or_noop / two_to_one, and the circuit digest of CircuitBuilder::build
//   circuit_digest = H(constants_sigmas_cap || H_pad(domain_separator = []) || degree_bits)
// (upstream plonky2 plonk/circuit_builder.rs, not vendored; pinned by the golden verifier_data.json of the three final
// proofs for the BN128 hasher: tests/test_oracle_verifier.py, tests/test_product_verifier.py)
// ------------------------------------------------------------------------------------------------
static inline void hash_or_noop(const u64 *x, size_t n, int hasher, u64 out[4]) {
    if (hasher == HASHER_BN254) {
        if (n <= 3) { for (int i = 0; i < 4; i++) out[i] = (size_t)i < n ? x[i] : 0; return; }
        fr_t s[4]; for (int i = 0; i < 4; i++) for (int k = 0; k < 4; k++) s[i].v[k] = 0;
        for (size_t i = 0; i < n; i += 9) {
            size_t m = n - i < 9 ? n - i : 9;
            for (size_t j = 0; j < m; j += 3) { u64 l[4] = {x[i + j], j + 1 < m ? x[i + j + 1] : 0, j + 2 < m ? x[i + j + 2] : 0, 0}; s[j / 3 + 1] = fr_from_canonical(l); }
            bn254_permute(s);
        }
        fr_to_canonical(s[0], out);
    } else {
        if (n <= 4) { for (int i = 0; i < 4; i++) out[i] = (size_t)i < n ? x[i] : 0; return; }
        poseidon_hash_no_pad_host(x, n, out);
    }
}
static inline void two_to_one(const u64 l[4], const u64 r[4], int hasher, u64 out[4]) {
    if (hasher == HASHER_BN254) bn254_two_to_one(l, r, out); else poseidon_two_to_one(l, r, out);
}

static inline void circuit_digest_host(const u64 *cap, size_t ncap, u32 degree_bits, int hasher, u64 out[4]) {
    std::vector<u64> parts;
    if (hasher == HASHER_BN254) {
        u64 pad[9] = {1, 0, 0, 0, 0, 0, 0, 0, 1}, dom[4], v[5];      // hash_pad([]): [1], zero-fill to 8, closing 1 (rate 9)
        hash_or_noop(pad, 9, hasher, dom);
        for (size_t i = 0; i < ncap; i++) { bn254_to_vec_host(cap + 4 * i, v); parts.insert(parts.end(), v, v + 5); }
        bn254_to_vec_host(dom, v); parts.insert(parts.end(), v, v + 5);
    } else {
        u64 pad[8] = {1, 0, 0, 0, 0, 0, 0, 1}, dom[4];               // rate 8
        hash_or_noop(pad, 8, hasher, dom);
        parts.assign(cap, cap + 4 * ncap);
        parts.insert(parts.end(), dom, dom + 4);
    }
    parts.push_back(degree_bits);
    hash_or_noop(parts.data(), parts.size(), hasher, out);
}
