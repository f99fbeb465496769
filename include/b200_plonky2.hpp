// C++ host-side mirror of the reference's plonky2 surface for the hot path, header-only over the C ABI
// (include/b200_plonky2.h).  Names follow the reference's Rust API so call sites read alike:
//   CircuitData::prove(pw)            near_bft_finality/src/prove_crypto/ed25519.rs:60, recursion.rs:95, sha256.rs:82
//   CircuitBuilder::build::<C>()      ed25519.rs:33, recursion.rs:94 (here: CircuitData::from_build_outputs — the part of
//                                     build() that the GPU takes over: constants/sigmas commitment)
//   ProofWithPublicInputs::to_bytes   prove_block_data/signatures.rs:225
// Error behaviour mirrors anyhow::Result: failures throw b200::Error carrying the status code and b200_last_error().
#pragma once
#include "b200_plonky2.h"
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error("b200 error " + std::to_string(c) + ": " + m), code(c) {}
};

class Context {                       // one per GPU / stream / host thread
public:
    explicit Context(int device = 0) { int rc = b200_ctx_create(device, &ctx_); if (rc) throw Error(rc, "b200_ctx_create failed (no CUDA device: there is no CPU fallback)"); }
    ~Context() { b200_ctx_destroy(ctx_); }
    Context(const Context &) = delete; Context &operator=(const Context &) = delete;
    b200_ctx *raw() const { return ctx_; }
    void check(int rc) const { if (rc) throw Error(rc, b200_last_error(ctx_)); }
    // page-locked witness buffer (b200_host_alloc); release with b200_host_free
    uint64_t *alloc_witness(size_t num_wires, size_t degree) const { void *p = nullptr; check(b200_host_alloc(ctx_, num_wires * degree * 8, &p)); return (uint64_t *)p; }
private:
    b200_ctx *ctx_ = nullptr;
};

// plonky2 CommonCircuitData (owning the arrays the C struct points to)
struct CommonCircuitData {
    b200_common c{};
    std::vector<std::string> gate_ids; std::vector<const char *> gate_ptrs;
    std::vector<uint32_t> selector_indices, group_starts, group_ends; std::vector<uint64_t> k_is;
    void finalize() {
        gate_ptrs.clear(); for (auto &g : gate_ids) gate_ptrs.push_back(g.c_str());
        c.num_gates = (uint32_t)gate_ids.size(); c.gate_ids = gate_ptrs.data(); c.selector_indices = selector_indices.data();
        c.num_selector_groups = (uint32_t)group_starts.size(); c.group_starts = group_starts.data(); c.group_ends = group_ends.data();
        c.k_is = k_is.data();
    }
    size_t degree() const { return size_t(1) << c.degree_bits; }
};

struct ProofWithPublicInputs {
    std::vector<uint8_t> bytes;                                     // ProofWithPublicInputs::to_bytes() layout
    const std::vector<uint8_t> &to_bytes() const { return bytes; }
};

class CircuitData {
public:
    // constants_sigmas_values: [num_constants + num_routed_wires][n] column-major values on H; circuit_digest: 4 words
    CircuitData(Context &ctx, CommonCircuitData common, const uint64_t *constants_sigmas_values, const uint64_t circuit_digest[4])
        : ctx_(ctx), common_(std::move(common)) {
        common_.finalize();
        cap_.resize(size_t(4) << common_.c.cap_height);
        ctx_.check(b200_circuit_upload(ctx_.raw(), &common_.c, constants_sigmas_values, circuit_digest, cap_.data(), &circuit_));
    }
    ~CircuitData() { b200_circuit_free(circuit_); }
    CircuitData(const CircuitData &) = delete; CircuitData &operator=(const CircuitData &) = delete;
    const std::vector<uint64_t> &constants_sigmas_cap() const { return cap_; }
    const CommonCircuitData &common() const { return common_; }
    // wires: the full witness [num_wires][n] column-major (what PartitionWitness::full_witness() yields)
    ProofWithPublicInputs prove(const uint64_t *wires, const std::vector<uint64_t> &public_inputs, uint64_t pow_witness_hint = 0) const {
        uint8_t *p = nullptr; size_t len = 0;
        ctx_.check(b200_prove(ctx_.raw(), circuit_, wires, public_inputs.data(), public_inputs.size(), pow_witness_hint, &p, &len));
        ProofWithPublicInputs out; out.bytes.assign(p, p + len);
        b200_free(p);
        return out;
    }
    // CircuitData::verify(proof) (header_bphash.rs:94, primitives.rs:110,160): host-only check, throws on an invalid proof
    void verify(const ProofWithPublicInputs &proof, const uint64_t circuit_digest[4]) const {
        b200_verifier_only vo; vo.constants_sigmas_cap = cap_.data();
        for (int i = 0; i < 4; i++) vo.circuit_digest[i] = circuit_digest[i];
        int rc = b200_verify(&common_.c, &vo, proof.bytes.data(), proof.bytes.size());
        if (rc) throw Error(rc, b200_verify_error());
    }
private:
    Context &ctx_; CommonCircuitData common_; b200_circuit *circuit_ = nullptr; std::vector<uint64_t> cap_;
};

}  // namespace b200
