/* libb200plonky2 -- C ABI of the B200-native plonky2 prover backend.
 *
 * Drop-in boundary: the Rust circuit code (CircuitBuilder, gadgets, witness generators)
 * stays untouched; a patched plonky2::plonk::prover::prove() hands the FILLED WIRE MATRIX plus handles to the
 * preprocessed circuit data across this ABI and gets plonky2's binary proof bytes back.
 *
 * Reference call sites each entry point replaces (paths relative to the reference tree):
 *   b200_circuit_upload  <- CircuitBuilder::build::<C>()   near_bft_finality/src/prove_crypto/ed25519.rs:33,75,
 *                           recursion.rs:94, sha256.rs:78, prove_block_data/keys_stakes.rs:243, primitives.rs:105,155,340
 *                           (the constants/sigmas PolynomialBatch + circuit digest part of build(); gate placement,
 *                           copy constraints and generator indexing stay in Rust)
 *   b200_prove           <- CircuitData::prove(pw)          ed25519.rs:60,100, recursion.rs:95, sha256.rs:82,
 *                           keys_stakes.rs:244, primitives.rs:108,158,343  (everything after generate_partial_witness)
 *   b200_proof wire fmt  <- ProofWithPublicInputs::to_bytes near_bft_finality/src/prove_block_data/signatures.rs:225,
 *                           src/bin/prove_block.rs:351
 *   b200_ntt_batch / b200_lde_commit / b200_poseidon_merkle : the prover's inner loops exposed for benchmarking
 *                           (upstream plonky2_field::fft, PolynomialBatch::from_values, MerkleTree::new; not in tree)
 *
 * Conventions: every uint64_t field element is little-endian canonical (< p = 2^64-2^32+1).  Matrices are
 * COLUMN-MAJOR: polynomial c occupies elements [c*n, (c+1)*n).  All pointers are caller-owned and borrowed for
 * the duration of the call unless stated.  Functions return 0 on success or a negative b200_status; the message
 * is available from b200_last_error().  Calls on one context are serialized by the caller (one context per
 * host thread / CUDA stream); distinct contexts may be used concurrently.  The library never aborts and has no
 * CPU fallback: without a CUDA device b200_ctx_create fails with B200_ERR_CUDA.
 */
#ifndef B200_PLONKY2_H
#define B200_PLONKY2_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    B200_OK = 0,
    B200_ERR_BAD_ARG = -1,
    B200_ERR_CUDA = -2,
    B200_ERR_UNSATISFIED = -3,   /* witness does not satisfy the circuit: the permutation grand product is not 1 (copy constraint
                                    violated) or the vanishing identity fails at zeta (a gate constraint is violated) */
    B200_ERR_OOM = -4,
    B200_ERR_UNSUPPORTED = -5,   /* e.g. unknown gate id, zero-knowledge blinding, lookups */
    B200_ERR_INVALID_PROOF = -6  /* b200_verify: the proof does not verify */
} b200_status;

typedef struct b200_ctx b200_ctx;
typedef struct b200_circuit b200_circuit;

/* Mirrors plonky2 CommonCircuitData (fields as serialized in common_data.json, see
 * near_bft_finality/proofs/random/CGZP.../common_data.json and gnark-plonky2-verifier/types/common_data.go:11-58). */
typedef struct {
    uint32_t num_wires;
    uint32_t num_routed_wires;
    uint32_t num_constants;          /* total constant columns incl. selector columns (common_data.num_constants) */
    uint32_t num_challenges;
    uint32_t degree_bits;
    uint32_t rate_bits;
    uint32_t cap_height;
    uint32_t proof_of_work_bits;
    uint32_t num_query_rounds;
    uint32_t quotient_degree_factor;
    uint32_t num_partial_products;
    uint32_t num_gate_constraints;
    uint32_t num_public_inputs;
    uint32_t hasher;                 /* 0 = PoseidonGoldilocksConfig; 1 = PoseidonBN128GoldilocksConfig (the final wrap) */
    uint32_t num_reduction_arity_bits;
    uint32_t reduction_arity_bits[16];
    uint32_t num_gates;
    const char *const *gate_ids;     /* plonky2 Gate::id() strings, in common_data.gates order */
    const uint32_t *selector_indices;/* [num_gates] */
    uint32_t num_selector_groups;
    const uint32_t *group_starts;    /* [num_selector_groups] */
    const uint32_t *group_ends;      /* [num_selector_groups] */
    const uint64_t *k_is;            /* [num_routed_wires] */
} b200_common;

/* ---- context ---- */
int b200_ctx_create(int device, b200_ctx **out);
void b200_ctx_destroy(b200_ctx *ctx);
/* Use an existing CUDA stream (cudaStream_t cast to void*) for all work of this context; NULL = own stream. */
int b200_ctx_set_stream(b200_ctx *ctx, void *cuda_stream);
int b200_ctx_synchronize(b200_ctx *ctx);
const char *b200_last_error(const b200_ctx *ctx);      /* valid until the next call on ctx */
void b200_free(void *p);                                /* frees buffers returned by the library */
/* Page-locked host memory for witness buffers handed to b200_prove / b200_lde_commit: a copy from pageable memory is
 * staged synchronously by the driver (~1 ms of host time for the 8.8 MB witness of a recursion proof), a copy from
 * these buffers overlaps the first kernels.  Replaces the Vec allocation of PartitionWitness::full_witness()
 * (plonky2 @ 2244a9d, iop/witness.rs) in the patched prove(); ordinary malloc'd buffers stay valid inputs. */
int b200_host_alloc(b200_ctx *ctx, size_t bytes, void **out);
void b200_host_free(void *p);
/* number of kernels this library launched on ctx since creation (bench.py's gpu_launches) */
uint64_t b200_kernel_launches(const b200_ctx *ctx);
/* Phase timing: when enabled, CUDA events are recorded on the context stream around the prover's phases
 * (lde, leaf_hash_wide = the dominant Poseidon leaf kernel, quotient, fri_combine, ...).  b200_ctx_timings
 * synchronizes and returns a JSON object {"phase": [total_ms, count], ...} accumulated since the last call. */
int b200_ctx_enable_timing(b200_ctx *ctx, int enable);
const char *b200_ctx_timings(b200_ctx *ctx);

/* ---- circuit (build-side) ----
 * constants_sigmas_values: [num_constants + num_routed_wires][n] column-major VALUES on the subgroup H (natural
 * order), i.e. plonky2's constant polys followed by the sigma polys.  Computes the constants/sigmas LDE + Merkle
 * cap on the GPU and keeps everything resident.  cap_out (optional): 2^cap_height * 4 u64.
 * circuit_digest: 4 u64 as computed by build(), or NULL to have the library derive it from the cap. */
int b200_circuit_upload(b200_ctx *ctx, const b200_common *common, const uint64_t *constants_sigmas_values,
                        const uint64_t *circuit_digest /* 4 u64, or NULL: derive it from the cap */, uint64_t *cap_out, b200_circuit **out);
void b200_circuit_free(b200_circuit *c);
/* CircuitBuilder::build::<C>()'s device half, from the builder's own data instead of finished polynomials
 * (near_bft_finality/src/prove_crypto/recursion.rs:94, ed25519.rs:33,75, sha256.rs:78: the constant / sigma polynomial
 * generation + commitment + digest that follow gate placement and copy-constraint resolution):
 *   row_gate        [n]                      index into common->gate_ids of the gate instantiated on each row
 *   gate_constants  [num_constants - num_selector_groups][n]  the gates' own constants, column-major (NULL if none)
 *   sigma_next      [num_routed_wires][n]    for routed slot s = col * n + row: the slot of the NEXT wire of its copy class
 *                                            (WirePartition::get_sigma_polys' cyclic successor; s itself if unconstrained)
 * The selector columns, the sigma polynomial values k_is[col'] * w^row', their LDE + Merkle cap and the circuit digest are
 * produced on the GPU / by the library.  Circuits are CACHED per context, keyed by a 128-bit hash of all inputs: the reference
 * rebuilds an identical recursion circuit on each of its k + 55 recursive_proof calls (recursion.rs:36-96); a repeated build
 * returns the resident circuit (*cache_hit = 1) after hashing its inputs.  Handles are reference counted: release each with
 * b200_circuit_free.  cap_out: 2^cap_height * 4 u64; digest_out: 4 u64 (both optional). */
int b200_circuit_build(b200_ctx *ctx, const b200_common *common, const uint32_t *row_gate, const uint64_t *gate_constants,
                       const uint32_t *sigma_next, uint64_t *cap_out, uint64_t *digest_out, int *cache_hit, b200_circuit **out);
/* circuit_digest = H(constants_sigmas_cap || H_pad([]) || degree_bits) with the config's hasher (host only). */
int b200_circuit_digest(uint32_t hasher, const uint64_t *cap, uint32_t cap_height, uint32_t degree_bits, uint64_t out[4]);
int b200_ctx_cache_stats(const b200_ctx *ctx, uint64_t *hits, uint64_t *misses, uint64_t *entries);

/* ---- witness generation, first slice: the SimpleGenerators of the u32 custom gates, batched on the device ----
 * For every row whose gate (common->gate_ids[row_gate[row]]) is U32ArithmeticGate / U32AddManyGate / U32SubtractionGate /
 * U32RangeCheckGate / ComparisonGate / U32InterleaveGate / UninterleaveToU32Gate / UninterleaveToB32Gate, the gate's DERIVED wires
 * (outputs, carries / borrows, inverse, 2-bit-limb and bit decompositions, chunk flags) are computed from its INPUT wires, in
 * place in the column-major wire matrix [num_wires][n]  (crypto/plonky2_u32/src/gates/*.rs `run_once`; one thread per (row, op)).
 * Rows of other gates are left untouched.  The dependency ordering of generators and the non-native / BigUint gadget generators
 * stay in Rust; the patched witness loop hands the rows of one dependency level to this call.  rows_filled (optional): rows touched. */
int b200_witness_fill_u32(b200_ctx *ctx, const b200_common *common, const uint32_t *row_gate, uint64_t *wires, uint32_t *rows_filled);
int b200_witness_fill_u32_dev(b200_ctx *ctx, const b200_common *common, const uint32_t *row_gate, uint64_t *d_wires, uint32_t *rows_filled);

/* ---- prove ----
 * wires: [num_wires][n] column-major (the full witness), public_inputs: [num_public_inputs].
 * pow_witness_hint: 0 = grind the minimum nonce (== reference with RAYON_NUM_THREADS=1); otherwise the value is
 * checked (16 leading zero bits) and used, so that a proof of the multi-threaded reference can be reproduced.
 * proof_bytes: malloc'd by the library (ProofWithPublicInputs::to_bytes layout), release with b200_free. */
int b200_prove(b200_ctx *ctx, const b200_circuit *circuit, const uint64_t *wires, const uint64_t *public_inputs,
               size_t n_public_inputs, uint64_t pow_witness_hint, uint8_t **proof_bytes, size_t *proof_len);
/* Same, wires already resident in device memory (device pointer). */
int b200_prove_dev(b200_ctx *ctx, const b200_circuit *circuit, const uint64_t *d_wires, const uint64_t *public_inputs,
                   size_t n_public_inputs, uint64_t pow_witness_hint, uint8_t **proof_bytes, size_t *proof_len);

/* ---- verify (host only, no device needed) ----
 * CircuitData::verify(proof) (near_bft_finality/src/prove_block_data/header_bphash.rs:94, primitives.rs:110,160).
 * Returns 0 if the proof verifies, B200_ERR_INVALID_PROOF otherwise (reason from b200_verify_error(), thread-local). */
typedef struct {
    const uint64_t *constants_sigmas_cap;   /* 2^cap_height hashes, 4 u64 each (32 LE bytes per hash for hasher 1) */
    uint64_t circuit_digest[4];
} b200_verifier_only;
int b200_verify(const b200_common *common, const b200_verifier_only *verifier_only, const uint8_t *proof, size_t len);
const char *b200_verify_error(void);

/* ---- inner loops (host-buffer entry points; *_dev variants take device pointers and run asynchronously on
 * the context stream) ---- */
/* Forward (inverse=0) or inverse NTT of `cols` polynomials of 2^log_n points, natural order in and out, in place. */
int b200_ntt_batch(b200_ctx *ctx, uint64_t *values, size_t cols, uint32_t log_n, int inverse);
int b200_ntt_batch_dev(b200_ctx *ctx, uint64_t *d_values, size_t cols, uint32_t log_n, int inverse, uint64_t *d_scratch);
/* PolynomialBatch::from_values: values [cols][n] -> coefficients (bit-reversed storage, internal), LDE
 * [cols][n << rate_bits] (row index bit-reversed, coset shift 7) and the Merkle cap (2^cap_height x 4 u64).
 * lde_out / coeffs_out may be NULL (host variant).  coeffs_out is in natural order. */
int b200_lde_commit(b200_ctx *ctx, const uint64_t *values, size_t cols, uint32_t log_n, uint32_t rate_bits,
                    uint32_t cap_height, uint64_t *cap_out, uint64_t *lde_out, uint64_t *coeffs_out);
int b200_lde_commit_dev(b200_ctx *ctx, const uint64_t *d_values, size_t cols, uint32_t log_n, uint32_t rate_bits,
                        uint32_t cap_height, uint64_t *d_coeffs_br, uint64_t *d_lde, uint64_t *d_digests);
/* Merkle cap over L = 2^log_L leaves given column-major data [cols][L]; digests_out (optional, host variant):
 * all levels, b200_merkle_num_u64(log_L, cap_height) u64, cap last. */
int b200_poseidon_merkle(b200_ctx *ctx, const uint64_t *data, size_t cols, uint32_t log_L, uint32_t cap_height,
                         uint64_t *cap_out, uint64_t *digests_out);
int b200_poseidon_merkle_dev(b200_ctx *ctx, const uint64_t *d_data, size_t cols, uint32_t log_L, uint32_t cap_height,
                             uint64_t *d_digests);
size_t b200_merkle_num_u64(uint32_t log_L, uint32_t cap_height);
/* One arity-2^arity_bits FRI fold of 2^log_len extension-field values (interleaved a,b; bit-reversed coset order)
 * by beta; shift = coset shift of the layer (7^(arity^layer)).  out: 2^(log_len-arity_bits) ext values. */
int b200_fri_fold(b200_ctx *ctx, const uint64_t *values, uint32_t log_len, uint32_t arity_bits, const uint64_t beta[2],
                  uint64_t shift, uint64_t *out);
/* ---- Groth16 wrap (next row, SURVEY 8f-3 / BASELINE config 4): BN254 G1 multi-scalar multiplication ----
 * (gnark-crypto ecc/bn254 MultiExp under groth16.Prove, gnark-plonky2-verifier/cmd/web-api.go:77)
 * points: n affine points, 8 u64 each (x limbs, y limbs; little-endian, canonical, not Montgomery; all-zero = infinity);
 * scalars: n x 4 u64 canonical (< r).  out_affine: 8 u64 (all-zero = infinity). */
int b200_msm_g1(b200_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, uint64_t out_affine[8]);
int b200_msm_g1_dev(b200_ctx *ctx, const uint64_t *d_points, const uint64_t *d_scalars, size_t n, uint64_t out_affine[8]);
/* Same on G2 (the twist y^2 = x^3 + 3/(9+u) over F_q2 = F_q[u]/(u^2+1)): the B-query MSM of groth16.Prove.
 * points: 16 u64 each = x.A0, x.A1, y.A0, y.A1 (gnark-crypto E2{A0, A1} order), 4 canonical limbs per coordinate;
 * out_affine: 16 u64 in the same order. */
int b200_msm_g2(b200_ctx *ctx, const uint64_t *points, const uint64_t *scalars, size_t n, uint64_t out_affine[16]);
int b200_msm_g2_dev(b200_ctx *ctx, const uint64_t *d_points, const uint64_t *d_scalars, size_t n, uint64_t out_affine[16]);
/* BN254 scalar-field NTT (the FFTs of groth16.Prove, gnark-crypto ecc/bn254/fr/fft: generator 5, 2-adicity 28):
 * values / out: 2^log_n elements x 4 u64 (canonical little-endian limbs), natural order in and out.
 * inverse = 0: A[k] = sum_i a[i] w^(ik)   (coset != 0: a[i] *= 5^i first, i.e. evaluation on the coset 5<w>)
 * inverse = 1: a[i] = n^-1 sum_k A[k] w^(-ik)   (coset != 0: a[i] *= 5^-i afterwards).  log_n <= 28. */
int b200_fr_ntt(b200_ctx *ctx, const uint64_t *values, uint32_t log_n, int inverse, int coset, uint64_t *out);
int b200_fr_ntt_dev(b200_ctx *ctx, const uint64_t *d_values, uint32_t log_n, int inverse, int coset, uint64_t *d_out);
/* groth16.Prove as a pipeline over the pieces above (gnark-plonky2-verifier/cmd/web-api.go:77 -> gnark v0.9.1
 * backend/groth16/bn254/prove.go, not vendored): H = (A.B - C)/Z by seven F_r transforms and a pointwise pass, the A / B1 / B2 /
 * K / Z multi-scalar multiplications, blinding by (r, s) and assembly.  The proving key lives on the device.
 *   wires    [num_wires][4]     the full R1CS solution, public wires (incl. the constant-one wire) first
 *   a, b, c  [2^log_domain][4]  <A_j, w>, <B_j, w>, <C_j, w> per constraint j, zero-padded to the domain (gnark's solver output)
 *   r, s     4 u64 each         blinding scalars (gnark draws them at random)
 *   proof    32 u64             Ar (G1, 8) | Bs (G2, 16) | Krs (G1, 8), affine canonical limbs
 * All scalars canonical < r (BN254 group order), 4 little-endian u64; points as for b200_msm_g1 / b200_msm_g2. */
typedef struct b200_groth16_pk b200_groth16_pk;
typedef struct {
    uint32_t log_domain;
    uint64_t num_wires, num_public;       /* num_public counts the constant-one wire */
    const uint64_t *g1_a, *g1_b;          /* [num_wires][8] */
    const uint64_t *g2_b;                 /* [num_wires][16] */
    const uint64_t *g1_k;                 /* [num_wires - num_public][8]: (beta A_i + alpha B_i + C_i)(tau) / delta for the private wires */
    const uint64_t *g1_z;                 /* [2^log_domain - 1][8]: tau^k Z(tau) / delta */
    uint64_t g1_alpha[8], g1_beta[8], g1_delta[8], g2_beta[16], g2_delta[16];
} b200_groth16_pk_desc;
int b200_groth16_pk_upload(b200_ctx *ctx, const b200_groth16_pk_desc *desc, b200_groth16_pk **out);
void b200_groth16_pk_free(b200_groth16_pk *pk);
int b200_groth16_prove(b200_ctx *ctx, const b200_groth16_pk *pk, const uint64_t *wires, const uint64_t *a, const uint64_t *b,
                       const uint64_t *c, const uint64_t r[4], const uint64_t s[4], uint64_t proof[32]);
/* Poseidon permutation of `count` independent 12-word states (self-test / throughput probe). */
int b200_poseidon_permute(b200_ctx *ctx, uint64_t *states, size_t count);

#ifdef __cplusplus
}
#endif
#endif
