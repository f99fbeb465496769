"""Synthetic circuits shaped like the reference's proofs (gate types, wire widths, row counts), built with the
mini CircuitBuilder.  'sig' mirrors the ed25519 signature circuit (wide_ecc_config, u32 gate mix:
crypto/plonky2_ed25519 + plonky2_sha512 + plonky2_u32 gates); 'rec' mirrors a recursive_proof circuit
(standard_recursion_config, Poseidon-heavy: near_bft_finality/src/prove_crypto/recursion.rs)."""
import numpy as np
from . import circuit_builder as cb

def _mix(total, weights):
    w = np.array(weights, dtype=float); w /= w.sum()
    cnt = np.floor(w * total).astype(int)
    return [max(1, c) for c in cnt]

def sig_circuit(degree_bits, seed=0, num_public_inputs=8):
    """ed25519-sub-proof-shaped circuit: 234 wires, 2^degree_bits rows."""
    cfg = cb.CircuitConfig.wide_ecc_config()
    rng = np.random.default_rng(seed)
    b = cb.CircuitBuilder(cfg)
    pis = [int(x) for x in rng.integers(0, 1 << 32, size=num_public_inputs)]
    b.register_public_inputs(pis)
    budget = (1 << degree_bits) - b.num_rows - 8
    assert budget > 40, "degree_bits too small for the sig-shaped mix"
    # gate histogram: a guess at the real circuit's (SURVEY App. B.1 needs a Rust box): u32 arithmetic dominates (non-native
    # field products), SHA-512 brings interleave/uninterleave/add-many, the windowed scalar multiplication brings
    # RandomAccessGate rows over 16-entry tables (curve_windowed_mul.rs:110-149), comparisons/range checks the reductions
    c = _mix(budget, [40, 6, 6, 8, 5, 8, 6, 6, 2, 8, 3, 2, 6, 1])
    cb.fill_u32_arithmetic(b, rng, c[0])
    cb.fill_u32_add_many(b, rng, c[1], 2)
    cb.fill_u32_add_many(b, rng, c[2], 5)
    cb.fill_u32_subtraction(b, rng, c[3])
    cb.fill_u32_range_check(b, rng, c[4], 8)
    cb.fill_u32_interleave(b, rng, c[5])
    cb.fill_uninterleave(b, rng, c[6], False)
    cb.fill_uninterleave(b, rng, c[7], True)
    cb.fill_comparison(b, rng, c[8])
    r1 = cb.fill_arithmetic(b, rng, c[9])
    cb.fill_base_sum(b, rng, c[10])
    cs = rng.integers(0, 1 << 62, size=(c[11], 2), dtype=np.uint64)
    rc = cb.fill_constant(b, cs)
    cb.fill_random_access(b, rng, c[12])
    cb.fill_poseidon_chain(b, rng, c[13])
    # cross-row copy constraints: pairs of arithmetic rows share their first multiplicand via a constant row
    k = min(len(rc), len(r1))
    return b, (rc, r1, k)

def rec_circuit(degree_bits, seed=0, num_public_inputs=4):
    """recursive_proof-shaped circuit: 135 wires, Poseidon-dominated."""
    cfg = cb.CircuitConfig.standard_recursion_config()
    rng = np.random.default_rng(seed)
    b = cb.CircuitBuilder(cfg)
    pis = [int(x) for x in rng.integers(0, 1 << 32, size=num_public_inputs)]
    b.register_public_inputs(pis)
    budget = (1 << degree_bits) - b.num_rows - 4
    assert budget > 24
    # gate mix of a recursive verifier circuit (all 13 gate types of common_data.json; Poseidon rows dominate)
    c = _mix(budget, [50, 12, 5, 5, 4, 4, 4, 4, 4, 4, 4])
    cb.fill_poseidon_chain(b, rng, c[0])
    cb.fill_arithmetic(b, rng, c[1])
    cb.fill_base_sum(b, rng, c[2])
    cb.fill_constant(b, rng.integers(0, 1 << 62, size=(c[3], 2), dtype=np.uint64))
    cb.fill_arithmetic_extension(b, rng, c[4])
    cb.fill_mul_extension(b, rng, c[5])
    cb.fill_reducing(b, rng, c[6], False)
    cb.fill_reducing(b, rng, c[7], True)
    cb.fill_random_access(b, rng, c[8])
    cb.fill_poseidon_mds(b, rng, c[9])
    cb.fill_coset_interpolation(b, rng, c[10])
    return b, None

def sha_circuit(degree_bits, seed=0, num_public_inputs=8):
    """sha256_proof_u32-shaped circuit (near_bft_finality/src/prove_crypto/sha256.rs:62-84, crypto/plonky2_sha256_u32):
    standard_recursion_config (135 wires) with the u32 interleave / uninterleave / add-many / arithmetic gates the
    interleave-trick SHA-256 uses, digest words as public inputs."""
    cfg = cb.CircuitConfig.standard_recursion_config()
    rng = np.random.default_rng(seed)
    b = cb.CircuitBuilder(cfg)
    b.register_public_inputs([int(x) for x in rng.integers(0, 1 << 32, size=num_public_inputs)])
    budget = (1 << degree_bits) - b.num_rows - 4
    assert budget > 20
    c = _mix(budget, [25, 20, 20, 15, 8, 6, 4, 2])
    cb.fill_u32_interleave(b, rng, c[0])
    cb.fill_uninterleave(b, rng, c[1], False)
    cb.fill_uninterleave(b, rng, c[2], True)
    cb.fill_u32_add_many(b, rng, c[3], 4)
    cb.fill_u32_arithmetic(b, rng, c[4])
    cb.fill_u32_range_check(b, rng, c[5], 4)
    cb.fill_arithmetic(b, rng, c[6])
    cb.fill_constant(b, rng.integers(0, 1 << 62, size=(c[7], 2), dtype=np.uint64))
    return b, None

def tiny_circuit(degree_bits, seed=0, num_public_inputs=0):
    """edge-case circuit: 2^degree_bits rows (>= 4) of arithmetic / constant gates, optional public inputs; small enough
    that FRI has no reduction layer (degree_bits <= 5)."""
    cfg = cb.CircuitConfig.standard_recursion_config()
    rng = np.random.default_rng(seed)
    b = cb.CircuitBuilder(cfg)
    if num_public_inputs:
        b.register_public_inputs([int(x) for x in rng.integers(0, 1 << 32, size=num_public_inputs)])
    budget = (1 << degree_bits) - b.num_rows
    a = max(1, budget // 2)
    cb.fill_arithmetic(b, rng, a)
    if budget - a > 0:
        cb.fill_constant(b, rng.integers(0, 1 << 62, size=(budget - a, 2), dtype=np.uint64))
    return b, None

def build(kind, degree_bits, seed=0, hasher="gl"):
    """kind: "sig" | "rec"; hasher "bn254" = PoseidonBN128GoldilocksConfig (the final wrap proof, prove_block.rs:279-300)."""
    if kind == "tiny": b, _ = tiny_circuit(degree_bits, seed)
    elif kind == "tiny_pi": b, _ = tiny_circuit(degree_bits, seed, 3)
    elif kind == "sha": b, _ = sha_circuit(degree_bits, seed)
    else: b, _ = (sig_circuit if kind == "sig" else rec_circuit)(degree_bits, seed)
    cd = b.build()
    cd.common["hasher"] = hasher
    assert cd.common["degree_bits"] == degree_bits, (cd.common["degree_bits"], degree_bits)
    return cd
