"""The C oracle prover (restating upstream plonky2's prove()) must produce proofs that the pinned oracle verifier
accepts, for both workload shapes; the circuit-digest formula is pinned on the golden BN128 wrapper data."""
import json, os, copy
import pytest
from oracle import prover, proof_io as io, verifier

G = os.path.join(os.path.dirname(__file__), "golden")

def test_circuit_digest_formula_matches_golden_wrapper():
    vd = json.load(open(os.path.join(G, "near_final_proofs", "random_CGZPhFRk", "verifier_data.json")))
    cap = [int(x) for x in vd["constants_sigmas_cap"]]
    assert prover.circuit_digest_bn254(cap, 12) == int(vd["circuit_digest"])
    assert int(vd["circuit_digest"]) == 15612857049394414280467902980451102556370837276268678037246302844075472795365

@pytest.mark.parametrize("kind,db", [("sig", 8), ("rec", 6), ("rec", 9), ("tiny", 2), ("tiny", 3), ("tiny_pi", 5), ("tiny", 6), ("sha", 7)])
def test_oracle_prover_accepted_by_oracle_verifier(kind, db):
    from b200_plonky2 import workloads
    cd = workloads.build(kind, db, seed=5)
    oc = prover.OracleCircuit(cd.common, cd.constants_sigmas)
    pb = oc.prove(cd.wires, cd.public_inputs)
    cdo = prover.common_data(cd.common)
    pr = io.proof_from_bytes(pb, cdo)
    assert io.proof_to_bytes(pr, "gl") == pb
    assert pr.public_inputs == cd.public_inputs
    assert verifier.verify(cdo, oc.verifier_only(), pr)
    # minimum-nonce PoW: no smaller witness passes
    assert pr.pow_witness < (1 << 22)
    # a wrong witness must not yield an accepted proof
    w = cd.wires.copy(); w[0, min(5, cd.n - 1)] ^= 1
    pb2 = oc.prove(w, cd.public_inputs)
    with pytest.raises(verifier.VerifyError):
        verifier.verify(cdo, oc.verifier_only(), io.proof_from_bytes(pb2, cdo))

def test_pow_hint_semantics():
    """hint == a valid nonce -> used as is; invalid hint -> falls back to the minimum-nonce grind."""
    from b200_plonky2 import workloads
    cd = workloads.build("rec", 6, seed=2)
    oc = prover.OracleCircuit(cd.common, cd.constants_sigmas)
    cdo = prover.common_data(cd.common)
    b0 = oc.prove(cd.wires, cd.public_inputs)
    p0 = io.proof_from_bytes(b0, cdo)
    assert oc.prove(cd.wires, cd.public_inputs, pow_hint=p0.pow_witness) == b0
    bad_hint = p0.pow_witness + 1 if p0.pow_witness else 1
    b1 = oc.prove(cd.wires, cd.public_inputs, pow_hint=bad_hint)
    p1 = io.proof_from_bytes(b1, cdo)
    assert p1.pow_witness in (p0.pow_witness, bad_hint) and verifier.verify(cdo, oc.verifier_only(), p1)

def test_c_bn254_permutation_kats():
    """C oracle Poseidon-BN254 vs the reference's 4 permutation vectors (crypto/plonky2_bn128/src/poseidon_bn128.rs:133-180)."""
    import ctypes, numpy as np
    from oracle import cbind
    kats = json.load(open(os.path.join(G, "kats.json")))["poseidon_bn254_permutation"]
    for v in kats:
        st = np.zeros(16, dtype=np.uint64)
        for i, x in enumerate(v["in"]):
            for k in range(4): st[4 * i + k] = (int(x) >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
        cbind.lib().orc_bn254_permute(st.ctypes.data_as(ctypes.c_void_p))
        out = [sum(int(st[4 * i + k]) << (64 * k) for k in range(4)) for i in range(4)]
        assert out == [int(x) for x in v["out"]]

def test_oracle_prover_bn128_config_accepted_by_verifier():
    """PoseidonBN128GoldilocksConfig: the verifier instance that accepts the reference's golden wrapper proofs must
    accept the oracle prover's BN128-config proofs (13-gate recursion-shaped circuit)."""
    from b200_plonky2 import workloads
    cd = workloads.build("rec", 6, seed=8, hasher="bn254")
    oc = prover.OracleCircuit(cd.common, cd.constants_sigmas)
    pb = oc.prove(cd.wires, cd.public_inputs)
    cdo = prover.common_data(cd.common)
    assert cdo.hasher == "bn254"
    pr = io.proof_from_bytes(pb, cdo)
    assert io.proof_to_bytes(pr, "bn254") == pb
    assert verifier.verify(cdo, oc.verifier_only(), pr)

def test_product_host_bn254_digest_matches_oracle():
    from b200_plonky2 import poseidon_bn254_py as pp
    from oracle import poseidon_bn254 as pbn
    assert pp.permute([0, 1, 2, 3]) == pbn.permute([0, 1, 2, 3])
    cap = [[i * 7 + 1, i * 3, 5, 9] for i in range(16)]
    ints = [sum(h[i] << (64 * i) for i in range(4)) for h in cap]
    d = prover.circuit_digest_bn254(ints, 12)
    assert pp.circuit_digest(cap, 12) == [(d >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
