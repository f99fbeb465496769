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

@pytest.mark.parametrize("kind,db", [("sig", 8), ("rec", 6), ("rec", 9)])
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
    w = cd.wires.copy(); w[3, 5] ^= 1
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
