"""b200_verify (the product's host-side CircuitData::verify) against the reference's golden proofs (no GPU needed):
it must accept the three real-block BN128 wrapper proofs and every proof of the oracle prover (both hashers, all 21 gate
kinds), and reject tampered ones."""
import json, os, struct
import numpy as np
import pytest
import b200_plonky2 as bp

G = os.path.join(os.path.dirname(__file__), "golden")
NEAR = sorted(os.listdir(os.path.join(G, "near_final_proofs")))

def _limbs(x): x = int(x); return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]

@pytest.mark.parametrize("name", NEAR)
def test_accepts_golden_near_wrapper_proofs(name):
    d = os.path.join(G, "near_final_proofs", name)
    common = bp.common_from_json(json.load(open(f"{d}/common_data.json")), hasher="bn254")
    vd = json.load(open(f"{d}/verifier_data.json"))
    cap = [_limbs(h) for h in vd["constants_sigmas_cap"]]
    raw = open(f"{d}/proof.bin", "rb").read()
    ok, why = bp.verify(common, cap, _limbs(vd["circuit_digest"]), raw)
    assert ok, why
    # tamper: a wire opening, a Merkle sibling byte, the PoW witness, a public input
    for off in (1536 + 16 * 90, 20000, len(raw) - 97 * 8 - 16, len(raw) - 8 * 50):
        bad = bytearray(raw); bad[off] ^= 1
        ok, why = bp.verify(common, cap, _limbs(vd["circuit_digest"]), bytes(bad))
        assert not ok and why
    ok, why = bp.verify(common, cap, _limbs(vd["circuit_digest"]), raw[:-8])
    assert not ok

def test_accepts_gnark_fixture():
    import gzip
    from oracle import proof_io as io
    d = os.path.join(G, "gnark_test_circuit")
    cj = json.load(open(f"{d}/common_circuit_data.json"))
    common = bp.common_from_json(cj, hasher="bn254")
    vd = json.load(open(f"{d}/verifier_only_circuit_data.json"))
    pr = io.proof_from_json(json.load(gzip.open(f"{d}/proof_with_public_inputs.json.gz")), "bn254")
    raw = io.proof_to_bytes(pr, "bn254")
    ok, why = bp.verify(common, [_limbs(h) for h in vd["constants_sigmas_cap"]], _limbs(vd["circuit_digest"]), raw)
    assert ok, why

@pytest.mark.parametrize("kind,db,hasher", [("sig", 8, "gl"), ("rec", 6, "gl"), ("rec", 6, "bn254"), ("tiny", 3, "gl"), ("tiny_pi", 5, "gl")])
def test_accepts_oracle_prover_proofs_and_rejects_wrong_witness(kind, db, hasher):
    from b200_plonky2 import workloads
    from oracle import prover
    cd = workloads.build(kind, db, seed=13, hasher=hasher)
    oc = prover.OracleCircuit(cd.common, cd.constants_sigmas)
    pb = oc.prove(cd.wires, cd.public_inputs)
    ok, why = bp.verify(cd.common, oc.cap, oc.digest, pb)
    assert ok, why
    w = cd.wires.copy(); w[0, min(5, cd.n - 1)] ^= 1
    ok, why = bp.verify(cd.common, oc.cap, oc.digest, oc.prove(w, cd.public_inputs))
    assert not ok


@pytest.mark.parametrize("name", NEAR + ["gnark"])
def test_circuit_digest_reproduces_golden_verifier_data(name):
    """b200_circuit_digest (the digest formula b200_circuit_build / b200_circuit_upload apply, BN128 hasher) on the golden
    constants_sigmas_cap == the circuit_digest stored next to it by the reference (G6: the three final wrapper circuits, all
    the same deterministic circuit; G5: the gnark fixture)."""
    if name == "gnark":
        d = os.path.join(G, "gnark_test_circuit")
        vd = json.load(open(f"{d}/verifier_only_circuit_data.json")); cj = json.load(open(f"{d}/common_circuit_data.json"))
    else:
        d = os.path.join(G, "near_final_proofs", name)
        vd = json.load(open(f"{d}/verifier_data.json")); cj = json.load(open(f"{d}/common_data.json"))
    cap = [_limbs(h) for h in vd["constants_sigmas_cap"]]
    got = bp.circuit_digest("bn254", cap, cj["fri_params"]["degree_bits"])
    assert got == _limbs(vd["circuit_digest"])

def test_circuit_digest_gl_matches_host_mirror():
    from b200_plonky2 import digest
    rng = np.random.default_rng(3)
    cap = rng.integers(0, 1 << 63, size=(16, 4), dtype=np.uint64)
    assert bp.circuit_digest("gl", cap, 13) == [int(x) for x in digest.circuit_digest(cap, 13)]
