"""tools/prove_dump.py: the one-command byte-parity harness for a box with the Rust reference.  Here the
"reference dump" is written from a synthetic circuit and the oracle prover's proof stands in for the Rust proof."""
import json, os, subprocess, sys
import numpy as np
import pytest
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

def write_dump(cd, d, proof=None, verifier_only=None):
    c = cd.common
    cj = {"config": {"num_wires": c["num_wires"], "num_routed_wires": c["num_routed_wires"], "num_challenges": c["num_challenges"]},
          "fri_params": {"degree_bits": c["degree_bits"], "reduction_arity_bits": c["reduction_arity_bits"],
                         "config": {"rate_bits": c["rate_bits"], "cap_height": c["cap_height"], "proof_of_work_bits": c["proof_of_work_bits"], "num_query_rounds": c["num_query_rounds"]}},
          "num_constants": c["num_constants"], "quotient_degree_factor": c["quotient_degree_factor"], "num_partial_products": c["num_partial_products"],
          "num_gate_constraints": c["num_gate_constraints"], "num_public_inputs": c["num_public_inputs"], "gates": c["gates"],
          "selectors_info": {"selector_indices": c["selector_indices"], "groups": [{"start": s, "end": e} for s, e in c["groups"]]}, "k_is": c["k_is"]}
    json.dump(cj, open(os.path.join(d, "common_data.json"), "w"))
    cd.constants_sigmas.astype("<u8").tofile(os.path.join(d, "constants_sigmas.bin"))
    cd.wires.astype("<u8").tofile(os.path.join(d, "wires.bin"))
    np.array(cd.public_inputs, dtype="<u8").tofile(os.path.join(d, "public_inputs.bin"))
    if proof is not None: open(os.path.join(d, "proof.bin"), "wb").write(proof)
    if verifier_only is not None: json.dump(verifier_only, open(os.path.join(d, "verifier_only.json"), "w"))

def test_dump_loader_round_trip(tmp_path):
    import prove_dump
    from b200_plonky2 import workloads
    cd = workloads.build("tiny_pi", 4, seed=2)
    write_dump(cd, str(tmp_path))
    back = prove_dump.load_dump(str(tmp_path))
    assert all(back.common[k] == cd.common[k] for k in back.common if k in cd.common)
    assert np.array_equal(back.constants_sigmas, cd.constants_sigmas) and np.array_equal(back.wires, cd.wires) and back.public_inputs == cd.public_inputs

@pytest.mark.gpu
def test_prove_dump_matches_reference_proof_bytes(tmp_path):
    from b200_plonky2 import workloads
    from oracle import prover
    cd = workloads.build("rec", 7, seed=8)
    oc = prover.OracleCircuit(cd.common, cd.constants_sigmas)
    ref = oc.prove(cd.wires, cd.public_inputs)
    vo = {"constants_sigmas_cap": [{"elements": [int(x) for x in h]} for h in oc.cap], "circuit_digest": {"elements": [int(x) for x in oc.digest]}}
    write_dump(cd, str(tmp_path), proof=ref, verifier_only=vo)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prove_dump.py"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0 and "proof bytes == reference proof.bin" in r.stdout, r.stdout + r.stderr
    bad = bytearray(ref); bad[100] ^= 1
    open(tmp_path / "proof.bin", "wb").write(bytes(bad))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prove_dump.py"), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 3 and "MISMATCH" in r.stdout
