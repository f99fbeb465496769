"""Used by __graft_entry__.smoke(): one tiny end-to-end proof on cuda:0, checked against the oracle (checker only)."""
def run(ctx):
    from . import workloads
    from oracle import prover, proof_io as io, verifier
    cd = workloads.build("rec", 6, seed=1)
    oc = prover.OracleCircuit(cd.common, cd.constants_sigmas)
    gpu = cd.prove(ctx)
    assert gpu == oc.prove(cd.wires, cd.public_inputs), "GPU proof bytes != oracle prover"
    cdo = prover.common_data(cd.common)
    assert verifier.verify(cdo, oc.verifier_only(), io.proof_from_bytes(gpu, cdo))
    cd.free()
    print("smoke ok: b200_prove (13-gate recursion-shaped circuit, 2^6 rows) byte-identical to the oracle prover and verified")
