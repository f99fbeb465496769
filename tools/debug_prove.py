#!/usr/bin/env python3
"""Triage helper: run the GPU prover with B200_DEBUG_DUMP and compare every intermediate with the C oracle."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-light-client-implementation_b200"))
import numpy as np

def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "rec"; db = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    import b200_plonky2
    from b200_plonky2 import workloads
    from oracle import prover, proof_io as io, verifier, gl
    cd = workloads.build(kind, db, seed=11)
    oc = prover.OracleCircuit(cd.common, cd.constants_sigmas)
    ctx = b200_plonky2.Context(0)
    d = tempfile.mkdtemp(); os.environ["B200_DEBUG_DUMP"] = d
    cd.upload(ctx)
    print("cs cap equal:", np.array_equal(cd.constants_sigmas_cap, oc.cap))
    cpu, dbg = oc.prove(cd.wires, cd.public_inputs, debug=True)
    try:
        gpu = cd.prove(ctx)
    except Exception as e:
        print("GPU prove raised:", e); gpu = None
    print(open(os.path.join(d, "challenges.txt")).read() if os.path.exists(os.path.join(d, "challenges.txt")) else "no challenges file")
    print("oracle:", {k: dbg[k] for k in ("betas", "gammas", "alphas", "zeta", "fri_alpha")})
    n = cd.n
    def load(name, shape, dt=np.uint64):
        p = os.path.join(d, name)
        return np.fromfile(p, dtype=dt).reshape(shape) if os.path.exists(p) else None
    zs = load("zs_pp_values.bin", dbg["zs"].shape)
    if zs is not None: print("zs_pp values equal:", np.array_equal(zs, dbg["zs"]), np.argwhere(zs != dbg["zs"])[:3])
    q = load("quotient_coeffs_br.bin", dbg["q"].shape)
    if q is not None:
        br = np.array([gl.reverse_bits(i, cd.common["degree_bits"]) for i in range(n)])
        qn = q[:, br]                       # natural[k] = br_storage[bitrev(k)]
        print("quotient coeffs equal:", np.array_equal(qn, dbg["q"]), np.argwhere(qn != dbg["q"])[:3])
    f0 = load("fri_values0.bin", dbg["f0"].shape)
    if f0 is not None: print("fri layer0 equal:", np.array_equal(f0, dbg["f0"]), np.argwhere(f0 != dbg["f0"])[:3])
    if gpu is not None:
        print("proof bytes equal:", gpu == cpu, len(gpu), len(cpu))
        if gpu != cpu:
            first = next(i for i in range(min(len(gpu), len(cpu))) if gpu[i] != cpu[i]); print("first differing byte:", first)
        cdo = prover.common_data(cd.common)
        try: print("verifier:", verifier.verify(cdo, oc.verifier_only(), io.proof_from_bytes(gpu, cdo)))
        except Exception as e: print("verifier rejected:", e)

if __name__ == "__main__":
    main()
