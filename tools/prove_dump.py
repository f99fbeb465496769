#!/usr/bin/env python3
"""Byte-parity harness for a Rust box: load a circuit + witness DUMPED from the
reference prover and prove it with b200_prove; if the dump carries the reference's proof, compare the bytes.

Dump directory layout (what the INTEGRATION.md shim writes with `B200_DUMP_DIR` set, all little-endian):
  common_data.json        plonky2 CommonCircuitData as serde writes it (prove_block.rs:343-356 does this for the final proof)
  constants_sigmas.bin    u64[(num_constants + num_routed_wires) * n]  polynomial VALUES on H, column-major
  wires.bin               u64[num_wires * n]                           the full witness (PartitionWitness::full_witness), column-major
  public_inputs.bin       u64[num_public_inputs]
  verifier_only.json      optional: {"constants_sigmas_cap": [...], "circuit_digest": ...} (JSON as in verifier_data.json)
  proof.bin               optional: ProofWithPublicInputs::to_bytes() of the reference for THIS witness (run it with
                          RAYON_NUM_THREADS=1 so that its PoW nonce is the minimum one; otherwise pass --use-ref-pow)
  hasher                  optional text file: "gl" (default) or "bn254"
Usage: python tools/prove_dump.py DIR [--use-ref-pow] [--out proof_b200.bin]
Exit code 0: proof produced, accepted by b200_verify, and (if proof.bin is present) byte-identical to it."""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-light-client-implementation_b200"))

def load_dump(d):
    import b200_plonky2 as bp
    from b200_plonky2 import circuit_builder as cb
    hasher = open(os.path.join(d, "hasher")).read().strip() if os.path.exists(os.path.join(d, "hasher")) else "gl"
    common = bp.common_from_json(json.load(open(os.path.join(d, "common_data.json"))), hasher=hasher)
    n = 1 << common["degree_bits"]
    cs = np.fromfile(os.path.join(d, "constants_sigmas.bin"), dtype="<u8")
    cols = common["num_constants"] + common["num_routed_wires"]
    if cs.size != cols * n: raise SystemExit(f"constants_sigmas.bin: expected {cols} x {n} u64, found {cs.size}")
    wires = np.fromfile(os.path.join(d, "wires.bin"), dtype="<u8")
    if wires.size != common["num_wires"] * n: raise SystemExit(f"wires.bin: expected {common['num_wires']} x {n} u64, found {wires.size}")
    pis = np.fromfile(os.path.join(d, "public_inputs.bin"), dtype="<u8")
    if pis.size != common["num_public_inputs"]: raise SystemExit("public_inputs.bin: wrong length")
    return cb.CircuitData(common, np.ascontiguousarray(cs.reshape(cols, n)), np.ascontiguousarray(wires.reshape(common["num_wires"], n)), [int(x) for x in pis])

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("dir"); ap.add_argument("--use-ref-pow", action="store_true"); ap.add_argument("--out")
    a = ap.parse_args()
    import b200_plonky2 as bp
    cd = load_dump(a.dir)
    ctx = bp.Context(0)
    cd.upload(ctx)
    vo = os.path.join(a.dir, "verifier_only.json")
    if os.path.exists(vo):
        v = json.load(open(vo))
        lim = lambda x: [(int(x) >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
        if cd.common.get("hasher") == "bn254": cap = [lim(h) for h in v["constants_sigmas_cap"]]; dg = lim(v["circuit_digest"])
        else: cap = [[int(e) for e in h["elements"]] for h in v["constants_sigmas_cap"]]; dg = [int(e) for e in v["circuit_digest"]["elements"]]
        ok = np.array_equal(np.array(cap, dtype=np.uint64), cd.constants_sigmas_cap) and list(dg) == list(cd.circuit_digest)
        print("verifier_only (constants_sigmas_cap, circuit_digest) == b200_circuit_upload:", ok)
        if not ok: sys.exit(2)
    ref = open(os.path.join(a.dir, "proof.bin"), "rb").read() if os.path.exists(os.path.join(a.dir, "proof.bin")) else None
    hint = 0
    if ref is not None and a.use_ref_pow:
        hint = int.from_bytes(ref[len(ref) - 8 * (len(cd.public_inputs) + 2):][:8], "little")     # pow_witness precedes the PI count
    proof = cd.prove(ctx, pow_witness_hint=hint)
    print(f"b200_prove: {len(proof)} bytes; b200_verify:", cd.verify(proof))
    if a.out: open(a.out, "wb").write(proof)
    if ref is not None:
        same = proof == ref
        if not same:
            k = next((i for i, (x, y) in enumerate(zip(proof, ref)) if x != y), min(len(proof), len(ref)))
            print(f"MISMATCH vs the reference proof: first differing byte {k} (lengths {len(proof)} / {len(ref)})")
            sys.exit(3)
        print("proof bytes == reference proof.bin")
    cd.free(); ctx.close()

if __name__ == "__main__":
    main()
