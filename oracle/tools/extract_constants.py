#!/usr/bin/env python3
"""Extract the published Poseidon parameter tables (numbers only) from the reference tree.

TEST INFRASTRUCTURE.  Run once in the build container (where /root/reference exists); the
JSON it writes under oracle/data/ is committed so nothing at test time reads the reference.

Sources (data tables, not code):
  gnark-plonky2-verifier/poseidon/goldilocks_constants.go:7-1199   Poseidon-Goldilocks t=12
  gnark-plonky2-verifier/poseidon/bn254_constants.go:15-536        Poseidon-BN254 t=4 (iden3)
"""
import json, re, sys, os

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(__file__), "..", "data")

def gl_tables():
    src = open(f"{REF}/gnark-plonky2-verifier/poseidon/goldilocks_constants.go").read()
    out = {}
    # split the file on "var NAME =" blocks and pull every uint64 literal out of each block
    blocks = re.split(r"\nvar (\w+) = ", src)
    for name, body in zip(blocks[1::2], blocks[2::2]):
        body = body.replace("ZERO_VAR", "uint64(0)")
        nums = [int(x, 0) for x in re.findall(r"uint64\((0x[0-9a-fA-F]+|\d+)\)", body)]
        out[name] = nums
    keep = {
        "ALL_ROUND_CONSTANTS": 360, "MDS_MATRIX_CIRC": 12, "MDS_MATRIX_DIAG": 12, "MDS0TO0": 1,
        "FAST_PARTIAL_FIRST_ROUND_CONSTANT": 12, "FAST_PARTIAL_ROUND_CONSTANTS": 22,
        "FAST_PARTIAL_ROUND_VS": 22 * 11, "FAST_PARTIAL_ROUND_W_HATS": 22 * 11,
        "FAST_PARTIAL_ROUND_INITIAL_MATRIX": 11 * 11,
    }
    res = {}
    for k, n in keep.items():
        assert len(out[k]) == n, (k, len(out[k]), n)
        res[k] = out[k]
    return res

def bn254_tables():
    src = open(f"{REF}/gnark-plonky2-verifier/poseidon/bn254_constants.go").read()
    c = {int(i): int(v) for i, v in re.findall(r'cConstants\[(\d+)\], _ = new\(big\.Int\)\.SetString\("(\d+)"', src)}
    s = {int(i): int(v) for i, v in re.findall(r'sConstants\[(\d+)\], _ = new\(big\.Int\)\.SetString\("(\d+)"', src)}
    m = {(int(i), int(j)): int(v) for i, j, v in re.findall(r'mMatrix\[(\d)\]\[(\d)\], _ = new\(big\.Int\)\.SetString\("(\d+)"', src)}
    p = {(int(i), int(j)): int(v) for i, j, v in re.findall(r'pMatrix\[(\d)\]\[(\d)\], _ = new\(big\.Int\)\.SetString\("(\d+)"', src)}
    assert len(c) == 88 and len(s) == 392 and len(m) == 16 and len(p) == 16
    return {
        "C": [str(c[i]) for i in range(88)],
        "S": [str(s[i]) for i in range(392)],
        "M": [[str(m[(i, j)]) for j in range(4)] for i in range(4)],
        "P": [[str(p[(i, j)]) for j in range(4)] for i in range(4)],
    }

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    json.dump(gl_tables(), open(os.path.join(OUT, "poseidon_gl_constants.json"), "w"))
    json.dump(bn254_tables(), open(os.path.join(OUT, "poseidon_bn254_constants.json"), "w"))
    print("ok")
