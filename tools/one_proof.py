#!/usr/bin/env python3
"""One proof of one shape with no warm-up: the target process for `ncu -k regex:... -c N` captures.
Usage: python tools/one_proof.py sig 17 [gl|bn254]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-light-client-implementation_b200"))
import b200_plonky2 as bp
from b200_plonky2 import workloads
kind = sys.argv[1]; db = int(sys.argv[2]); hs = sys.argv[3] if len(sys.argv) > 3 else "gl"
ctx = bp.Context(0)
cd = workloads.build(kind, db, seed=3, hasher=hs); cd.upload(ctx)
print(len(cd.prove(ctx)), "proof bytes")
