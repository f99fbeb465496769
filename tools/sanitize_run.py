#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): proves and verifies a few small circuits."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-light-client-implementation_b200"))
import numpy as np
import b200_plonky2 as bp
from b200_plonky2 import workloads
ctx = bp.Context(0)
for kind, db, hs in (("rec", 7, "gl"), ("sig", 9, "gl"), ("sha", 7, "gl"), ("tiny", 3, "gl"), ("rec", 6, "bn254")):
    cd = workloads.build(kind, db, seed=5, hasher=hs)
    p = cd.prove(ctx)
    assert cd.verify(p)
    cd.free()
v = np.arange(3 << 13, dtype=np.uint64).reshape(3, 1 << 13)
assert np.array_equal(ctx.ntt_batch(ctx.ntt_batch(v, 13), 13, inverse=True), v)
print("sanitize run ok")
