#!/usr/bin/env python3
"""Per-phase CUDA-event timings of one proof shape.  Usage: python tools/phase_breakdown.py rec 13 [gl|bn254]"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-light-client-implementation_b200"))
import b200_plonky2 as bp
from b200_plonky2 import workloads
kind = sys.argv[1]; db = int(sys.argv[2]); hs = sys.argv[3] if len(sys.argv) > 3 else "gl"
ctx = bp.Context(0)
cd = workloads.build(kind, db, seed=3, hasher=hs); cd.upload(ctx)
w = ctx.pinned_copy(cd.wires) if os.environ.get('PAGEABLE') != '1' else cd.wires
for _ in range(3): cd.prove(ctx, wires=w)
ctx.enable_timing(True); ctx.timings()
l0 = ctx.kernel_launches(); t0 = time.time()
N = 10
for _ in range(N): cd.prove(ctx, wires=w)
wall = (time.time() - t0) / N * 1e3
ph = ctx.timings()
print(json.dumps({"shape": f"{kind} 2^{db} {hs}", "wall_ms_per_proof": round(wall, 3), "launches_per_proof": (ctx.kernel_launches() - l0) / N,
                  "phases_ms": {k: round(v[0] / N, 4) for k, v in ph.items()}, "sum_phases_ms": round(sum(v[0] for k, v in ph.items() if not k.startswith("_")) / N, 3),
                  "host_busy_ms": round((ph["_host_total"][0] - ph["_host_sync_wait"][0]) / N, 3) if "_host_total" in ph else None}))
