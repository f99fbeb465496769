#!/usr/bin/env python3
"""BASELINE config 4 (next row): BN254 G1 / G2 MSM timing, N in {2^18, 2^20, 2^22}, uniform 252-bit scalars, points drawn from
a pool of distinct curve points.  Device-resident inputs, CUDA events, best of 3, per-phase times of the last run.
Usage: msm_bench.py [g1|g2] [log_n ...].  Output: JSON lines."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-light-client-implementation_b200"))
import numpy as np, torch
import b200_plonky2 as bp
from b200_plonky2 import lib, c_void_p
from oracle import bn254_curve as ec          # input generation + spot check only (tools/, not the product path)

def limbs(x): return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
def main():
    peak = 6574.8
    try: peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception: pass
    torch.cuda.set_device(0)
    ctx = bp.Context(0); stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    args = sys.argv[1:]
    g2 = bool(args) and args[0] == "g2"
    if args and args[0] in ("g1", "g2"): args = args[1:]
    if g2:
        p = ec.G2; pool = []
        for k in range(1024): pool.append(p); p = ec.g2_add(p, ec.G2)
        pool_np = np.array([limbs(q[0][0]) + limbs(q[0][1]) + limbs(q[1][0]) + limbs(q[1][1]) for q in pool], dtype=np.uint64)
    else:
        p = ec.G; pool = []
        for k in range(4096): pool.append(p); p = ec.add(p, ec.G)
        pool_np = np.array([limbs(q[0]) + limbs(q[1]) for q in pool], dtype=np.uint64)
    fn = lib.b200_msm_g2_dev if g2 else lib.b200_msm_g1_dev
    pt_bytes = 128 if g2 else 64
    rng = np.random.default_rng(7)
    for logn in (int(a) for a in (args or ["18", "20", "22"])):
        n = 1 << logn
        idx = rng.integers(0, len(pool), size=n)
        P = torch.from_numpy(pool_np[idx].view(np.int64)).cuda()
        raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
        raw[:, 3] &= np.uint64((1 << 60) - 1)                                       # < 2^252 < r
        S = torch.from_numpy(raw.view(np.int64)).cuda()
        out = np.zeros(16 if g2 else 8, dtype=np.uint64)
        def run():
            rc = fn(ctx.h, c_void_p(P.data_ptr()), c_void_p(S.data_ptr()), n, out.ctypes.data_as(c_void_p))
            if rc: raise RuntimeError(lib.b200_last_error(ctx.h))
        run(); torch.cuda.synchronize(); best = 1e30
        for _ in range(3):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
        ctx.enable_timing(True); ctx.timings(); run(); ph = ctx.timings(); ctx.enable_timing(False)
        ab = pt_bytes + 32.0
        rec = {"op": "bn254_g2_msm" if g2 else "bn254_g1_msm", "log_n": logn, "ms": round(best, 3), "Mpoints_per_s": round(n / best / 1e3, 2),
               "alg_GBps": round(ab * n / (best * 1e-3) / 1e9, 2), "frac_of_hbm_peak": round(ab * n / (best * 1e-3) / 1e9 / peak, 5),
               "phases_ms": {k: round(v[0], 3) for k, v in ph.items() if not k.startswith("_")}}
        if logn <= 18 and not g2:     # closed-form check through the discrete logs of the pool
            ks = [int(i) + 1 for i in idx]
            sc = [sum(int(raw[j, t]) << (64 * t) for t in range(4)) for j in range(n)]
            exp = ec.mul(sum(s * k for s, k in zip(sc, ks)) % ec.R, ec.G)
            got = (sum(int(out[i]) << (64 * i) for i in range(4)), sum(int(out[4 + i]) << (64 * i) for i in range(4)))
            rec["matches_oracle"] = (got == exp)
        print(json.dumps(rec), flush=True)
        del P, S; torch.cuda.empty_cache()
if __name__ == "__main__":
    main()
