#!/usr/bin/env python3
"""BASELINE config 4 (next row): BN254 F_r NTT timing (the FFTs of groth16.Prove), device-resident, CUDA events, best of 3.
Usage: fr_ntt_bench.py [log_n ...]   (default 20 22 24).  Algorithmic bytes = 64 n (read + write 32-byte elements)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-light-client-implementation_b200"))
import numpy as np, torch
import b200_plonky2 as bp
from b200_plonky2 import lib, c_void_p

def main():
    peak = 6574.8
    try: peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception: pass
    torch.cuda.set_device(0)
    ctx = bp.Context(0); stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    rng = np.random.default_rng(3)
    for logn in (int(a) for a in (sys.argv[1:] or ["20", "22", "24"])):
        n = 1 << logn
        raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64); raw[:, 3] &= np.uint64((1 << 60) - 1)
        A = torch.from_numpy(raw.view(np.int64)).cuda(); B = torch.empty_like(A); C = torch.empty_like(A)
        def run(src, dst, inv, cs):
            rc = lib.b200_fr_ntt_dev(ctx.h, c_void_p(src.data_ptr()), logn, inv, cs, c_void_p(dst.data_ptr()))
            if rc: raise RuntimeError(lib.b200_last_error(ctx.h))
        rec = {"op": "bn254_fr_ntt", "log_n": logn}
        for name, inv, cs in (("forward", 0, 0), ("forward_coset", 0, 1), ("inverse_coset", 1, 1)):
            run(A, B, inv, cs); torch.cuda.synchronize(); best = 1e30
            for _ in range(3):
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(); run(A, B, inv, cs); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
            rec[name + "_ms"] = round(best, 3)
        ctx.enable_timing(True); ctx.timings(); run(A, B, 0, 1); ph = ctx.timings(); ctx.enable_timing(False)
        rec["phases_ms_forward_coset"] = {k: round(v[0], 3) for k, v in ph.items() if not k.startswith("_")}
        run(B, C, 1, 1); torch.cuda.synchronize()
        rec["round_trip_exact"] = bool(torch.equal(A, C))
        ms = rec["forward_ms"]
        rec["Melem_per_s"] = round(n / ms / 1e3, 1); rec["alg_GBps"] = round(64.0 * n / (ms * 1e-3) / 1e9, 1); rec["frac_of_hbm_peak"] = round(64.0 * n / (ms * 1e-3) / 1e9 / peak, 4)
        print(json.dumps(rec), flush=True)
        del A, B, C; torch.cuda.empty_cache()
if __name__ == "__main__":
    main()
