#!/usr/bin/env python3
"""BASELINE.json config 5: Goldilocks NTT + Poseidon-Merkle microbench, 2^16..2^24 rows, achieved algorithmic GB/s vs the
measured HBM peak.  Device-resident inputs, CUDA events on the context stream, best of 5 after 2 warm-ups.
Usage (on a GPU box): python tools/sweep_config5.py > profiles/config5_sweep_r01.jsonl"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-light-client-implementation_b200"))
import torch
import b200_plonky2 as bp
from b200_plonky2 import lib, c_void_p

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); best = 1e30
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best

def main():
    peak = 6574.8
    try: peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception: pass
    torch.cuda.set_device(0)
    ctx = bp.Context(0); stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    P = lambda t: c_void_p(t.data_ptr())
    def chk(rc):
        if rc: raise RuntimeError(lib.b200_last_error(ctx.h))
    for log_n, cols in [(16, 234), (16, 16), (18, 135), (18, 20), (20, 85), (20, 16), (22, 16), (24, 1)]:
        n = 1 << log_n
        g = torch.Generator(device="cuda"); g.manual_seed(log_n * 1000 + cols)
        v = torch.randint(0, 1 << 62, (cols, n), dtype=torch.int64, device="cuda", generator=g)
        scratch = torch.empty_like(v)
        ms = timeit(lambda: chk(lib.b200_ntt_batch_dev(ctx.h, P(v), cols, log_n, 0, P(scratch))))
        rec = {"op": "ntt_forward_natural_to_natural", "log_n": log_n, "cols": cols, "ms": round(ms, 4),
               "alg_GBps": round(16.0 * n * cols / (ms * 1e-3) / 1e9, 1), "frac_of_hbm_peak": round(16.0 * n * cols / (ms * 1e-3) / 1e9 / peak, 4)}
        print(json.dumps(rec), flush=True)
        nd = lib.b200_merkle_num_u64(log_n, 4)
        dg = torch.empty(nd, dtype=torch.int64, device="cuda")
        ms = timeit(lambda: chk(lib.b200_poseidon_merkle_dev(ctx.h, P(v), cols, log_n, 4, P(dg))))
        perms = n * ((cols + 7) // 8 if cols > 4 else 0) + n - 16
        algb = 8.0 * n * cols + 32.0 * (2 * n - 16)
        print(json.dumps({"op": "poseidon_merkle_cap4", "log_L": log_n, "cols": cols, "ms": round(ms, 4), "Mperm_s": round(perms / (ms * 1e-3) / 1e6, 1),
                          "alg_GBps": round(algb / (ms * 1e-3) / 1e9, 1), "frac_of_hbm_peak": round(algb / (ms * 1e-3) / 1e9 / peak, 4)}), flush=True)
        if log_n <= 20 and cols * (n << 3) * 8 < 40e9:
            co = torch.empty_like(v); lde = torch.empty((cols, n << 3), dtype=torch.int64, device="cuda")
            dg2 = torch.empty(lib.b200_merkle_num_u64(log_n + 3, 4), dtype=torch.int64, device="cuda")
            ms = timeit(lambda: chk(lib.b200_lde_commit_dev(ctx.h, P(v), cols, log_n, 3, 4, P(co), P(lde), P(dg2))), iters=3, warm=1)
            algb = 72.0 * n * cols + 512.0 * n
            print(json.dumps({"op": "lde_commit(iNTT+x8 coset LDE+leaf hash+cap)", "log_n": log_n, "cols": cols, "ms": round(ms, 4),
                              "alg_GBps": round(algb / (ms * 1e-3) / 1e9, 1), "frac_of_hbm_peak": round(algb / (ms * 1e-3) / 1e9 / peak, 4)}), flush=True)
            del co, lde, dg2
        del v, scratch, dg
        torch.cuda.empty_cache()

if __name__ == "__main__":
    main()
