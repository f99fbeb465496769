#!/usr/bin/env python3
"""Scratch micro-benchmark of the inner-loop kernels (device-resident inputs, CUDA events on the ctx stream).
Not the judged bench (that is bench.py); used to pick what to optimise.  Usage: python tools/microbench.py [log_n] [cols]"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-light-client-implementation_b200"))
import torch
import b200_plonky2 as bp
from b200_plonky2 import lib, c_void_p

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]

def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 17
    cols = int(sys.argv[2]) if len(sys.argv) > 2 else 234
    torch.cuda.set_device(0)
    ctx = bp.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    n = 1 << log_n; N = n << 3
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    vals = torch.randint(0, (1 << 62), (cols, n), dtype=torch.int64, device="cuda", generator=g)
    coeffs = torch.empty_like(vals); lde = torch.empty((cols, N), dtype=torch.int64, device="cuda")
    nd = lib.b200_merkle_num_u64(log_n + 3, 4)
    dg = torch.empty(nd, dtype=torch.int64, device="cuda")
    scratch = torch.empty_like(vals)
    P = lambda t: c_void_p(t.data_ptr())
    res = {}
    def chk(rc):
        if rc: raise RuntimeError(lib.b200_last_error(ctx.h))
    # full lde_commit
    t = timeit(lambda: chk(lib.b200_lde_commit_dev(ctx.h, P(vals), cols, log_n, 3, 4, P(coeffs), P(lde), P(dg))))
    res["lde_commit_ms"] = t
    # NTT alone (forward natural->natural incl. bitrev permutation)
    t = timeit(lambda: chk(lib.b200_ntt_batch_dev(ctx.h, P(vals), cols, log_n, 0, P(scratch))))
    res["ntt_batch_ms"] = t
    res["ntt_GBps_alg"] = 16.0 * n * cols / (t[0] * 1e-3) / 1e9
    # Merkle alone on the LDE
    t = timeit(lambda: chk(lib.b200_poseidon_merkle_dev(ctx.h, P(lde), cols, log_n + 3, 4, P(dg))))
    res["poseidon_merkle_ms"] = t
    perms = N * ((cols + 7) // 8) + N - 16
    res["poseidon_Mperm_s"] = perms / (t[0] * 1e-3) / 1e6
    res["merkle_GBps_alg"] = (8.0 * N * cols + 32.0 * (2 * N - 16)) / (t[0] * 1e-3) / 1e9
    lde_ms = res["lde_commit_ms"][0] - res["poseidon_merkle_ms"][0]
    res["lde_only_ms_est"] = lde_ms
    res["lde_GBps_alg"] = 72.0 * n * cols / (lde_ms * 1e-3) / 1e9
    res["config"] = {"log_n": log_n, "cols": cols}
    print(json.dumps(res))

if __name__ == "__main__":
    main()
