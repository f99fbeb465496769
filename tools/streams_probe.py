#!/usr/bin/env python3
"""Throughput of S concurrent proving contexts (one host thread + stream each) on one GPU, sig-shaped 2^17 x 234 proofs.
Usage: streams_probe.py [log_n] [S ...]"""
import os, sys, json, time, threading, ctypes, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-light-client-implementation_b200"))
import numpy as np, torch
import b200_plonky2 as bp
from b200_plonky2 import workloads, lib, c_void_p, c_size_t
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 17
Ss = [int(a) for a in sys.argv[2:]] or [1, 2, 3]
torch.cuda.set_device(0)
cd0 = workloads.build("sig", log_n, seed=1)
host = torch.from_numpy(cd0.wires).pin_memory()
pis = np.array(cd0.public_inputs, dtype=np.uint64)
for S in Ss:
    ctxs = [bp.Context(0) for _ in range(S)]
    cds = []
    for c in ctxs:
        d = copy.copy(cd0)          # same circuit, its own device copy per context
        d._handle = None; d.upload(c); cds.append(d)
    dev = [host.to("cuda") for _ in range(S)]
    def prove(i, on_dev):
        out = ctypes.POINTER(ctypes.c_uint8)(); ln = c_size_t()
        fn = lib.b200_prove_dev if on_dev else lib.b200_prove
        ptr = dev[i].data_ptr() if on_dev else host.data_ptr()
        rc = fn(ctxs[i].h, cds[i]._handle, c_void_p(ptr), pis.ctypes.data_as(c_void_p), len(pis), ctypes.c_uint64(0), ctypes.byref(out), ctypes.byref(ln))
        assert rc == 0, lib.b200_last_error(ctxs[i].h)
        lib.b200_free(out)
    res = {"streams": S, "log_n": log_n}
    for on_dev in (True, False):
        for i in range(S): prove(i, on_dev); prove(i, on_dev)
        torch.cuda.synchronize()
        K = 6
        def work(i):
            for _ in range(K): prove(i, on_dev)
        ths = [threading.Thread(target=work, args=(i,)) for i in range(S)]
        t0 = time.perf_counter()
        for t in ths: t.start()
        for t in ths: t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res["proofs_per_s_" + ("dev" if on_dev else "host")] = round(S * K / dt, 3)
    print(json.dumps(res), flush=True)
    for d in cds: d.free()
    for c in ctxs: c.close()
    del dev; torch.cuda.empty_cache()
