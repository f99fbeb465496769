"""ctypes binding of the C oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE."""
import ctypes, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

def build(force=False):
    srcs = [os.path.join(_HERE, "c", f) for f in os.listdir(os.path.join(_HERE, "c")) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        cs = [s for s in srcs if s.endswith(".c")]
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-fPIC", "-Wall", "-Wno-unused-function", "-shared",
                               "-o", _SO] + cs)
    return _SO

_lib = None
def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_merkle_num_u64.restype = ctypes.c_size_t
    return _lib

def _p(a): return a.ctypes.data_as(ctypes.c_void_p)
_sz = ctypes.c_size_t

def poseidon_permute(state):
    a = np.array(state, dtype=np.uint64); lib().orc_poseidon_permute(_p(a)); return a

def ntt_batch(values, log_n, inverse=False):
    a = np.array(values, dtype=np.uint64, copy=True)
    lib().orc_ntt_batch(_p(a), _sz(a.shape[0]), log_n, int(inverse)); return a

def merkle_num_u64(log_L, cap_height): return lib().orc_merkle_num_u64(log_L, cap_height)

def merkle_colmajor(data, log_L, cap_height):
    a = np.ascontiguousarray(data, dtype=np.uint64)
    dg = np.zeros(merkle_num_u64(log_L, cap_height), dtype=np.uint64)
    lib().orc_merkle_colmajor(_p(a), _sz(a.shape[0]), log_L, cap_height, _p(dg))
    return dg

def commit_values(values, log_n, rate_bits=3, cap_height=4):
    """-> (coeffs [cols][n] natural, lde [cols][8n] bit-reversed rows, digests (all levels), cap [2^h][4])"""
    v = np.ascontiguousarray(values, dtype=np.uint64)
    cols, n = v.shape[0], 1 << log_n
    co = np.zeros((cols, n), dtype=np.uint64); lde = np.zeros((cols, n << rate_bits), dtype=np.uint64)
    dg = np.zeros(merkle_num_u64(log_n + rate_bits, cap_height), dtype=np.uint64)
    lib().orc_commit_values(_p(v), _sz(cols), log_n, rate_bits, cap_height, _p(co), _p(lde), _p(dg))
    return co, lde, dg, dg[-(4 << cap_height):].reshape(-1, 4)
