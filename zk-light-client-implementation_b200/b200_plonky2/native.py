"""ctypes binding of libb200plonky2.so (the C ABI in include/b200_plonky2.h).

The CUDA library is the only implementation: importing this module without the built .so, or creating a
Context without a CUDA device, raises -- there is no CPU fallback.
"""
import ctypes, os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(os.path.dirname(_HERE), "libb200plonky2.so")

class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libb200plonky2 error {code}: {msg}")
        self.code = code

def _load():
    if not os.path.exists(_LIB_PATH):
        raise ImportError(f"{_LIB_PATH} not built: run zk-light-client-implementation_b200/build.sh "
                          "(or __graft_entry__.build()); there is no CPU fallback")
    return ctypes.CDLL(_LIB_PATH)

lib = _load()
u64p = ctypes.POINTER(ctypes.c_uint64)
c_void_p, c_size_t, c_int, c_u32, c_u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64

class Common(ctypes.Structure):
    _fields_ = [(n, c_u32) for n in (
        "num_wires", "num_routed_wires", "num_constants", "num_challenges", "degree_bits", "rate_bits", "cap_height",
        "proof_of_work_bits", "num_query_rounds", "quotient_degree_factor", "num_partial_products",
        "num_gate_constraints", "num_public_inputs", "hasher", "num_reduction_arity_bits")] + [
        ("reduction_arity_bits", c_u32 * 16), ("num_gates", c_u32),
        ("gate_ids", ctypes.POINTER(ctypes.c_char_p)), ("selector_indices", ctypes.POINTER(c_u32)),
        ("num_selector_groups", c_u32), ("group_starts", ctypes.POINTER(c_u32)), ("group_ends", ctypes.POINTER(c_u32)),
        ("k_is", u64p)]

class VerifierOnly(ctypes.Structure):
    _fields_ = [("constants_sigmas_cap", u64p), ("circuit_digest", c_u64 * 4)]

def _sig(name, restype, *argtypes):
    f = getattr(lib, name); f.restype = restype; f.argtypes = list(argtypes); return f

_sig("b200_ctx_create", c_int, c_int, ctypes.POINTER(c_void_p))
_sig("b200_ctx_destroy", None, c_void_p)
_sig("b200_ctx_set_stream", c_int, c_void_p, c_void_p)
_sig("b200_ctx_synchronize", c_int, c_void_p)
_sig("b200_last_error", ctypes.c_char_p, c_void_p)
_sig("b200_free", None, c_void_p)
_sig("b200_kernel_launches", c_u64, c_void_p)
_sig("b200_host_alloc", c_int, c_void_p, c_size_t, c_void_p)
_sig("b200_host_free", None, c_void_p)
_sig("b200_merkle_num_u64", c_size_t, c_u32, c_u32)
_sig("b200_ctx_enable_timing", c_int, c_void_p, c_int)
_sig("b200_ctx_timings", ctypes.c_char_p, c_void_p)
_sig("b200_ntt_batch", c_int, c_void_p, c_void_p, c_size_t, c_u32, c_int)
_sig("b200_ntt_batch_dev", c_int, c_void_p, c_void_p, c_size_t, c_u32, c_int, c_void_p)
_sig("b200_lde_commit", c_int, c_void_p, c_void_p, c_size_t, c_u32, c_u32, c_u32, c_void_p, c_void_p, c_void_p)
_sig("b200_lde_commit_dev", c_int, c_void_p, c_void_p, c_size_t, c_u32, c_u32, c_u32, c_void_p, c_void_p, c_void_p)
_sig("b200_poseidon_merkle", c_int, c_void_p, c_void_p, c_size_t, c_u32, c_u32, c_void_p, c_void_p)
_sig("b200_poseidon_merkle_dev", c_int, c_void_p, c_void_p, c_size_t, c_u32, c_u32, c_void_p)
_sig("b200_poseidon_permute", c_int, c_void_p, c_void_p, c_size_t)
_sig("b200_fri_fold", c_int, c_void_p, c_void_p, c_u32, c_u32, c_void_p, c_u64, c_void_p)
_sig("b200_msm_g1", c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p)
_sig("b200_msm_g1_dev", c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p)
_sig("b200_msm_g2", c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p)
_sig("b200_msm_g2_dev", c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p)
_sig("b200_fr_ntt", c_int, c_void_p, c_void_p, c_u32, c_int, c_int, c_void_p)
_sig("b200_fr_ntt_dev", c_int, c_void_p, c_void_p, c_u32, c_int, c_int, c_void_p)
_sig("b200_circuit_build", c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p)
_sig("b200_circuit_digest", c_int, c_u32, c_void_p, c_u32, c_u32, c_void_p)
_sig("b200_ctx_cache_stats", c_int, c_void_p, c_void_p, c_void_p, c_void_p)
_sig("b200_witness_fill_u32", c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p)
_sig("b200_witness_fill_u32_dev", c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p)
class Groth16PkDesc(ctypes.Structure):
    _fields_ = [("log_domain", c_u32), ("num_wires", c_u64), ("num_public", c_u64), ("g1_a", c_void_p), ("g1_b", c_void_p), ("g2_b", c_void_p),
                ("g1_k", c_void_p), ("g1_z", c_void_p), ("g1_alpha", c_u64 * 8), ("g1_beta", c_u64 * 8), ("g1_delta", c_u64 * 8),
                ("g2_beta", c_u64 * 16), ("g2_delta", c_u64 * 16)]
_sig("b200_groth16_pk_upload", c_int, c_void_p, c_void_p, ctypes.POINTER(c_void_p))
_sig("b200_groth16_pk_free", None, c_void_p)
_sig("b200_groth16_prove", c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p)
_sig("b200_verify", c_int, c_void_p, c_void_p, c_void_p, c_size_t)
_sig("b200_verify_error", ctypes.c_char_p)

EXPORTS = ["b200_ctx_create", "b200_ctx_destroy", "b200_ctx_set_stream", "b200_ctx_synchronize", "b200_last_error",
           "b200_free", "b200_host_alloc", "b200_host_free", "b200_kernel_launches", "b200_ctx_enable_timing", "b200_ctx_timings", "b200_circuit_upload", "b200_circuit_build", "b200_circuit_digest", "b200_ctx_cache_stats", "b200_circuit_free", "b200_witness_fill_u32", "b200_witness_fill_u32_dev", "b200_prove",
           "b200_prove_dev", "b200_ntt_batch", "b200_ntt_batch_dev", "b200_lde_commit", "b200_lde_commit_dev",
           "b200_poseidon_merkle", "b200_poseidon_merkle_dev", "b200_merkle_num_u64", "b200_fri_fold",
           "b200_poseidon_permute", "b200_verify", "b200_verify_error", "b200_msm_g1", "b200_msm_g1_dev", "b200_msm_g2", "b200_msm_g2_dev", "b200_fr_ntt", "b200_fr_ntt_dev", "b200_groth16_pk_upload", "b200_groth16_pk_free", "b200_groth16_prove"]

def _np(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    return a, a.ctypes.data_as(c_void_p)

class Context:
    """One per GPU / stream (b200_ctx)."""
    def __init__(self, device=0):
        h = c_void_p()
        rc = lib.b200_ctx_create(device, ctypes.byref(h))
        if rc != 0:
            raise B200Error(rc, "b200_ctx_create failed (no CUDA device?) -- no CPU fallback exists")
        self.h = h
    def close(self):
        if self.h:
            lib.b200_ctx_destroy(self.h); self.h = None
    def __del__(self):
        try: self.close()
        except Exception: pass
    def _chk(self, rc):
        if rc != 0:
            raise B200Error(rc, lib.b200_last_error(self.h).decode())
    def set_stream(self, cuda_stream_ptr):
        self._chk(lib.b200_ctx_set_stream(self.h, c_void_p(cuda_stream_ptr)))
    def synchronize(self):
        self._chk(lib.b200_ctx_synchronize(self.h))
    def kernel_launches(self):
        return int(lib.b200_kernel_launches(self.h))
    def pinned_empty(self, shape, dtype=np.uint64):
        """numpy array over page-locked memory from b200_host_alloc (freed when the array and its views are collected)."""
        import weakref
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = c_void_p()
        self._chk(lib.b200_host_alloc(self.h, nbytes, ctypes.byref(p)))
        buf = (ctypes.c_uint8 * max(nbytes, 1)).from_address(p.value)
        weakref.finalize(buf, lib.b200_host_free, c_void_p(p.value))
        return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    def pinned_copy(self, a):
        out = self.pinned_empty(a.shape, a.dtype); out[...] = a
        return out
    def enable_timing(self, on=True):
        self._chk(lib.b200_ctx_enable_timing(self.h, int(on)))
    def cache_stats(self):
        h, m, e = c_u64(), c_u64(), c_u64()
        lib.b200_ctx_cache_stats(self.h, ctypes.byref(h), ctypes.byref(m), ctypes.byref(e))
        return {"hits": h.value, "misses": m.value, "entries": e.value}
    def timings(self):
        import json
        return json.loads(lib.b200_ctx_timings(self.h).decode())

    # ---- host-buffer entry points ----
    def ntt_batch(self, values, log_n, inverse=False):
        """values: [cols, 2^log_n] uint64 (natural order) -> transformed copy."""
        a, p = _np(np.array(values, dtype=np.uint64, copy=True))
        cols = a.shape[0]
        self._chk(lib.b200_ntt_batch(self.h, p, cols, log_n, int(inverse)))
        return a
    def lde_commit(self, values, log_n, rate_bits=3, cap_height=4, want_lde=True, want_coeffs=True):
        a, p = _np(values)
        cols = a.shape[0]; n = 1 << log_n
        cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
        lde = np.zeros((cols, n << rate_bits), dtype=np.uint64) if want_lde else None
        co = np.zeros((cols, n), dtype=np.uint64) if want_coeffs else None
        self._chk(lib.b200_lde_commit(self.h, p, cols, log_n, rate_bits, cap_height, cap.ctypes.data_as(c_void_p),
                                      lde.ctypes.data_as(c_void_p) if want_lde else None,
                                      co.ctypes.data_as(c_void_p) if want_coeffs else None))
        return cap, lde, co
    def poseidon_merkle(self, data, log_L, cap_height=4, want_digests=False):
        a, p = _np(data)
        cols = a.shape[0]
        cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
        dg = np.zeros(lib.b200_merkle_num_u64(log_L, cap_height), dtype=np.uint64) if want_digests else None
        self._chk(lib.b200_poseidon_merkle(self.h, p, cols, log_L, cap_height, cap.ctypes.data_as(c_void_p),
                                           dg.ctypes.data_as(c_void_p) if want_digests else None))
        return (cap, dg) if want_digests else cap
    def msm_g1(self, points, scalars):
        """points [n][8] u64 (affine x,y limbs), scalars [n][4] u64 -> (x, y) Python ints, or None for infinity."""
        p, pp = _np(points); s, sp = _np(scalars)
        out = np.zeros(8, dtype=np.uint64)
        self._chk(lib.b200_msm_g1(self.h, pp, sp, p.shape[0], out.ctypes.data_as(c_void_p)))
        x = sum(int(out[i]) << (64 * i) for i in range(4)); y = sum(int(out[4 + i]) << (64 * i) for i in range(4))
        return None if x == 0 and y == 0 else (x, y)
    def msm_g2(self, points, scalars):
        """points [n][16] u64 (x.A0, x.A1, y.A0, y.A1 limbs), scalars [n][4] u64 -> ((x0, x1), (y0, y1)) or None for infinity."""
        p, pp = _np(points); s, sp = _np(scalars)
        out = np.zeros(16, dtype=np.uint64)
        self._chk(lib.b200_msm_g2(self.h, pp, sp, p.shape[0], out.ctypes.data_as(c_void_p)))
        v = [sum(int(out[4 * k + i]) << (64 * i) for i in range(4)) for k in range(4)]
        return None if not any(v) else ((v[0], v[1]), (v[2], v[3]))
    def groth16_pk_upload(self, log_domain, num_public, g1_a, g1_b, g2_b, g1_k, g1_z, g1_alpha, g1_beta, g1_delta, g2_beta, g2_delta):
        """arrays of affine points as u64 limb rows ([n][8] for G1, [n][16] for G2); returns an opaque handle (free with groth16_pk_free)"""
        arrs = [np.ascontiguousarray(x, dtype=np.uint64) for x in (g1_a, g1_b, g2_b, g1_k, g1_z)]
        d = Groth16PkDesc(); d.log_domain = log_domain; d.num_wires = arrs[0].shape[0]; d.num_public = num_public
        d.g1_a, d.g1_b, d.g2_b, d.g1_k, d.g1_z = [a.ctypes.data_as(c_void_p) if a.size else None for a in arrs]
        for name, v, n in (("g1_alpha", g1_alpha, 8), ("g1_beta", g1_beta, 8), ("g1_delta", g1_delta, 8), ("g2_beta", g2_beta, 16), ("g2_delta", g2_delta, 16)):
            for i in range(n): getattr(d, name)[i] = int(v[i])
        h = c_void_p()
        self._chk(lib.b200_groth16_pk_upload(self.h, ctypes.byref(d), ctypes.byref(h)))
        return h
    def groth16_pk_free(self, h): lib.b200_groth16_pk_free(h)
    def groth16_prove(self, pk, wires, a, b, c, r, s):
        """wires [num_wires][4], a/b/c [domain][4] u64 limbs, r/s ints -> 32 u64 limbs: Ar (8) | Bs in G2 (16) | Krs (8)"""
        W, pw = _np(wires); A, pa = _np(a); B, pb = _np(b); C, pc = _np(c)
        lim = lambda x: np.array([(int(x) >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
        R, S = lim(r), lim(s)
        out = np.zeros(32, dtype=np.uint64)
        self._chk(lib.b200_groth16_prove(self.h, pk, pw, pa, pb, pc, R.ctypes.data_as(c_void_p), S.ctypes.data_as(c_void_p), out.ctypes.data_as(c_void_p)))
        return out
    def fr_ntt(self, values, inverse=False, coset=False):
        """values: [n][4] u64 canonical limbs of BN254 F_r elements (natural order) -> transformed [n][4]."""
        a, p = _np(values)
        n = a.shape[0]; log_n = n.bit_length() - 1
        assert n == 1 << log_n and a.shape[1] == 4
        out = np.zeros_like(a)
        self._chk(lib.b200_fr_ntt(self.h, p, log_n, int(inverse), int(coset), out.ctypes.data_as(c_void_p)))
        return out
    def poseidon_permute(self, states):
        a, p = _np(np.array(states, dtype=np.uint64, copy=True))
        self._chk(lib.b200_poseidon_permute(self.h, p, a.shape[0]))
        return a


def common_struct(common: dict):
    """dict with the common_data.json fields (+ "hasher": "gl" | "bn254") -> (b200_common, keep-alive objects)."""
    c = common
    cc = Common()
    for k in ("num_wires", "num_routed_wires", "num_constants", "num_challenges", "degree_bits", "rate_bits", "cap_height",
              "proof_of_work_bits", "num_query_rounds", "quotient_degree_factor", "num_partial_products",
              "num_gate_constraints", "num_public_inputs"):
        setattr(cc, k, c[k])
    cc.hasher = 1 if c.get("hasher", "gl") == "bn254" else 0
    cc.num_reduction_arity_bits = len(c["reduction_arity_bits"])
    for i, b in enumerate(c["reduction_arity_bits"]): cc.reduction_arity_bits[i] = b
    cc.num_gates = len(c["gates"])
    keep = dict(ids=(ctypes.c_char_p * len(c["gates"]))(*[g.encode() for g in c["gates"]]),
                sel=(c_u32 * len(c["gates"]))(*c["selector_indices"]),
                gs=(c_u32 * len(c["groups"]))(*[s for s, e in c["groups"]]), ge=(c_u32 * len(c["groups"]))(*[e for s, e in c["groups"]]),
                kis=np.array(c["k_is"], dtype=np.uint64))
    cc.gate_ids = ctypes.cast(keep["ids"], ctypes.POINTER(ctypes.c_char_p))
    cc.selector_indices = ctypes.cast(keep["sel"], ctypes.POINTER(c_u32))
    cc.num_selector_groups = len(c["groups"])
    cc.group_starts = ctypes.cast(keep["gs"], ctypes.POINTER(c_u32)); cc.group_ends = ctypes.cast(keep["ge"], ctypes.POINTER(c_u32))
    cc.k_is = keep["kis"].ctypes.data_as(u64p)
    return cc, keep

def common_from_json(d: dict, hasher="gl"):
    """plonky2 common_data.json (as written by near_bft_finality/src/bin/prove_block.rs:343-356) -> the dict used here."""
    cfg, fp = d["config"], d["fri_params"]
    return dict(num_wires=cfg["num_wires"], num_routed_wires=cfg["num_routed_wires"], num_constants=d["num_constants"],
                num_challenges=cfg["num_challenges"], degree_bits=fp["degree_bits"], rate_bits=fp["config"]["rate_bits"],
                cap_height=fp["config"]["cap_height"], proof_of_work_bits=fp["config"]["proof_of_work_bits"],
                num_query_rounds=fp["config"]["num_query_rounds"], quotient_degree_factor=d["quotient_degree_factor"],
                num_partial_products=d["num_partial_products"], num_gate_constraints=d["num_gate_constraints"],
                num_public_inputs=d["num_public_inputs"], reduction_arity_bits=list(fp["reduction_arity_bits"]),
                gates=list(d["gates"]), selector_indices=list(d["selectors_info"]["selector_indices"]),
                groups=[(g["start"], g["end"]) for g in d["selectors_info"]["groups"]], k_is=list(d["k_is"]), hasher=hasher)

def circuit_digest(hasher, cap, degree_bits):
    """CircuitBuilder::build's circuit_digest from the constants/sigmas cap ([2^cap_height][4] u64), host only."""
    cap = np.ascontiguousarray(cap, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros(4, dtype=np.uint64)
    ch = int(cap.shape[0]).bit_length() - 1
    assert cap.shape[0] == 1 << ch
    rc = lib.b200_circuit_digest(1 if hasher == "bn254" else 0, cap.ctypes.data_as(c_void_p), ch, degree_bits, out.ctypes.data_as(c_void_p))
    if rc: raise B200Error(rc, "b200_circuit_digest: bad arguments")
    return [int(x) for x in out]

def verify(common: dict, constants_sigmas_cap, circuit_digest, proof_bytes: bytes):
    """CircuitData::verify: constants_sigmas_cap [2^cap_height][4] u64 words, circuit_digest 4 words.  Returns (ok, reason)."""
    cc, keep = common_struct(common)
    cap = np.ascontiguousarray(constants_sigmas_cap, dtype=np.uint64)
    vo = VerifierOnly(); vo.constants_sigmas_cap = cap.ctypes.data_as(u64p)
    for i in range(4): vo.circuit_digest[i] = int(circuit_digest[i])
    buf = (ctypes.c_uint8 * len(proof_bytes)).from_buffer_copy(proof_bytes)
    rc = lib.b200_verify(ctypes.byref(cc), ctypes.byref(vo), buf, len(proof_bytes))
    return rc == 0, lib.b200_verify_error().decode()
