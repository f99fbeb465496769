"""Python front-end of the C oracle prover (oracle/c/prover.c).  TEST INFRASTRUCTURE.

circuit_digest restates upstream CircuitBuilder::build (hash_no_pad(cap.flatten() || hash_pad(domain_separator=[])
|| degree_bits)); the formula is pinned against near_bft_finality/proofs/*/verifier_data.json with the BN254 hasher
in tests/test_oracle_prover.py."""
import ctypes
import numpy as np
from . import cbind, cgates, poseidon_gl as pg, poseidon_bn254 as pb, proof_io as io

class OrcCommon(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "num_wires", "num_routed", "num_constants", "num_challenges", "degree_bits", "rate_bits", "cap_height", "pow_bits",
        "num_queries", "qdf", "npp", "num_gate_constraints", "num_pi", "num_selectors", "n_arity")] + [
        ("arity_bits", ctypes.c_uint32 * 16), ("num_gates", ctypes.c_uint32), ("gates", ctypes.c_void_p), ("k_is", ctypes.c_void_p),
        ("hasher", ctypes.c_uint32)]

class OrcDebug(ctypes.Structure):
    _fields_ = [("betas", ctypes.c_uint64 * 4), ("gammas", ctypes.c_uint64 * 4), ("alphas", ctypes.c_uint64 * 4),
                ("zeta", ctypes.c_uint64 * 2), ("fri_alpha", ctypes.c_uint64 * 2), ("fri_betas", (ctypes.c_uint64 * 2) * 16),
                ("zs_pp_values", ctypes.c_void_p), ("quotient_coeffs", ctypes.c_void_p), ("fri_values0", ctypes.c_void_p)]

def hash_pad_gl(x):
    x = list(x) + [1]
    while (len(x) + 1) % 8 != 0: x.append(0)
    x.append(1)
    return pg.hash_no_pad(x)

def circuit_digest_gl(cap, degree_bits):
    parts = [int(v) for v in np.asarray(cap).reshape(-1)] + hash_pad_gl([]) + [degree_bits]
    return pg.hash_no_pad(parts)

def circuit_digest_bn254(cap, degree_bits):
    x = [1]
    while (len(x) + 1) % 9 != 0: x.append(0)
    x.append(1)
    parts = []
    for h in cap: parts += pb.to_vec(h)
    parts += pb.to_vec(pb.hash_no_pad(x)) + [degree_bits]
    return pb.hash_no_pad(parts)

def common_data(common: dict) -> io.CommonData:
    c = common
    return io.CommonData(
        num_wires=c["num_wires"], num_routed_wires=c["num_routed_wires"], num_constants_cfg=c.get("num_constants_cfg", 2),
        num_challenges=c["num_challenges"],
        fri=io.FriConfig(c["rate_bits"], c["cap_height"], c["proof_of_work_bits"], c["num_query_rounds"]),
        degree_bits=c["degree_bits"], reduction_arity_bits=list(c["reduction_arity_bits"]), gates=list(c["gates"]),
        selector_indices=list(c["selector_indices"]), groups=[tuple(g) for g in c["groups"]],
        quotient_degree_factor=c["quotient_degree_factor"], num_gate_constraints=c["num_gate_constraints"],
        num_constants=c["num_constants"], num_public_inputs=c["num_public_inputs"], k_is=list(c["k_is"]),
        num_partial_products=c["num_partial_products"], hasher=c.get("hasher", "gl"))

class OracleCircuit:
    """build()'s prover-side half on the CPU: constants/sigmas commitment + digest."""
    def __init__(self, common: dict, constants_sigmas):
        self.lib = cbind.lib()
        self.common = common
        self.hasher = common.get("hasher", "gl")
        c = common
        self.gate_arr, self._keep = cgates.make_gate_array(c["gates"], c["selector_indices"], c["groups"])
        self.kis = np.array(c["k_is"], dtype=np.uint64)
        oc = OrcCommon()
        oc.num_wires = c["num_wires"]; oc.num_routed = c["num_routed_wires"]; oc.num_constants = c["num_constants"]
        oc.num_challenges = c["num_challenges"]; oc.degree_bits = c["degree_bits"]; oc.rate_bits = c["rate_bits"]
        oc.cap_height = c["cap_height"]; oc.pow_bits = c["proof_of_work_bits"]; oc.num_queries = c["num_query_rounds"]
        oc.qdf = c["quotient_degree_factor"]; oc.npp = c["num_partial_products"]; oc.num_gate_constraints = c["num_gate_constraints"]
        oc.num_pi = c["num_public_inputs"]; oc.num_selectors = len(c["groups"]); oc.n_arity = len(c["reduction_arity_bits"])
        for i, b in enumerate(c["reduction_arity_bits"]): oc.arity_bits[i] = b
        oc.num_gates = len(c["gates"]); oc.gates = ctypes.addressof(self.gate_arr); oc.k_is = self.kis.ctypes.data
        oc.hasher = 1 if self.hasher == "bn254" else 0
        self.oc = oc
        cs = np.ascontiguousarray(constants_sigmas, dtype=np.uint64)
        cap = np.zeros((1 << c["cap_height"], 4), dtype=np.uint64)
        zero = (ctypes.c_uint64 * 4)(0, 0, 0, 0)
        self.lib.orc_circuit_new.restype = ctypes.c_void_p
        self.h = ctypes.c_void_p(self.lib.orc_circuit_new(ctypes.byref(oc), cs.ctypes.data_as(ctypes.c_void_p), zero,
                                                          cap.ctypes.data_as(ctypes.c_void_p)))
        self.cap = cap
        if self.hasher == "bn254":
            self.cap_ints = [sum(int(h[i]) << (64 * i) for i in range(4)) for h in cap]
            d = circuit_digest_bn254(self.cap_ints, c["degree_bits"])
            self.digest_int = d
            self.digest = [(d >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
        else:
            self.digest = circuit_digest_gl(cap, c["degree_bits"])
        self.lib.orc_circuit_set_digest(self.h, (ctypes.c_uint64 * 4)(*self.digest))

    def verifier_only(self):
        if self.hasher == "bn254":
            return io.VerifierOnly(list(self.cap_ints), self.digest_int)
        return io.VerifierOnly([tuple(int(x) for x in h) for h in self.cap], tuple(self.digest))

    def prove(self, wires, public_inputs, pow_hint=0, debug=False):
        c = self.common
        w = np.ascontiguousarray(wires, dtype=np.uint64); pis = np.array(public_inputs, dtype=np.uint64)
        out = ctypes.POINTER(ctypes.c_uint8)(); ln = ctypes.c_size_t()
        dbg = None; keep = None
        if debug:
            n = 1 << c["degree_bits"]; nzp = c["num_challenges"] * (1 + c["num_partial_products"])
            keep = dict(zs=np.zeros((nzp, n), np.uint64), q=np.zeros((c["num_challenges"] * c["quotient_degree_factor"], n), np.uint64),
                        f0=np.zeros((n << c["rate_bits"], 2), np.uint64))
            dbg = OrcDebug(); dbg.zs_pp_values = keep["zs"].ctypes.data; dbg.quotient_coeffs = keep["q"].ctypes.data
            dbg.fri_values0 = keep["f0"].ctypes.data
        rc = self.lib.orc_prove(self.h, w.ctypes.data_as(ctypes.c_void_p), pis.ctypes.data_as(ctypes.c_void_p),
                                ctypes.c_uint64(pow_hint), ctypes.byref(out), ctypes.byref(ln), ctypes.byref(dbg) if dbg else None)
        if rc != 0: raise RuntimeError(f"orc_prove failed: {rc}")
        b = ctypes.string_at(out, ln.value)
        self.lib.orc_free(out)
        if debug:
            keep.update(betas=list(dbg.betas)[:2], gammas=list(dbg.gammas)[:2], alphas=list(dbg.alphas)[:2], zeta=list(dbg.zeta),
                        fri_alpha=list(dbg.fri_alpha))
            return b, keep
        return b

    def close(self):
        if self.h: self.lib.orc_circuit_free(self.h); self.h = None
    def __del__(self):
        try: self.close()
        except Exception: pass
