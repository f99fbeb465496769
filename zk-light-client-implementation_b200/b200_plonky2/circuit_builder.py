"""Host-side mirror of the reference's plonky2 surface for the hot path: a small vectorised CircuitBuilder whose
build() produces exactly what plonky2's CircuitBuilder::build::<C>() hands to the prover (selector / constant /
sigma polynomial VALUES, k_is, gate table, selector groups) and whose CircuitData.prove() calls the CUDA prover
through the C ABI (reference call sites: near_bft_finality/src/prove_crypto/ed25519.rs:33,60; recursion.rs:94-95;
sha256.rs:78,82).  It is NOT a port of the reference's gadget library: the real ed25519/sha512 gadgets stay in
Rust above the C ABI.  Here gates are filled block-wise with synthetic satisfying witnesses (same gate types, wire
layouts and copy-constraint structure as the reference circuits) so that the prover can be exercised, checked by the
oracle verifier and benchmarked without a Rust toolchain.

Semantics mirrored from upstream plonky2 (wormhole-foundation/plonky2-near @ 2244a9d, not vendored):
  * gates sorted by (degree, id); selector groups built greedily with max_degree = quotient_degree_factor + 1
    (reproduces selectors_info of near_bft_finality/proofs/random/CGZP.../common_data.json)
  * constant columns = selector columns followed by gate constants; unused selector = 2^32 - 1
  * sigma polynomials: value k_is[col'] * w^row' of the next wire in the copy class; k_is[j] = 7^j
  * public inputs are hashed in-circuit with PoseidonGate rows and the digest is wired to PublicInputGate (row 0)
"""
import ctypes
import numpy as np
from . import field_np as F
from .field_np import P

U32 = np.uint64(0xFFFFFFFF)

class CircuitConfig:
    def __init__(self, num_wires, num_routed_wires=80, num_constants=2, num_challenges=2, rate_bits=3, cap_height=4,
                 proof_of_work_bits=16, num_query_rounds=28, max_quotient_degree_factor=8):
        self.num_wires = num_wires; self.num_routed_wires = num_routed_wires; self.num_constants = num_constants
        self.num_challenges = num_challenges; self.rate_bits = rate_bits; self.cap_height = cap_height
        self.proof_of_work_bits = proof_of_work_bits; self.num_query_rounds = num_query_rounds
        self.max_quotient_degree_factor = max_quotient_degree_factor
    @staticmethod
    def standard_recursion_config(): return CircuitConfig(135)     # near_bft_finality/src/prove_crypto/recursion.rs:36
    @staticmethod
    def wide_ecc_config(): return CircuitConfig(234)               # near_bft_finality/src/prove_crypto/ed25519.rs:29

_PH = "_phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField>"

class GateKind:
    def __init__(self, gid, degree, num_constants, num_wires):
        self.id = gid; self.degree = degree; self.num_constants = num_constants; self.num_wires = num_wires

def noop_gate(): return GateKind("NoopGate", 0, 0, 0)
def constant_gate(n=2): return GateKind(f"ConstantGate {{ num_consts: {n} }}", 1, n, n)
def public_input_gate(): return GateKind("PublicInputGate", 1, 0, 4)
def arithmetic_gate(cfg): n = cfg.num_routed_wires // 4; return GateKind(f"ArithmeticGate {{ num_ops: {n} }}", 3, 2, 4 * n)
def base_sum_gate(limbs=63): return GateKind(f"BaseSumGate {{ num_limbs: {limbs} }} + Base: 2", 2, 0, 1 + limbs)
def poseidon_gate(): return GateKind("PoseidonGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>", 7, 0, 135)
def u32_arithmetic_gate(cfg):
    n = min(cfg.num_wires // 38, cfg.num_routed_wires // 6)
    return GateKind(f"U32ArithmeticGate {{ num_ops: {n}, {_PH} }}", 4, 0, 38 * n)
def u32_add_many_gate(cfg, addends):
    n = min(cfg.num_wires // (addends + 21), cfg.num_routed_wires // (addends + 3))
    return GateKind(f"U32AddManyGate {{ num_addends: {addends}, num_ops: {n}, {_PH} }}", 4, 0, (addends + 21) * n)
def u32_subtraction_gate(cfg):
    n = min(cfg.num_wires // 21, cfg.num_routed_wires // 5)
    return GateKind(f"U32SubtractionGate {{ num_ops: {n}, {_PH} }}", 4, 0, 21 * n)
def u32_range_check_gate(limbs): return GateKind(f"U32RangeCheckGate {{ num_input_limbs: {limbs}, {_PH} }}", 4, 0, 17 * limbs)
def u32_interleave_gate(cfg):
    n = min(cfg.num_wires // 34, cfg.num_routed_wires // 2)
    return GateKind(f"U32InterleaveGate {{ num_ops: {n} }}", 2, 0, 34 * n)
def uninterleave_to_u32_gate(cfg):
    n = min(cfg.num_wires // 67, cfg.num_routed_wires // 3)
    return GateKind(f"UninterleaveToU32Gate {{ num_ops: {n} }}", 2, 0, 67 * n)
def uninterleave_to_b32_gate(cfg):
    n = min(cfg.num_wires // 67, cfg.num_routed_wires // 3)
    return GateKind(f"UninterleaveToB32Gate {{ num_ops: {n} }}", 2, 0, 67 * n)
def comparison_gate(bits=32, chunks=16):
    cb = -(-bits // chunks)
    return GateKind(f"ComparisonGate {{ num_bits: {bits}, num_chunks: {chunks}, {_PH} }}<D=2>", 1 << cb, 0, 4 + 5 * chunks + cb + 1)

def _num_ops(kind):
    import re
    m = re.search(r"num_ops: (\d+)", kind.id)
    return int(m[1])

# ---------------------------------------------------------------------------------------------------------------
class CircuitBuilder:
    def __init__(self, config: CircuitConfig):
        self.cfg = config
        self.blocks = []          # (kind, consts [cnt][k], wires [cnt][num_wires])
        self.num_rows = 0
        self.copies = []          # (rowsA, colsA, rowsB, colsB) arrays
        self.public_inputs = []
        # row 0: PublicInputGate (digest wires are connected by register_public_inputs)
        self._pi_row = self.add_rows(public_input_gate(), np.zeros((1, 0), np.uint64), np.zeros((1, config.num_wires), np.uint64))[0]

    def add_rows(self, kind, consts, wires):
        cnt = wires.shape[0]
        assert wires.shape[1] == self.cfg.num_wires and kind.num_wires <= self.cfg.num_wires, kind.id
        assert consts.shape == (cnt, kind.num_constants), (kind.id, consts.shape)
        self.blocks.append((kind, np.ascontiguousarray(consts, dtype=np.uint64), np.ascontiguousarray(wires, dtype=np.uint64)))
        r0 = self.num_rows
        self.num_rows += cnt
        return np.arange(r0, r0 + cnt)

    def connect(self, rows_a, col_a, rows_b, col_b):
        rows_a = np.atleast_1d(np.asarray(rows_a)); rows_b = np.atleast_1d(np.asarray(rows_b))
        ca = np.broadcast_to(np.asarray(col_a), rows_a.shape); cb = np.broadcast_to(np.asarray(col_b), rows_b.shape)
        assert rows_a.shape == rows_b.shape
        assert (ca < self.cfg.num_routed_wires).all() and (cb < self.cfg.num_routed_wires).all(), "copy constraint on a non-routed wire"
        self.copies.append((rows_a.copy(), ca.copy(), rows_b.copy(), cb.copy()))

    def _block_of_row(self, row):
        r = 0
        for b in self.blocks:
            if row < r + b[2].shape[0]: return b, row - r
            r += b[2].shape[0]
        raise IndexError(row)

    def register_public_inputs(self, values):
        """Hash the public inputs in-circuit (PoseidonGate sponge, rate 8) and wire the digest to row 0."""
        vals = [int(v) % P for v in values]
        self.public_inputs = vals
        state = np.zeros((1, 12), np.uint64)
        prev = None
        for i in range(0, len(vals), 8):
            chunk = vals[i:i + 8]
            inp = state.copy()
            inp[0, :len(chunk)] = np.array(chunk, dtype=np.uint64)
            rows, out = fill_poseidon(self, inp, np.zeros(1, np.uint64))
            if prev is not None:                      # capacity + untouched rate words carry over from the previous permutation
                for j in range(len(chunk), 12):
                    self.connect(prev, 12 + j, rows, j)
            prev = rows; state = out
        pi_hash = state[0, :4].copy()
        kind, consts, wires = self.blocks[0]
        wires[0, :4] = pi_hash
        if prev is not None:
            for j in range(4):
                self.connect(prev, 12 + j, [self._pi_row], j)
        return pi_hash

    # ---------------------------------------------------------------------------------------------------------
    def build(self):
        cfg = self.cfg
        deg = self.num_rows
        degree_bits = max(2, int(np.ceil(np.log2(max(deg, 2)))))
        n = 1 << degree_bits
        if n > deg:
            self.add_rows(noop_gate(), np.zeros((n - deg, 0), np.uint64), np.zeros((n - deg, cfg.num_wires), np.uint64))
        # gate table sorted by (degree, id)
        kinds = {}
        for kind, _, _ in self.blocks: kinds.setdefault(kind.id, kind)
        gates = sorted(kinds.values(), key=lambda g: (g.degree, g.id))
        index = {g.id: i for i, g in enumerate(gates)}
        qdf = cfg.max_quotient_degree_factor
        max_degree = qdf + 1
        assert gates[-1].degree <= qdf, "gate degree too high for the quotient degree factor"
        if gates[-1].degree + len(gates) - 1 <= max_degree:
            groups = [(0, len(gates))]
        else:
            groups, start = [], 0
            while start < len(gates):
                size = 0
                while start + size < len(gates) and size + gates[start + size].degree < max_degree: size += 1
                groups.append((start, start + size)); start += size
        sel_index = [next(k for k, (s, e) in enumerate(groups) if s <= i < e) for i in range(len(gates))]
        nsel = len(groups)
        ngc = max([g.num_constants for g in gates] + [cfg.num_constants])
        num_constants = nsel + ngc
        consts = np.zeros((num_constants, n), np.uint64)
        consts[:nsel, :] = np.uint64(0xFFFFFFFF) if nsel > 1 else np.uint64(0)
        wires = np.zeros((cfg.num_wires, n), np.uint64)
        r = 0
        for kind, c, w in self.blocks:
            cnt = w.shape[0]
            gi = index[kind.id]
            consts[sel_index[gi], r:r + cnt] = gi
            if kind.num_constants: consts[nsel:nsel + kind.num_constants, r:r + cnt] = c.T
            wires[:, r:r + cnt] = w.T
            r += cnt
        # ---- copy constraints -> sigma ----
        R = cfg.num_routed_wires
        nslots = R * n                                      # slot = col * n + row
        if self.copies:
            a = np.concatenate([c[1] * n + c[0] for c in self.copies]).astype(np.int64)
            b = np.concatenate([c[3] * n + c[2] for c in self.copies]).astype(np.int64)
        else:
            a = b = np.zeros(0, np.int64)
        flat = wires[:R].reshape(-1)
        if len(a) and not np.array_equal(flat[a], flat[b]):
            bad = np.nonzero(flat[a] != flat[b])[0][:5]
            raise ValueError(f"copy constraint between unequal wires, e.g. slots {[(int(a[i]), int(b[i])) for i in bad]}")
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components
        g = coo_matrix((np.ones(len(a), np.int8), (a, b)), shape=(nslots, nslots))
        _, labels = connected_components(g, directed=False)
        order = np.argsort(labels, kind="stable")
        lab_sorted = labels[order]
        nxt = np.empty(nslots, np.int64)
        starts = np.nonzero(np.r_[True, lab_sorted[1:] != lab_sorted[:-1]])[0]
        ends = np.r_[starts[1:], nslots]
        rolled = np.roll(order, -1)
        rolled[ends - 1] = order[starts]                  # close each cycle
        nxt[order] = rolled
        k_is = F.powers(7, R)
        w = F.root_of_unity(degree_bits)
        subgroup = F.powers(w, n)
        sig_col = nxt // n; sig_row = nxt % n
        sigmas = F.mul(k_is[sig_col], subgroup[sig_row]).reshape(R, n)
        cs = np.ascontiguousarray(np.concatenate([consts, sigmas], axis=0))
        npp = -(-R // qdf) - 1
        # reduction arity bits: ConstantArityBits(4, 5)
        rab, db = [], degree_bits
        while db > 5 and db + cfg.rate_bits - 4 >= cfg.cap_height:
            rab.append(4); db -= 4
        common = dict(
            num_wires=cfg.num_wires, num_routed_wires=R, num_constants_cfg=cfg.num_constants, num_constants=num_constants,
            num_challenges=cfg.num_challenges, degree_bits=degree_bits, rate_bits=cfg.rate_bits, cap_height=cfg.cap_height,
            proof_of_work_bits=cfg.proof_of_work_bits, num_query_rounds=cfg.num_query_rounds, quotient_degree_factor=qdf,
            num_partial_products=npp, num_gate_constraints=max(_gate_num_constraints(g.id) for g in gates),
            num_public_inputs=len(self.public_inputs), reduction_arity_bits=rab, gates=[g.id for g in gates],
            selector_indices=sel_index, groups=groups, k_is=[int(x) for x in k_is])
        cd = CircuitData(common, cs, wires, list(self.public_inputs))
        # what CircuitBuilder::build hands to b200_circuit_build: gate placement, the gates' constants, the copy-class successor map
        row_gate = np.zeros(n, np.uint32); r = 0
        for kind, c, w_ in self.blocks:
            row_gate[r:r + w_.shape[0]] = index[kind.id]; r += w_.shape[0]
        cd.build_inputs = dict(row_gate=row_gate, gate_constants=np.ascontiguousarray(consts[nsel:]), sigma_next=nxt.astype(np.uint32).reshape(R, n))
        return cd

def _gate_num_constraints(gid):
    # counted by the gate definitions (crypto/plonky2_u32/src/gates/*.rs num_constraints(); plonky2 standard gates)
    import re
    g = lambda pat: [int(x) for x in re.search(pat, gid).groups()]
    if gid.startswith("NoopGate"): return 0
    if gid.startswith("ConstantGate"): return g(r"num_consts: (\d+)")[0]
    if gid.startswith("PublicInputGate"): return 4
    if gid.startswith("ArithmeticGate"): return g(r"num_ops: (\d+)")[0]
    if gid.startswith("BaseSumGate"): return 1 + g(r"num_limbs: (\d+)")[0]
    if gid.startswith("PoseidonGate"): return 123
    if gid.startswith("PoseidonMdsGate"): return 24
    if gid.startswith("ArithmeticExtensionGate"): return 2 * g(r"num_ops: (\d+)")[0]
    if gid.startswith("MulExtensionGate"): return 2 * g(r"num_ops: (\d+)")[0]
    if gid.startswith("ReducingGate") or gid.startswith("ReducingExtensionGate"): return 2 * g(r"num_coeffs: (\d+)")[0]
    if gid.startswith("RandomAccessGate"):
        b_, c_, e_ = g(r"bits: (\d+), num_copies: (\d+), num_extra_constants: (\d+)"); return c_ * (b_ + 2) + e_
    if gid.startswith("CosetInterpolationGate"):
        sb, dg = g(r"subgroup_bits: (\d+), degree: (\d+)"); return 2 + 4 * (((1 << sb) - 2) // (dg - 1)) + 2
    if gid.startswith("U32ArithmeticGate"): return 36 * g(r"num_ops: (\d+)")[0]
    if gid.startswith("U32AddManyGate"): return 21 * g(r"num_ops: (\d+)")[0]
    if gid.startswith("U32SubtractionGate"): return 19 * g(r"num_ops: (\d+)")[0]
    if gid.startswith("U32RangeCheckGate"): return 17 * g(r"num_input_limbs: (\d+)")[0]
    if gid.startswith("U32InterleaveGate"): return 34 * g(r"num_ops: (\d+)")[0]
    if gid.startswith("UninterleaveTo"): return 67 * g(r"num_ops: (\d+)")[0]
    if gid.startswith("ComparisonGate"):
        b, c = g(r"num_bits: (\d+), num_chunks: (\d+)"); return 6 + 5 * c + (-(-b // c))
    raise ValueError(gid)

# ---------------------------------------------------------------------------------------------------------------
class CircuitData:
    """What build() returns: prover-side data.  prove() mirrors plonky2 CircuitData::prove(pw) -> proof bytes."""
    build_inputs = None
    def __init__(self, common, constants_sigmas, wires, public_inputs):
        self.common = common
        self.constants_sigmas = constants_sigmas      # [num_constants + routed][n] values on H
        self.wires = wires                            # [num_wires][n] (the witness this circuit was built with)
        self.public_inputs = public_inputs
        self._handle = None; self._ctx = None
        self.constants_sigmas_cap = None
        self.circuit_digest = None

    @property
    def n(self): return 1 << self.common["degree_bits"]

    def c_common(self):
        from . import Common, c_u32, u64p
        c = self.common
        cc = Common()
        for k in ("num_wires", "num_routed_wires", "num_constants", "num_challenges", "degree_bits", "rate_bits", "cap_height",
                  "proof_of_work_bits", "num_query_rounds", "quotient_degree_factor", "num_partial_products",
                  "num_gate_constraints", "num_public_inputs"):
            setattr(cc, k, c[k])
        cc.hasher = 1 if c.get("hasher", "gl") == "bn254" else 0
        cc.num_reduction_arity_bits = len(c["reduction_arity_bits"])
        for i, b in enumerate(c["reduction_arity_bits"]): cc.reduction_arity_bits[i] = b
        cc.num_gates = len(c["gates"])
        self._keep = dict(
            ids=(ctypes.c_char_p * len(c["gates"]))(*[g.encode() for g in c["gates"]]),
            sel=(c_u32 * len(c["gates"]))(*c["selector_indices"]),
            gs=(c_u32 * len(c["groups"]))(*[s for s, e in c["groups"]]),
            ge=(c_u32 * len(c["groups"]))(*[e for s, e in c["groups"]]),
            kis=np.array(c["k_is"], dtype=np.uint64))
        cc.gate_ids = ctypes.cast(self._keep["ids"], ctypes.POINTER(ctypes.c_char_p))
        cc.selector_indices = ctypes.cast(self._keep["sel"], ctypes.POINTER(c_u32))
        cc.num_selector_groups = len(c["groups"])
        cc.group_starts = ctypes.cast(self._keep["gs"], ctypes.POINTER(c_u32))
        cc.group_ends = ctypes.cast(self._keep["ge"], ctypes.POINTER(c_u32))
        cc.k_is = self._keep["kis"].ctypes.data_as(u64p)
        return cc

    def upload(self, ctx):
        """CircuitBuilder::build()'s commitment half from finished constant / sigma polynomial VALUES (b200_circuit_upload);
        the library derives the circuit digest from the cap."""
        from . import lib, c_void_p
        if self._handle is not None: return self
        cc = self.c_common()
        cap = np.zeros((1 << self.common["cap_height"], 4), np.uint64)
        h = c_void_p()
        ctx._chk(lib.b200_circuit_upload(ctx.h, ctypes.byref(cc), self.constants_sigmas.ctypes.data_as(c_void_p), None,
                                         cap.ctypes.data_as(c_void_p), ctypes.byref(h)))
        self.constants_sigmas_cap = cap
        from . import circuit_digest
        self.circuit_digest = circuit_digest(self.common.get("hasher", "gl"), cap, self.common["degree_bits"])
        self._handle = h; self._ctx = ctx
        return self

    def build_on_device(self, ctx):
        """CircuitBuilder::build()'s device half from the builder's own data (b200_circuit_build): selector / sigma polynomial
        generation, commitment and digest on the GPU, with the per-context circuit cache.  Returns True on a cache hit."""
        from . import lib, c_void_p
        assert self._handle is None and self.build_inputs is not None
        bi = self.build_inputs
        cc = self.c_common()
        cap = np.zeros((1 << self.common["cap_height"], 4), np.uint64); dg = np.zeros(4, np.uint64)
        h = c_void_p(); hit = ctypes.c_int(0)
        gc = bi["gate_constants"]
        ctx._chk(lib.b200_circuit_build(ctx.h, ctypes.byref(cc), bi["row_gate"].ctypes.data_as(c_void_p),
                                        gc.ctypes.data_as(c_void_p) if gc.size else None, bi["sigma_next"].ctypes.data_as(c_void_p),
                                        cap.ctypes.data_as(c_void_p), dg.ctypes.data_as(c_void_p), ctypes.byref(hit), ctypes.byref(h)))
        self.constants_sigmas_cap = cap; self.circuit_digest = [int(x) for x in dg]
        self._handle = h; self._ctx = ctx
        return bool(hit.value)

    def fill_u32_witness(self, ctx, wires):
        """run the u32 gates' generators on the device over `wires` ([num_wires][n], in place); returns the number of rows filled"""
        from . import lib, c_void_p
        cc = self.c_common()
        w = np.ascontiguousarray(wires, dtype=np.uint64)
        cnt = ctypes.c_uint32(0)
        ctx._chk(lib.b200_witness_fill_u32(ctx.h, ctypes.byref(cc), self.build_inputs["row_gate"].ctypes.data_as(c_void_p), w.ctypes.data_as(c_void_p), ctypes.byref(cnt)))
        if w is not wires: wires[...] = w
        return int(cnt.value)

    def prove(self, ctx, wires=None, public_inputs=None, pow_witness_hint=0):
        """CircuitData::prove(pw): returns ProofWithPublicInputs::to_bytes()."""
        from . import lib, c_void_p, c_size_t
        self.upload(ctx)
        w = self.wires if wires is None else np.ascontiguousarray(wires, dtype=np.uint64)
        pis = np.array(self.public_inputs if public_inputs is None else public_inputs, dtype=np.uint64)
        out = ctypes.POINTER(ctypes.c_uint8)(); ln = c_size_t()
        ctx._chk(lib.b200_prove(ctx.h, self._handle, w.ctypes.data_as(c_void_p), pis.ctypes.data_as(c_void_p), len(pis),
                                ctypes.c_uint64(pow_witness_hint), ctypes.byref(out), ctypes.byref(ln)))
        b = ctypes.string_at(out, ln.value)
        lib.b200_free(out)
        return b

    def verify(self, proof_bytes):
        """CircuitData::verify(proof): host-side check by the product verifier (b200_verify).  Raises on an invalid proof."""
        from . import verify, B200Error
        assert self.constants_sigmas_cap is not None, "upload() first (the cap and digest come from the build-side commitment)"
        ok, why = verify(self.common, self.constants_sigmas_cap, self.circuit_digest, proof_bytes)
        if not ok: raise B200Error(-6, why)
        return True

    def free(self):
        from . import lib
        if self._handle is not None:
            lib.b200_circuit_free(self._handle); self._handle = None

# ---------------------------------------------------------------------------------------------------------------
# Gate fillers: append `cnt` rows of one gate type with a satisfying witness.  Return (rows, extra).
def _rand_u32(rng, shape): return rng.integers(0, 1 << 32, size=shape, dtype=np.uint64)

def fill_constant(b, consts):
    cnt = consts.shape[0]
    w = np.zeros((cnt, b.cfg.num_wires), np.uint64); w[:, :consts.shape[1]] = consts
    return b.add_rows(constant_gate(consts.shape[1]), consts, w)

def fill_arithmetic(b, rng, cnt, chain=True):
    kind = arithmetic_gate(b.cfg); n = _num_ops(kind)
    c = rng.integers(1, 1 << 62, size=(cnt, 2), dtype=np.uint64)
    w = np.zeros((cnt, b.cfg.num_wires), np.uint64)
    prev = None
    for i in range(n):
        m0 = F.canon(rng.integers(0, 1 << 63, size=cnt, dtype=np.uint64)) if (prev is None or not chain) else prev
        m1 = F.canon(rng.integers(0, 1 << 63, size=cnt, dtype=np.uint64)); ad = F.canon(rng.integers(0, 1 << 63, size=cnt, dtype=np.uint64))
        out = F.add(F.mul(F.mul(m0, m1), c[:, 0]), F.mul(ad, c[:, 1]))
        w[:, 4 * i] = m0; w[:, 4 * i + 1] = m1; w[:, 4 * i + 2] = ad; w[:, 4 * i + 3] = out
        prev = out
    rows = b.add_rows(kind, c, w)
    if chain:
        for i in range(1, n): b.connect(rows, 4 * (i - 1) + 3, rows, 4 * i)
    return rows

def fill_base_sum(b, rng, cnt, limbs=63):
    kind = base_sum_gate(limbs)
    v = rng.integers(0, 1 << 62, size=cnt, dtype=np.uint64)
    w = np.zeros((cnt, b.cfg.num_wires), np.uint64); w[:, 0] = v
    for i in range(limbs): w[:, 1 + i] = (v >> np.uint64(i)) & np.uint64(1)
    return b.add_rows(kind, np.zeros((cnt, 0), np.uint64), w)

def fill_u32_arithmetic(b, rng, cnt, chain=True):
    kind = u32_arithmetic_gate(b.cfg); n = _num_ops(kind)
    w = np.zeros((cnt, b.cfg.num_wires), np.uint64)
    prev = None
    for i in range(n):
        m0 = _rand_u32(rng, cnt) >> np.uint64(1) if (prev is None or not chain) else prev     # hi != u32::MAX
        m1 = _rand_u32(rng, cnt) >> np.uint64(1); ad = _rand_u32(rng, cnt)
        prod = m0 * m1 + ad
        lo = prod & U32; hi = prod >> np.uint64(32)
        inv = F.inv(F.sub(np.full(cnt, 0xFFFFFFFF, np.uint64), hi))
        w[:, 6 * i:6 * i + 6] = np.stack([m0, m1, ad, lo, hi, inv], axis=1)
        for j in range(32): w[:, 6 * n + 32 * i + j] = (prod >> np.uint64(2 * j)) & np.uint64(3)
        prev = lo >> np.uint64(1)
        if chain and i + 1 < n: pass
    # chained input must equal the wire it is copied from: use lo itself (not lo>>1) for exact equality
    if chain:
        for i in range(1, n):
            m0 = w[:, 6 * (i - 1) + 3] >> np.uint64(0)
            m0 = m0 & np.uint64(0x7FFFFFFF) if False else m0
        # recompute sequentially with m0 := previous lo (u32, may exceed 2^31: hi can then reach u32::MAX only with negligible probability)
        for i in range(1, n):
            m0 = w[:, 6 * (i - 1) + 3]; m1 = w[:, 6 * i + 1]; ad = w[:, 6 * i + 2]
            prod = m0 * m1 + ad
            lo = prod & U32; hi = prod >> np.uint64(32)
            assert not (hi == U32).any()
            inv = F.inv(F.sub(np.full(cnt, 0xFFFFFFFF, np.uint64), hi))
            w[:, 6 * i:6 * i + 6] = np.stack([m0, m1, ad, lo, hi, inv], axis=1)
            for j in range(32): w[:, 6 * n + 32 * i + j] = (prod >> np.uint64(2 * j)) & np.uint64(3)
    rows = b.add_rows(kind, np.zeros((cnt, 0), np.uint64), w)
    if chain:
        for i in range(1, n): b.connect(rows, 6 * (i - 1) + 3, rows, 6 * i)
    return rows

def fill_u32_add_many(b, rng, cnt, addends):
    kind = u32_add_many_gate(b.cfg, addends); n = _num_ops(kind); r = addends + 3
    w = np.zeros((cnt, b.cfg.num_wires), np.uint64)
    for i in range(n):
        ad = _rand_u32(rng, (cnt, addends)); cin = rng.integers(0, 4, size=cnt, dtype=np.uint64)
        s = ad.sum(axis=1, dtype=np.uint64) + cin
        res = s & U32; cout = s >> np.uint64(32)
        w[:, r * i:r * i + addends] = ad; w[:, r * i + addends] = cin; w[:, r * i + addends + 1] = res; w[:, r * i + addends + 2] = cout
        for j in range(16): w[:, r * n + 18 * i + j] = (res >> np.uint64(2 * j)) & np.uint64(3)
        for j in range(2): w[:, r * n + 18 * i + 16 + j] = (cout >> np.uint64(2 * j)) & np.uint64(3)
    return b.add_rows(kind, np.zeros((cnt, 0), np.uint64), w)

def fill_u32_subtraction(b, rng, cnt):
    kind = u32_subtraction_gate(b.cfg); n = _num_ops(kind)
    w = np.zeros((cnt, b.cfg.num_wires), np.uint64)
    for i in range(n):
        x = _rand_u32(rng, cnt); y = _rand_u32(rng, cnt); bi = rng.integers(0, 2, size=cnt, dtype=np.uint64)
        under = x < y + bi
        res = np.where(under, x + np.uint64(1 << 32) - y - bi, x - y - bi)
        w[:, 5 * i:5 * i + 5] = np.stack([x, y, bi, res, under.astype(np.uint64)], axis=1)
        for j in range(16): w[:, 5 * n + 16 * i + j] = (res >> np.uint64(2 * j)) & np.uint64(3)
    return b.add_rows(kind, np.zeros((cnt, 0), np.uint64), w)

def fill_u32_range_check(b, rng, cnt, limbs=8):
    kind = u32_range_check_gate(limbs)
    w = np.zeros((cnt, b.cfg.num_wires), np.uint64)
    x = _rand_u32(rng, (cnt, limbs)); w[:, :limbs] = x
    for i in range(limbs):
        for j in range(16): w[:, limbs + 16 * i + j] = (x[:, i] >> np.uint64(2 * j)) & np.uint64(3)
    return b.add_rows(kind, np.zeros((cnt, 0), np.uint64), w)

def _spread(x):     # bit k of x -> bit 2k
    out = np.zeros_like(x)
    for k in range(32): out |= ((x >> np.uint64(k)) & np.uint64(1)) << np.uint64(2 * k)
    return out

def fill_u32_interleave(b, rng, cnt):
    kind = u32_interleave_gate(b.cfg); n = _num_ops(kind)
    w = np.zeros((cnt, b.cfg.num_wires), np.uint64)
    for i in range(n):
        x = _rand_u32(rng, cnt)
        w[:, 2 * i] = x; w[:, 2 * i + 1] = _spread(x)
        for j in range(32): w[:, 2 * n + 32 * i + j] = (x >> np.uint64(31 - j)) & np.uint64(1)
    return b.add_rows(kind, np.zeros((cnt, 0), np.uint64), w)

def fill_uninterleave(b, rng, cnt, to_b32):
    kind = uninterleave_to_b32_gate(b.cfg) if to_b32 else uninterleave_to_u32_gate(b.cfg); n = _num_ops(kind)
    w = np.zeros((cnt, b.cfg.num_wires), np.uint64)
    for i in range(n):
        x = rng.integers(0, 1 << 63, size=cnt, dtype=np.uint64)
        ev = np.zeros(cnt, np.uint64); od = np.zeros(cnt, np.uint64)
        for j in range(64): w[:, 3 * n + 64 * i + j] = (x >> np.uint64(63 - j)) & np.uint64(1)
        for j in range(32):
            sh = np.uint64(2 * (31 - j)) if to_b32 else np.uint64(31 - j)
            ev |= w[:, 3 * n + 64 * i + 2 * j] << sh; od |= w[:, 3 * n + 64 * i + 2 * j + 1] << sh
        w[:, 3 * i] = x; w[:, 3 * i + 1] = ev; w[:, 3 * i + 2] = od
    return b.add_rows(kind, np.zeros((cnt, 0), np.uint64), w)

def fill_comparison(b, rng, cnt, bits=32, chunks=16):
    kind = comparison_gate(bits, chunks); cb = -(-bits // chunks)
    w = np.zeros((cnt, b.cfg.num_wires), np.uint64)
    first = _rand_u32(rng, cnt); second = _rand_u32(rng, cnt)
    eqmask = rng.integers(0, 8, size=cnt) == 0
    second = np.where(eqmask, first, second)
    w[:, 0] = first; w[:, 1] = second
    msd = np.zeros(cnt, np.uint64)
    mask = np.uint64((1 << cb) - 1)
    for i in range(chunks):
        f = (first >> np.uint64(cb * i)) & mask; s = (second >> np.uint64(cb * i)) & mask
        diff = F.sub(s, f); eq = (f == s)
        dummy = np.where(eq, np.uint64(1), F.inv(np.where(eq, np.uint64(1), diff)))      # ONE when equal, as ComparisonGenerator (comparison.rs:472-476)
        inter = np.where(eq, msd, np.uint64(0))
        w[:, 4 + i] = f; w[:, 4 + chunks + i] = s; w[:, 4 + 2 * chunks + i] = dummy
        w[:, 4 + 3 * chunks + i] = eq.astype(np.uint64); w[:, 4 + 4 * chunks + i] = inter
        msd = np.where(eq, msd, diff)
    w[:, 3] = msd
    tot = F.add(np.full(cnt, 1 << cb, np.uint64), msd)          # in [0, 2^(cb+1))
    for i in range(cb + 1): w[:, 4 + 5 * chunks + i] = (tot >> np.uint64(i)) & np.uint64(1)
    w[:, 2] = w[:, 4 + 5 * chunks + cb]
    return b.add_rows(kind, np.zeros((cnt, 0), np.uint64), w)

# ---- Poseidon (vectorised, with the S-box-input trace PoseidonGate stores as wires) ----
from . import poseidon_np

def fill_poseidon(b, inputs, swap):
    """inputs [cnt][12], swap [cnt] in {0,1}.  Returns (rows, outputs [cnt][12])."""
    cnt = inputs.shape[0]
    w = np.zeros((cnt, b.cfg.num_wires), np.uint64)
    w[:, 0:12] = inputs; w[:, 24] = swap
    st = inputs.copy()
    for i in range(4):
        delta = F.mul(swap, F.sub(inputs[:, i + 4], inputs[:, i]))
        w[:, 25 + i] = delta
        st[:, i] = F.add(inputs[:, i], delta); st[:, i + 4] = F.sub(inputs[:, i + 4], delta)
    out, full0, partial, full1 = poseidon_np.permute_trace(st)
    for r in range(1, 4): w[:, 29 + 12 * (r - 1):29 + 12 * r] = full0[r]
    for r in range(22): w[:, 65 + r] = partial[r]
    for r in range(4): w[:, 87 + 12 * r:87 + 12 * (r + 1)] = full1[r]
    w[:, 12:24] = out
    rows = b.add_rows(poseidon_gate(), np.zeros((cnt, 0), np.uint64), w)
    return rows, out

def fill_poseidon_chain(b, rng, cnt, chain_len=16):
    """cnt rows organised as Merkle-path-like chains of `chain_len` rows: row r+1 input[4..8) = row r output[0..4)
    (cross-row copy constraints).  Chains are independent, so every step is vectorised over all chains."""
    chain_len = max(1, min(chain_len, cnt))
    nchains = -(-cnt // chain_len)
    total = nchains * chain_len
    all_in = np.zeros((chain_len, nchains, 12), np.uint64)
    sw = rng.integers(0, 2, size=(chain_len, nchains), dtype=np.uint64)
    cur = F.canon(rng.integers(0, 1 << 63, size=(nchains, 12), dtype=np.uint64))
    for step in range(chain_len):
        all_in[step] = cur
        st = cur.copy()
        m = sw[step].astype(bool)
        st[m, 0:4], st[m, 4:8] = cur[m, 4:8], cur[m, 0:4]
        out = poseidon_np.permute(st)
        nxt = F.canon(rng.integers(0, 1 << 63, size=(nchains, 12), dtype=np.uint64)); nxt[:, 4:8] = out[:, 0:4]
        cur = nxt
    # row order: chain-major (rows of one chain are consecutive); drop the surplus rows of the last chain
    inp = all_in.transpose(1, 0, 2).reshape(total, 12)[:cnt]
    swf = sw.transpose(1, 0).reshape(total)[:cnt]
    rows, _ = fill_poseidon(b, inp, swf)
    idx = np.arange(cnt)
    link = idx[(idx % chain_len) != chain_len - 1]
    link = link[link + 1 < cnt]
    for j in range(4):
        b.connect(rows[link], 12 + j, rows[link + 1], 4 + j)
    return rows

# ---------------------------------------------------------------------------------------------------------------
# The remaining standard gates of a recursion circuit (common_data.json: gates[2..11]); extension-field values are
# pairs of base wires, arithmetic in F_p[X]/(X^2-7) vectorised over rows.
def arithmetic_extension_gate(cfg): n = cfg.num_routed_wires // 8; return GateKind(f"ArithmeticExtensionGate {{ num_ops: {n} }}", 3, 2, 8 * n)
def mul_extension_gate(cfg): n = cfg.num_routed_wires // 6; return GateKind(f"MulExtensionGate {{ num_ops: {n} }}", 3, 1, 6 * n)
def reducing_gate(cfg): n = (cfg.num_wires - 2 - 2 * 2) // 3 - (0); n = min(n, (cfg.num_routed_wires - 6)); n = 43 if cfg.num_wires == 135 else n; return GateKind(f"ReducingGate {{ num_coeffs: {n} }}", 2, 0, 6 + n + 2 * (n - 1))
def reducing_extension_gate(cfg): n = 32 if cfg.num_wires == 135 else (cfg.num_wires - 4) // 4 - 1; return GateKind(f"ReducingExtensionGate {{ num_coeffs: {n} }}", 2, 0, 6 + 2 * n + 2 * (n - 1))
def random_access_gate(bits=4, copies=4, extra=2):
    return GateKind(f"RandomAccessGate {{ bits: {bits}, num_copies: {copies}, num_extra_constants: {extra}, {_PH} }}<D=2>", bits + 1, extra,
                    (2 + (1 << bits)) * copies + extra + copies * bits)
def poseidon_mds_gate(): return GateKind("PoseidonMdsGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>", 1, 0, 48)
COSET_WEIGHTS_4 = [17293822565076172801, 18374686475376656385, 18446744069413535745, 281474976645120, 17592186044416, 256,
                   18446744000695107601, 18446744065119617025, 1152921504338411520, 72057594037927936, 1048576,
                   18446462594437939201, 18446726477228539905, 18446744069414584065, 68719476720, 4294967296]
def coset_interpolation_gate():
    ws = ", ".join(str(x) for x in COSET_WEIGHTS_4)
    return GateKind(f"CosetInterpolationGate {{ subgroup_bits: 4, degree: 6, barycentric_weights: [{ws}], {_PH} }}<D=2>", 6, 0, 47)

def _rf(rng, shape): return F.canon(rng.integers(0, 1 << 63, size=shape, dtype=np.uint64))
def _emul(a, b):   # a, b: tuples (re, im) of arrays
    seven = np.uint64(7)
    return (F.add(F.mul(a[0], b[0]), F.mul(F.mul(a[1], b[1]), np.full_like(a[0], seven))), F.add(F.mul(a[0], b[1]), F.mul(a[1], b[0])))
def _eadd(a, b): return (F.add(a[0], b[0]), F.add(a[1], b[1]))
def _esub(a, b): return (F.sub(a[0], b[0]), F.sub(a[1], b[1]))
def _escale(a, s): return (F.mul(a[0], s), F.mul(a[1], s))

def fill_arithmetic_extension(b, rng, cnt):
    kind = arithmetic_extension_gate(b.cfg); n = _num_ops(kind)
    c = _rf(rng, (cnt, 2)); w = np.zeros((cnt, b.cfg.num_wires), np.uint64)
    for i in range(n):
        m0 = (_rf(rng, cnt), _rf(rng, cnt)); m1 = (_rf(rng, cnt), _rf(rng, cnt)); ad = (_rf(rng, cnt), _rf(rng, cnt))
        out = _eadd(_escale(_emul(m0, m1), c[:, 0]), _escale(ad, c[:, 1]))
        for k, v in enumerate((m0, m1, ad, out)): w[:, 8 * i + 2 * k] = v[0]; w[:, 8 * i + 2 * k + 1] = v[1]
    return b.add_rows(kind, c, w)

def fill_mul_extension(b, rng, cnt):
    kind = mul_extension_gate(b.cfg); n = _num_ops(kind)
    c = _rf(rng, (cnt, 1)); w = np.zeros((cnt, b.cfg.num_wires), np.uint64)
    for i in range(n):
        m0 = (_rf(rng, cnt), _rf(rng, cnt)); m1 = (_rf(rng, cnt), _rf(rng, cnt))
        out = _escale(_emul(m0, m1), c[:, 0])
        for k, v in enumerate((m0, m1, out)): w[:, 6 * i + 2 * k] = v[0]; w[:, 6 * i + 2 * k + 1] = v[1]
    return b.add_rows(kind, c, w)

def fill_reducing(b, rng, cnt, ext):
    import re
    kind = reducing_extension_gate(b.cfg) if ext else reducing_gate(b.cfg)
    n = int(re.search(r"num_coeffs: (\d+)", kind.id)[1])
    w = np.zeros((cnt, b.cfg.num_wires), np.uint64)
    alpha = (_rf(rng, cnt), _rf(rng, cnt)); acc = (_rf(rng, cnt), _rf(rng, cnt))
    w[:, 2], w[:, 3] = alpha; w[:, 4], w[:, 5] = acc
    start_accs = 6 + (2 * n if ext else n)
    for i in range(n):
        if ext:
            co = (_rf(rng, cnt), _rf(rng, cnt)); w[:, 6 + 2 * i], w[:, 7 + 2 * i] = co
        else:
            co = (_rf(rng, cnt), np.zeros(cnt, np.uint64)); w[:, 6 + i] = co[0]
        acc = _eadd(_emul(acc, alpha), co)
        if i == n - 1: w[:, 0], w[:, 1] = acc
        else: w[:, start_accs + 2 * i], w[:, start_accs + 2 * i + 1] = acc
    return b.add_rows(kind, np.zeros((cnt, 0), np.uint64), w)

def fill_random_access(b, rng, cnt, bits=4, copies=4, extra=2):
    kind = random_access_gate(bits, copies, extra); vec = 1 << bits
    c = _rf(rng, (cnt, extra)); w = np.zeros((cnt, b.cfg.num_wires), np.uint64)
    routed = (2 + vec) * copies + extra
    for cp in range(copies):
        base = (2 + vec) * cp
        idx = rng.integers(0, vec, size=cnt, dtype=np.uint64)
        items = _rf(rng, (cnt, vec))
        w[:, base] = idx; w[:, base + 2:base + 2 + vec] = items
        w[:, base + 1] = items[np.arange(cnt), idx.astype(np.int64)]
        for i in range(bits): w[:, routed + cp * bits + i] = (idx >> np.uint64(i)) & np.uint64(1)
    for i in range(extra): w[:, (2 + vec) * copies + i] = c[:, i]
    return b.add_rows(kind, c, w)

def fill_poseidon_mds(b, rng, cnt):
    from . import poseidon_np
    w = np.zeros((cnt, b.cfg.num_wires), np.uint64)
    inp = _rf(rng, (cnt, 24)); w[:, :24] = inp
    for comp in range(2):
        out = poseidon_np._mds(inp[:, comp::2])
        w[:, 24 + comp:48:2] = out
    return b.add_rows(poseidon_mds_gate(), np.zeros((cnt, 0), np.uint64), w)

def fill_coset_interpolation(b, rng, cnt):
    """FRI-fold check gate: interpolate 16 extension values on shift*<w_16> and evaluate at a point (barycentric,
    split in chunks of `degree` with stored intermediates).  coset_interpolation_gate.go:151-226."""
    kind = coset_interpolation_gate(); sb, degree, npts = 4, 6, 16
    n_inter = (npts - 2) // (degree - 1); start_inter = 1 + 2 * npts + 4
    w = np.zeros((cnt, b.cfg.num_wires), np.uint64)
    shift = _rf(rng, cnt) | np.uint64(1)
    w[:, 0] = shift
    vals = [(_rf(rng, cnt), _rf(rng, cnt)) for _ in range(npts)]
    for i, v in enumerate(vals): w[:, 1 + 2 * i], w[:, 2 + 2 * i] = v
    ep = (_rf(rng, cnt), _rf(rng, cnt))
    sinv = F.inv(shift)
    sep = _escale(ep, sinv)                                        # evaluation_point = shift * shifted_evaluation_point
    w[:, 1 + 2 * npts], w[:, 2 + 2 * npts] = ep
    w[:, start_inter + 4 * n_inter], w[:, start_inter + 4 * n_inter + 1] = sep
    dom = F.powers(F.root_of_unity(sb), npts)
    ev = (np.zeros(cnt, np.uint64), np.zeros(cnt, np.uint64)); prod = (np.ones(cnt, np.uint64), np.zeros(cnt, np.uint64))
    nxt, k = degree, 0
    for i in range(npts):
        if i == nxt and k < n_inter:
            w[:, start_inter + 2 * k], w[:, start_inter + 2 * k + 1] = ev
            w[:, start_inter + 2 * (n_inter + k)], w[:, start_inter + 2 * (n_inter + k) + 1] = prod
            k += 1; nxt += degree - 1
        term = _esub(sep, (np.full(cnt, dom[i], np.uint64), np.zeros(cnt, np.uint64)))
        wv = _escale(vals[i], np.full(cnt, COSET_WEIGHTS_4[i] % P, np.uint64))
        ev = _eadd(_emul(ev, term), _emul(wv, prod)); prod = _emul(prod, term)
    w[:, 1 + 2 * npts + 2], w[:, 1 + 2 * npts + 3] = ev
    return b.add_rows(kind, np.zeros((cnt, 0), np.uint64), w)
