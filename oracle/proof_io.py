"""plonky2 proof / circuit-data wire formats (serde JSON and the binary `to_bytes` layout).

TEST INFRASTRUCTURE.  JSON shape: gnark-plonky2-verifier/types/deserialize.go, variables/deserialize.go,
types/common_data.go:11-58.  Binary layout: upstream plonky2 util/serialization (not in tree), pinned
byte-for-byte by near_bft_finality/proofs/*/proof.bin <-> proof.json (tests/test_oracle_proof_io.py):
  caps (wires, zs_pp, quotient) | openings (ext = 2 x u64 LE, no length prefixes) |
  FRI: commit caps | per query: 4 x (leaf u64s, u8 n_siblings, siblings) ; per step (arity ext evals,
  u8 n_siblings, siblings) | final poly ext coeffs | pow_witness u64 | u64 n_public_inputs | PIs u64 LE.
Hash encoding: "bn254" = 32 bytes LE of Fr (JSON: decimal string); "gl" = 4 x u64 LE (JSON: {"elements":[..]}).
"""
import json, struct
from dataclasses import dataclass, field
from typing import List, Any

@dataclass
class FriConfig:
    rate_bits: int = 3
    cap_height: int = 4
    proof_of_work_bits: int = 16
    num_query_rounds: int = 28

@dataclass
class CommonData:
    num_wires: int
    num_routed_wires: int
    num_constants_cfg: int
    num_challenges: int
    fri: FriConfig
    degree_bits: int
    reduction_arity_bits: List[int]
    gates: List[str]
    selector_indices: List[int]
    groups: List[tuple]
    quotient_degree_factor: int
    num_gate_constraints: int
    num_constants: int
    num_public_inputs: int
    k_is: List[int]
    num_partial_products: int
    hasher: str = "gl"       # "gl" | "bn254"  (Merkle/sponge hasher of C::Hasher)

    @property
    def num_selectors(self): return len(self.groups)
    @property
    def lde_bits(self): return self.degree_bits + self.fri.rate_bits
    @property
    def final_poly_len(self): return 1 << (self.degree_bits - sum(self.reduction_arity_bits))
    @property
    def num_zs_pp(self): return self.num_challenges * (1 + self.num_partial_products)
    @property
    def num_quotient(self): return self.num_challenges * self.quotient_degree_factor
    @property
    def num_preprocessed(self): return self.num_constants + self.num_routed_wires
    def oracle_widths(self):
        return [self.num_preprocessed, self.num_wires, self.num_zs_pp, self.num_quotient]

def common_from_json(d, hasher):
    c = d["config"]; f = d["fri_params"]
    return CommonData(
        num_wires=c["num_wires"], num_routed_wires=c["num_routed_wires"], num_constants_cfg=c["num_constants"],
        num_challenges=c["num_challenges"],
        fri=FriConfig(f["config"]["rate_bits"], f["config"]["cap_height"], f["config"]["proof_of_work_bits"],
                      f["config"]["num_query_rounds"]),
        degree_bits=f["degree_bits"], reduction_arity_bits=list(f["reduction_arity_bits"]),
        gates=list(d["gates"]), selector_indices=list(d["selectors_info"]["selector_indices"]),
        groups=[(g["start"], g["end"]) for g in d["selectors_info"]["groups"]],
        quotient_degree_factor=d["quotient_degree_factor"], num_gate_constraints=d["num_gate_constraints"],
        num_constants=d["num_constants"], num_public_inputs=d["num_public_inputs"], k_is=list(d["k_is"]),
        num_partial_products=d["num_partial_products"], hasher=hasher)

def common_to_json(cd: CommonData):
    fc = {"rate_bits": cd.fri.rate_bits, "cap_height": cd.fri.cap_height,
          "proof_of_work_bits": cd.fri.proof_of_work_bits,
          "reduction_strategy": {"ConstantArityBits": [4, 5]}, "num_query_rounds": cd.fri.num_query_rounds}
    return {
        "config": {"num_wires": cd.num_wires, "num_routed_wires": cd.num_routed_wires,
                   "num_constants": cd.num_constants_cfg, "use_base_arithmetic_gate": True, "security_bits": 100,
                   "num_challenges": cd.num_challenges, "zero_knowledge": False, "max_quotient_degree_factor": 8,
                   "fri_config": fc},
        "fri_params": {"config": fc, "hiding": False, "degree_bits": cd.degree_bits,
                       "reduction_arity_bits": cd.reduction_arity_bits},
        "gates": cd.gates,
        "selectors_info": {"selector_indices": cd.selector_indices,
                           "groups": [{"start": s, "end": e} for s, e in cd.groups]},
        "quotient_degree_factor": cd.quotient_degree_factor, "num_gate_constraints": cd.num_gate_constraints,
        "num_constants": cd.num_constants, "num_public_inputs": cd.num_public_inputs, "k_is": cd.k_is,
        "num_partial_products": cd.num_partial_products, "num_lookup_polys": 0, "num_lookup_selectors": 0, "luts": []}

@dataclass
class VerifierOnly:
    constants_sigmas_cap: List[Any]
    circuit_digest: Any

@dataclass
class QueryStep:
    evals: List[tuple]
    siblings: List[Any]

@dataclass
class QueryRound:
    initial: List[tuple]          # per oracle: (leaf_elements, siblings)
    steps: List[QueryStep]

@dataclass
class Proof:
    wires_cap: List[Any]
    zs_pp_cap: List[Any]
    quotient_cap: List[Any]
    constants: List[tuple]
    plonk_sigmas: List[tuple]
    wires: List[tuple]
    plonk_zs: List[tuple]
    plonk_zs_next: List[tuple]
    partial_products: List[tuple]
    quotient_polys: List[tuple]
    commit_phase_caps: List[List[Any]]
    query_rounds: List[QueryRound]
    final_poly: List[tuple]
    pow_witness: int
    public_inputs: List[int] = field(default_factory=list)

# ---------------- hashes ----------------
def _hash_from_json(h, hasher):
    if hasher == "bn254":
        return int(h)
    return tuple(h["elements"])
def _hash_to_json(h, hasher):
    if hasher == "bn254":
        return str(h)
    return {"elements": list(h)}
def _hash_to_bytes(h, hasher):
    if hasher == "bn254":
        return int(h).to_bytes(32, "little")
    return struct.pack("<4Q", *h)
def _hash_size(hasher): return 32

def verifier_only_from_json(d, hasher):
    return VerifierOnly([_hash_from_json(h, hasher) for h in d["constants_sigmas_cap"]],
                        _hash_from_json(d["circuit_digest"], hasher))

def verifier_only_to_json(v: VerifierOnly, hasher):
    return {"constants_sigmas_cap": [_hash_to_json(h, hasher) for h in v.constants_sigmas_cap],
            "circuit_digest": _hash_to_json(v.circuit_digest, hasher)}

# ---------------- JSON ----------------
def proof_from_json(d, hasher):
    p = d["proof"]; o = p["openings"]; f = p["opening_proof"]
    ext = lambda xs: [tuple(x) for x in xs]
    H = lambda xs: [_hash_from_json(h, hasher) for h in xs]
    rounds = []
    for q in f["query_round_proofs"]:
        init = [(list(ev), H(mp["siblings"])) for ev, mp in q["initial_trees_proof"]["evals_proofs"]]
        steps = [QueryStep(ext(s["evals"]), H(s["merkle_proof"]["siblings"])) for s in q["steps"]]
        rounds.append(QueryRound(init, steps))
    return Proof(H(p["wires_cap"]), H(p["plonk_zs_partial_products_cap"]), H(p["quotient_polys_cap"]),
                 ext(o["constants"]), ext(o["plonk_sigmas"]), ext(o["wires"]), ext(o["plonk_zs"]),
                 ext(o["plonk_zs_next"]), ext(o["partial_products"]), ext(o["quotient_polys"]),
                 [H(c) for c in f["commit_phase_merkle_caps"]], rounds, ext(f["final_poly"]["coeffs"]),
                 f["pow_witness"], list(d["public_inputs"]))

def proof_to_json(pr: Proof, hasher):
    H = lambda xs: [_hash_to_json(h, hasher) for h in xs]
    ext = lambda xs: [list(x) for x in xs]
    return {"proof": {
        "wires_cap": H(pr.wires_cap), "plonk_zs_partial_products_cap": H(pr.zs_pp_cap),
        "quotient_polys_cap": H(pr.quotient_cap),
        "openings": {"constants": ext(pr.constants), "plonk_sigmas": ext(pr.plonk_sigmas), "wires": ext(pr.wires),
                     "plonk_zs": ext(pr.plonk_zs), "plonk_zs_next": ext(pr.plonk_zs_next),
                     "partial_products": ext(pr.partial_products), "quotient_polys": ext(pr.quotient_polys),
                     "lookup_zs": [], "lookup_zs_next": []},
        "opening_proof": {
            "commit_phase_merkle_caps": [H(c) for c in pr.commit_phase_caps],
            "query_round_proofs": [{
                "initial_trees_proof": {"evals_proofs": [[list(ev), {"siblings": H(sib)}] for ev, sib in q.initial]},
                "steps": [{"evals": ext(s.evals), "merkle_proof": {"siblings": H(s.siblings)}} for s in q.steps]}
                for q in pr.query_rounds],
            "final_poly": {"coeffs": ext(pr.final_poly)}, "pow_witness": pr.pow_witness}},
        "public_inputs": list(pr.public_inputs)}

# ---------------- binary ----------------
def proof_to_bytes(pr: Proof, hasher):
    out = bytearray()
    def cap(c):
        for h in c: out.extend(_hash_to_bytes(h, hasher))
    def exts(xs):
        for a, b in xs: out.extend(struct.pack("<2Q", a, b))
    def mproof(sib):
        out.append(len(sib))
        for h in sib: out.extend(_hash_to_bytes(h, hasher))
    cap(pr.wires_cap); cap(pr.zs_pp_cap); cap(pr.quotient_cap)
    for v in (pr.constants, pr.plonk_sigmas, pr.wires, pr.plonk_zs, pr.plonk_zs_next, pr.partial_products,
              pr.quotient_polys):
        exts(v)
    for c in pr.commit_phase_caps: cap(c)
    for q in pr.query_rounds:
        for ev, sib in q.initial:
            out.extend(struct.pack(f"<{len(ev)}Q", *ev)); mproof(sib)
        for s in q.steps:
            exts(s.evals); mproof(s.siblings)
    exts(pr.final_poly)
    out.extend(struct.pack("<Q", pr.pow_witness))
    out.extend(struct.pack("<Q", len(pr.public_inputs)))
    out.extend(struct.pack(f"<{len(pr.public_inputs)}Q", *pr.public_inputs))
    return bytes(out)

class _Reader:
    def __init__(self, b): self.b = b; self.o = 0
    def u8(self): v = self.b[self.o]; self.o += 1; return v
    def u64(self): v = struct.unpack_from("<Q", self.b, self.o)[0]; self.o += 8; return v
    def u64s(self, n): v = list(struct.unpack_from(f"<{n}Q", self.b, self.o)); self.o += 8 * n; return v
    def exts(self, n): v = self.u64s(2 * n); return [(v[2 * i], v[2 * i + 1]) for i in range(n)]
    def hash(self, hasher):
        if hasher == "bn254":
            v = int.from_bytes(self.b[self.o:self.o + 32], "little"); self.o += 32; return v
        return tuple(self.u64s(4))

def proof_from_bytes(b, cd: CommonData):
    r = _Reader(b); hs = cd.hasher
    ncap = 1 << cd.fri.cap_height
    cap = lambda: [r.hash(hs) for _ in range(ncap)]
    mproof = lambda: [r.hash(hs) for _ in range(r.u8())]
    wires_cap, zs_cap, q_cap = cap(), cap(), cap()
    constants = r.exts(cd.num_constants); sigmas = r.exts(cd.num_routed_wires); wires = r.exts(cd.num_wires)
    zs = r.exts(cd.num_challenges); zs_next = r.exts(cd.num_challenges)
    pps = r.exts(cd.num_challenges * cd.num_partial_products); quot = r.exts(cd.num_quotient)
    caps = [cap() for _ in cd.reduction_arity_bits]
    rounds = []
    for _ in range(cd.fri.num_query_rounds):
        init = []
        for w in cd.oracle_widths():
            ev = r.u64s(w); init.append((ev, mproof()))
        steps = []
        for ab in cd.reduction_arity_bits:
            ev = r.exts(1 << ab); steps.append(QueryStep(ev, mproof()))
        rounds.append(QueryRound(init, steps))
    final_poly = r.exts(cd.final_poly_len)
    pow_witness = r.u64()
    npi = r.u64()
    pis = r.u64s(npi)
    assert r.o == len(b), (r.o, len(b))
    return Proof(wires_cap, zs_cap, q_cap, constants, sigmas, wires, zs, zs_next, pps, quot, caps, rounds,
                 final_poly, pow_witness, pis)
