"""Full plonky2 verifier (CPU, exact).  TEST INFRASTRUCTURE — the checker for every proof the CUDA prover emits.

Restates:
  gnark-plonky2-verifier/verifier/verifier.go:45-82   challenge derivation (transcript order)
  gnark-plonky2-verifier/plonk/plonk.go:57-250        vanishing-polynomial identity at zeta
  gnark-plonky2-verifier/fri/fri.go:63-548            FRI: PoW, initial Merkle openings, combine, fold, final poly
  gnark-plonky2-verifier/challenger/challenger.go:117-144  FRI challenges
Generic over C::Hasher in {"gl" (PoseidonGoldilocksConfig), "bn254" (PoseidonBN128GoldilocksConfig)}.
Pinned by: testdata/test_circuit (G5 + the G4 Fiat-Shamir KATs) and the three near_bft_finality/proofs (G6).
"""
from . import gl, poseidon_gl as pg, poseidon_bn254 as pb
from .gl import P, Fp2
from .challenger import Challenger
from .gates import gate_from_id, eval_filtered_constraints
from .proof_io import CommonData, Proof, VerifierOnly

class VerifyError(Exception):
    pass

class Hasher:
    def __init__(self, kind):
        self.kind = kind
    def hash_or_noop(self, xs):
        return pb.hash_or_noop(xs) if self.kind == "bn254" else tuple(pg.hash_or_noop(xs))
    def two_to_one(self, l, r):
        return pb.two_to_one(l, r) if self.kind == "bn254" else tuple(pg.two_to_one(l, r))
    def to_vec(self, h):
        return pb.to_vec(h) if self.kind == "bn254" else list(h)

def get_challenges(cd: CommonData, vo: VerifierOnly, pr: Proof, pi_hash):
    H = Hasher(cd.hasher)
    ch = Challenger()
    ch.observe_many(H.to_vec(vo.circuit_digest))
    ch.observe_many(pi_hash)
    for h in pr.wires_cap: ch.observe_many(H.to_vec(h))
    betas = ch.get_n(cd.num_challenges)
    gammas = ch.get_n(cd.num_challenges)
    for h in pr.zs_pp_cap: ch.observe_many(H.to_vec(h))
    alphas = ch.get_n(cd.num_challenges)
    for h in pr.quotient_cap: ch.observe_many(H.to_vec(h))
    zeta = ch.get_ext()
    for v in (pr.constants, pr.plonk_sigmas, pr.wires, pr.plonk_zs, pr.partial_products, pr.quotient_polys):
        ch.observe_ext_many(v)
    ch.observe_ext_many(pr.plonk_zs_next)
    fri_alpha = ch.get_ext()
    fri_betas = []
    for cap in pr.commit_phase_caps:
        for h in cap: ch.observe_many(H.to_vec(h))
        fri_betas.append(ch.get_ext())
    ch.observe_ext_many(pr.final_poly)
    ch.observe(pr.pow_witness)
    pow_response = ch.get_challenge()
    query_indices = ch.get_n(cd.fri.num_query_rounds)
    return dict(betas=betas, gammas=gammas, alphas=alphas, zeta=zeta, fri_alpha=fri_alpha, fri_betas=fri_betas,
                pow_response=pow_response, query_indices=query_indices)

def eval_vanishing(cd, gate_objs, pr, ch, pi_hash):
    """plonk.go:121-207 in F_p^2."""
    F = Fp2
    zeta = Fp2.lift(ch["zeta"])
    n = 1 << cd.degree_bits
    zeta_pow_n = Fp2.lift(gl.eexp_pow2(ch["zeta"], cd.degree_bits))
    consts = [Fp2.lift(x) for x in pr.constants]
    wires = [Fp2.lift(x) for x in pr.wires]
    constraint_terms = eval_filtered_constraints(F, cd, gate_objs, consts, wires, pi_hash)
    # L_0(zeta) = (zeta^n - 1) / (n (zeta - 1))
    l0 = Fp2.lift(gl.ediv((zeta_pow_n - 1).tup(), (zeta * n - n).tup()))
    sigmas = [Fp2.lift(x) for x in pr.plonk_sigmas]
    z1_terms, pp_terms = [], []
    npp = cd.num_partial_products; qd = cd.quotient_degree_factor
    for i in range(cd.num_challenges):
        z = Fp2.lift(pr.plonk_zs[i]); zn = Fp2.lift(pr.plonk_zs_next[i])
        z1_terms.append(l0 * (z - 1))
        beta, gamma = ch["betas"][i], ch["gammas"][i]
        nums = [wires[j] + zeta * (cd.k_is[j] * beta % P) + gamma for j in range(cd.num_routed_wires)]
        dens = [wires[j] + sigmas[j] * beta + gamma for j in range(cd.num_routed_wires)]
        accs = [z] + [Fp2.lift(x) for x in pr.partial_products[i * npp:(i + 1) * npp]] + [zn]
        for k in range(npp + 1):
            np_, dp_ = F.const(1), F.const(1)
            for j in range(k * qd, min((k + 1) * qd, cd.num_routed_wires)):
                np_ = np_ * nums[j]; dp_ = dp_ * dens[j]
            pp_terms.append(accs[k] * np_ - accs[k + 1] * dp_)
    terms = z1_terms + pp_terms + constraint_terms
    out = []
    for a in ch["alphas"]:
        acc = F.const(0)
        for t in reversed(terms):
            acc = acc * a + t
        out.append(acc)
    return out, zeta_pow_n

def verify_merkle(H, leaf, index, siblings, cap):
    d = H.hash_or_noop(leaf)
    for s in siblings:
        d = H.two_to_one(s, d) if index & 1 else H.two_to_one(d, s)
        index >>= 1
    if cap[index] != d:
        raise VerifyError("merkle proof does not match cap")

def fri_batches(cd):
    """fri_utils.go:118-193 : (oracle, poly) lists for the zeta batch and the g*zeta batch."""
    allp = []
    for o, wdt in enumerate(cd.oracle_widths()):
        allp += [(o, i) for i in range(wdt)]
    zs = [(2, i) for i in range(cd.num_challenges)]
    return [allp, zs]

def compute_evaluation(x, x_index_within_coset, arity_bits, evals, beta):
    """fri.go:244-316 : interpolate the arity evals on the coset and evaluate at beta (barycentric)."""
    arity = 1 << arity_bits
    g = gl.primitive_root_of_unity(arity_bits)
    g_inv = gl.inv(g)
    perm = [None] * arity
    for i in range(arity):
        perm[gl.reverse_bits(i, arity_bits)] = evals[i]
    rev = gl.reverse_bits(x_index_within_coset, arity_bits)
    coset_start = x * pow(g_inv, rev, P) % P
    xs = [coset_start * pow(g, i, P) % P for i in range(arity)]
    # barycentric weights
    ws = []
    for i in range(arity):
        w = 1
        for j in range(arity):
            if i != j: w = w * (xs[i] - xs[j]) % P
        ws.append(gl.inv(w))
    lx = gl.E_ONE
    for xi in xs: lx = gl.emul(lx, gl.esub(beta, (xi, 0)))
    s = gl.E_ZERO
    for xi, yi, wi in zip(xs, perm, ws):
        d = gl.esub(beta, (xi, 0))
        if d == gl.E_ZERO: return yi
        s = gl.eadd(s, gl.emul(yi, gl.escale(gl.einv(d), wi)))
    return gl.emul(lx, s)

def verify_fri(cd: CommonData, vo: VerifierOnly, pr: Proof, ch):
    H = Hasher(cd.hasher)
    if ch["pow_response"] >> (64 - cd.fri.proof_of_work_bits) != 0:
        raise VerifyError("proof of work failed")
    if len(pr.query_rounds) != cd.fri.num_query_rounds: raise VerifyError("bad number of query rounds")
    if len(pr.final_poly) != cd.final_poly_len: raise VerifyError("bad final poly length")
    zeta = ch["zeta"]
    gN = gl.primitive_root_of_unity(cd.degree_bits)
    points = [zeta, gl.escale(zeta, gN)]
    batches = fri_batches(cd)
    openings = [pr.constants + pr.plonk_sigmas + pr.wires + pr.plonk_zs + pr.partial_products + pr.quotient_polys,
                pr.plonk_zs_next]
    alpha = ch["fri_alpha"]
    reduced_openings = [gl.reduce_with_powers_ext(o, alpha) for o in openings]
    caps = [vo.constants_sigmas_cap, pr.wires_cap, pr.zs_pp_cap, pr.quotient_cap]
    lde_bits = cd.lde_bits
    w_lde = gl.primitive_root_of_unity(lde_bits)
    for qi, rp in zip(ch["query_indices"], pr.query_rounds):
        x_index = qi % (1 << lde_bits)
        if len(rp.initial) != 4: raise VerifyError("bad initial proof")
        for o, (leaf, sib) in enumerate(rp.initial):
            if len(leaf) != cd.oracle_widths()[o] or len(sib) + cd.fri.cap_height != lde_bits:
                raise VerifyError("bad initial proof shape")
            verify_merkle(H, leaf, x_index, sib, caps[o])
        subgroup_x = gl.GENERATOR * pow(w_lde, gl.reverse_bits(x_index, lde_bits), P) % P
        # fri_combine_initial (fri.go:170-211)
        s = gl.E_ZERO
        for b, point, ro in zip(batches, points, reduced_openings):
            evs = [(rp.initial[o][0][i], 0) for (o, i) in b]
            red = gl.reduce_with_powers_ext(evs, alpha)
            num = gl.esub(red, ro)
            den = gl.esub((subgroup_x, 0), point)
            s = gl.emul(gl.eexp(alpha, len(evs)), s)
            s = gl.eadd(s, gl.ediv(num, den))
        old_eval = s
        if len(rp.steps) != len(cd.reduction_arity_bits): raise VerifyError("bad steps")
        bits_left = lde_bits
        for i, ab in enumerate(cd.reduction_arity_bits):
            st = rp.steps[i]
            if len(st.evals) != (1 << ab): raise VerifyError("bad evals length")
            within = x_index & ((1 << ab) - 1)
            coset_index = x_index >> ab
            if st.evals[within] != old_eval:
                raise VerifyError(f"fri fold consistency failed at step {i}")
            old_eval = compute_evaluation(subgroup_x, within, ab, st.evals, ch["fri_betas"][i])
            flat = []
            for a, b2 in st.evals: flat += [a, b2]
            bits_left -= ab
            if len(st.siblings) + cd.fri.cap_height != bits_left: raise VerifyError("bad step proof shape")
            verify_merkle(H, flat, coset_index, st.siblings, pr.commit_phase_caps[i])
            subgroup_x = pow(subgroup_x, 1 << ab, P)
            x_index = coset_index
        fe = gl.E_ZERO
        for c in reversed(pr.final_poly):
            fe = gl.eadd(gl.emul(fe, (subgroup_x, 0)), c)
        if fe != old_eval:
            raise VerifyError("final polynomial evaluation mismatch")

def verify(cd: CommonData, vo: VerifierOnly, pr: Proof, return_challenges=False):
    if len(pr.public_inputs) != cd.num_public_inputs:
        raise VerifyError("wrong number of public inputs")
    gate_objs = [gate_from_id(s) for s in cd.gates]
    pi_hash = pg.hash_no_pad(pr.public_inputs)
    ch = get_challenges(cd, vo, pr, pi_hash)
    van, zeta_pow_n = eval_vanishing(cd, gate_objs, pr, ch, pi_hash)
    zh = zeta_pow_n - 1
    qd = cd.quotient_degree_factor
    for i, v in enumerate(van):
        chunk = pr.quotient_polys[i * qd:(i + 1) * qd]
        t = Fp2.lift(gl.reduce_with_powers_ext(chunk, zeta_pow_n.tup()))
        if not (v == zh * t):
            raise VerifyError(f"vanishing polynomial identity failed for challenge {i}")
    verify_fri(cd, vo, pr, ch)
    return ch if return_challenges else True
