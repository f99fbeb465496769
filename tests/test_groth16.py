"""groth16.Prove as a pipeline (b200_groth16_prove: H-polynomial by seven F_r transforms, five MSMs, blinding; SURVEY 8 f3,
gnark-plonky2-verifier/cmd/web-api.go:77) against a proving key built from KNOWN toxic waste, so that the discrete logarithm of
every proof element is known in closed form:
    Ar  = [alpha + A(tau) + r delta] G1        Bs = [beta + B(tau) + s delta] G2
    Krs = [ sum_priv w_i (beta A_i + alpha B_i + C_i)(tau) / delta + H(tau) Z(tau) / delta + s a + r b - r s delta ] G1
with A(tau) = sum_i w_i A_i(tau) etc. and H = (A B - C) / Z exact because the R1CS is satisfied.  This pins the whole pipeline
(domain conventions of the transforms, Z on the coset, query layouts, blinding algebra) without a pairing."""
import numpy as np
import pytest
from oracle import bn254_curve as ec
from oracle import bn254_fr_ntt as frn

R = ec.R
def _limbs(x, n=4): return [(int(x) >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]
def _g1(k): p = ec.mul(k % R, ec.G); return [0] * 8 if p is None else _limbs(p[0]) + _limbs(p[1])
def _g2(k): p = ec.g2_mul(k % R, ec.G2); return [0] * 16 if p is None else _limbs(p[0][0]) + _limbs(p[0][1]) + _limbs(p[1][0]) + _limbs(p[1][1])
def _pt1(l): x = sum(int(l[i]) << (64 * i) for i in range(4)); y = sum(int(l[4 + i]) << (64 * i) for i in range(4)); return None if x == 0 and y == 0 else (x, y)
def _pt2(l):
    v = [sum(int(l[4 * k + i]) << (64 * i) for i in range(4)) for k in range(4)]
    return None if not any(v) else ((v[0], v[1]), (v[2], v[3]))

def _instance(log_n, num_public, seed):
    """multiplication-gate R1CS: wire 0 = 1, then inputs, then one output wire per constraint"""
    import random
    rng = random.Random(seed)
    n = 1 << log_n
    ncons = n - 3                                            # fewer constraints than the domain: zero padding as in gnark
    nin = 5
    nw = 1 + nin + ncons
    w = [1] + [rng.randrange(R) for _ in range(nin)] + [0] * ncons
    A = [dict() for _ in range(n)]; B = [dict() for _ in range(n)]; C = [dict() for _ in range(n)]
    for j in range(ncons):
        prev = list(range(0, 1 + nin + j))
        for M in (A[j], B[j]):
            for i in rng.sample(prev, min(3, len(prev))): M[i] = rng.randrange(1, R) if rng.random() < 0.5 else rng.choice([1, 2, R - 1])
        a = sum(v * w[i] for i, v in A[j].items()) % R; b = sum(v * w[i] for i, v in B[j].items()) % R
        o = 1 + nin + j; w[o] = a * b % R; C[j][o] = 1
    av = [sum(v * w[i] for i, v in A[j].items()) % R for j in range(n)]
    bv = [sum(v * w[i] for i, v in B[j].items()) % R for j in range(n)]
    cv = [sum(v * w[i] for i, v in C[j].items()) % R for j in range(n)]
    assert all(x * y % R == z for x, y, z in zip(av, bv, cv))
    return n, nw, w, A, B, C, av, bv, cv

@pytest.mark.gpu
@pytest.mark.parametrize("log_n,num_public", [(4, 3), (7, 1)])
def test_groth16_prove_matches_toxic_waste_closed_form(log_n, num_public):
    import random, b200_plonky2 as bp
    n, nw, w, A, B, C, av, bv, cv = _instance(log_n, num_public, 100 + log_n)
    rng = random.Random(7)
    tau, alpha, beta, delta = (rng.randrange(2, R) for _ in range(4))
    omega = frn.root_of_unity(log_n)
    zt = (pow(tau, n, R) - 1) % R
    ninv = pow(n, -1, R)
    L = [zt * pow(omega, j, R) % R * ninv % R * pow((tau - pow(omega, j, R)) % R, -1, R) % R for j in range(n)]       # Lagrange basis at tau
    At = [0] * nw; Bt = [0] * nw; Ct = [0] * nw
    for j in range(n):
        for i, v in A[j].items(): At[i] = (At[i] + v * L[j]) % R
        for i, v in B[j].items(): Bt[i] = (Bt[i] + v * L[j]) % R
        for i, v in C[j].items(): Ct[i] = (Ct[i] + v * L[j]) % R
    dinv = pow(delta, -1, R)
    Kdl = [(beta * At[i] + alpha * Bt[i] + Ct[i]) * dinv % R for i in range(num_public, nw)]
    Zdl = [pow(tau, k, R) * zt % R * dinv % R for k in range(n - 1)]
    ctx = bp.Context(0)
    pk = ctx.groth16_pk_upload(log_n, num_public, [_g1(x) for x in At], [_g1(x) for x in Bt], [_g2(x) for x in Bt], [_g1(x) for x in Kdl], [_g1(x) for x in Zdl],
                               _g1(alpha), _g1(beta), _g1(delta), _g2(beta), _g2(delta))
    r, s = rng.randrange(R), rng.randrange(R)
    lim = lambda xs: np.array([_limbs(x) for x in xs], dtype=np.uint64)
    proof = ctx.groth16_prove(pk, lim(w), lim(av), lim(bv), lim(cv), r, s)
    a_t = sum(x * y for x, y in zip(w, At)) % R; b_t = sum(x * y for x, y in zip(w, Bt)) % R; c_t = sum(x * y for x, y in zip(w, Ct)) % R
    h_t = (a_t * b_t - c_t) * pow(zt, -1, R) % R
    a_dl = (alpha + a_t + r * delta) % R
    b_dl = (beta + b_t + s * delta) % R
    k_dl = (sum(w[num_public + i] * Kdl[i] for i in range(nw - num_public)) + h_t * zt % R * dinv + s * a_dl + r * b_dl - r * s % R * delta) % R
    assert _pt1(proof[0:8]) == ec.mul(a_dl, ec.G)
    assert _pt2(proof[8:24]) == ec.g2_mul(b_dl, ec.G2)
    assert _pt1(proof[24:32]) == ec.mul(k_dl, ec.G)
    # a second proof with other blinding factors from the same resident key
    proof2 = ctx.groth16_prove(pk, lim(w), lim(av), lim(bv), lim(cv), 0, 0)
    assert _pt1(proof2[0:8]) == ec.mul((alpha + a_t) % R, ec.G)
    ctx.groth16_pk_free(pk); ctx.close()
