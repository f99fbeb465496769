"""BN254 (alt_bn128) G1 group law and a naive multi-scalar multiplication with Python integers.
TEST INFRASTRUCTURE: the oracle for b200_msm_g1 (gnark-crypto ecc/bn254 is not vendored, the curve
y^2 = x^3 + 3 over F_q and the generator (1, 2) are the published parameters gnark-plonky2-verifier/go.mod:8-9 pins)."""
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
G = (1, 2)

def add(p, q):
    if p is None: return q
    if q is None: return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if (y1 + y2) % Q == 0: return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, Q) % Q
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    return (x3, (lam * (x1 - x3) - y1) % Q)

def mul(k, p):
    r = None
    while k:
        if k & 1: r = add(r, p)
        p = add(p, p); k >>= 1
    return r

def msm(points, scalars):
    acc = None
    for p, s in zip(points, scalars):
        acc = add(acc, mul(s % R, p))
    return acc

def on_curve(p):
    return p is None or (p[1] * p[1] - p[0] ** 3 - 3) % Q == 0
