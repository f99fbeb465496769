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


# ---------------------------------------------------------------------------------------------------------------
# G2: the sextic twist  y^2 = x^3 + 3/(9+u)  over F_q2 = F_q[u]/(u^2+1)  (the B-query MSM of groth16.Prove runs here).
# Elements of F_q2 are pairs (a0, a1) = a0 + a1 u; points are ((x0, x1), (y0, y1)) or None.
# ---------------------------------------------------------------------------------------------------------------
def f2_add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
def f2_sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
    return (a[0] * d % Q, -a[1] * d % Q)
def f2_scal(k, a): return (k * a[0] % Q, k * a[1] % Q)

B2 = f2_mul((3, 0), f2_inv((9, 1)))
G2 = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
       11559732032986387107991004021392285783925812861821192530917403151452391805634),
      (8495653923123431417604973247489272438418190587263600148770280649306958101930,
       4082367875863433681332203403145435568316851327593401208105741076214120093531))

def g2_add(p, q):
    if p is None: return q
    if q is None: return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if f2_add(y1, y2) == (0, 0): return None
        lam = f2_mul(f2_scal(3, f2_mul(x1, x1)), f2_inv(f2_scal(2, y1)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))

def g2_mul(k, p):
    r = None
    while k:
        if k & 1: r = g2_add(r, p)
        p = g2_add(p, p); k >>= 1
    return r

def g2_msm(points, scalars):
    acc = None
    for p, s in zip(points, scalars):
        acc = g2_add(acc, g2_mul(s % R, p))
    return acc

def g2_on_curve(p):
    if p is None: return True
    x, y = p
    return f2_sub(f2_mul(y, y), f2_add(f2_mul(f2_mul(x, x), x), B2)) == (0, 0)
