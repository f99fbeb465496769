"""BN254 G1 MSM (next row): oracle sanity on CPU, bit-exact parity + linearity properties on the GPU."""
import numpy as np
import pytest
from oracle import bn254_curve as ec

def test_oracle_known_points():
    # 2G as used by the EVM ecAdd precompile test vectors
    assert ec.add(ec.G, ec.G) == (0x030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3,
                                  0x15ed738c0e0a7c92e7845f96b2ae9c0a68a6a449e3538fc7ff3ebf7a5a18a2c4)
    assert ec.mul(ec.R, ec.G) is None and ec.on_curve(ec.mul(123456789, ec.G))
    assert ec.add(ec.mul(5, ec.G), ec.mul(ec.R - 5, ec.G)) is None

def _limbs(x, n=4): return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]

def _rand_instance(n, seed, small=False):
    rng = np.random.default_rng(seed)
    ks = [int(x) for x in rng.integers(1, 1 << 62, size=n)]
    pts = [ec.mul(k, ec.G) for k in ks]
    if small: scal = [int(x) for x in rng.integers(0, 1 << 20, size=n)]
    else: scal = [int.from_bytes(rng.bytes(32), "little") % ec.R for _ in range(n)]
    P = np.array([_limbs(p[0]) + _limbs(p[1]) for p in pts], dtype=np.uint64)
    S = np.array([_limbs(s) for s in scal], dtype=np.uint64)
    return pts, scal, ks, P, S

@pytest.fixture(scope="module")
def ctx():
    import b200_plonky2
    c = b200_plonky2.Context(0)
    yield c
    c.close()

@pytest.mark.gpu
@pytest.mark.parametrize("n,small", [(1, False), (7, False), (300, False), (300, True), (2048, False)])
def test_msm_matches_oracle(ctx, n, small):
    pts, scal, ks, P, S = _rand_instance(n, 10 + n, small)
    # via the discrete logs: sum s_i k_i G  (fast) and, for small n, the naive oracle MSM as well
    expect = ec.mul(sum(s * k for s, k in zip(scal, ks)) % ec.R, ec.G)
    if n <= 300: assert ec.msm(pts, scal) == expect
    assert ctx.msm_g1(P, S) == expect

@pytest.mark.gpu
def test_msm_edge_cases(ctx):
    pts, scal, ks, P, S = _rand_instance(64, 5)
    S0 = S.copy(); S0[:] = 0
    assert ctx.msm_g1(P, S0) is None                                   # all-zero scalars -> infinity
    P2 = P.copy(); P2[3] = 0                                           # a point at infinity contributes nothing
    exp = ec.msm([p for i, p in enumerate(pts) if i != 3], [s for i, s in enumerate(scal) if i != 3])
    assert ctx.msm_g1(P2, S) == exp
    # the same point many times with equal digits exercises the doubling / cancellation branches of the bucket adds
    Pm = np.repeat(P[:1], 50, axis=0); Sm = np.repeat(S[:1], 50, axis=0)
    assert ctx.msm_g1(Pm, Sm) == ec.mul(50 * scal[0] * ks[0] % ec.R, ec.G)
    Sn = Sm.copy(); neg = (ec.R - scal[0]) % ec.R
    Sn[25:] = np.array(_limbs(neg), dtype=np.uint64)
    assert ctx.msm_g1(Pm, Sn) is None                                  # 25 s P + 25 (-s) P = O
    Smax = S.copy(); Smax[:] = np.array(_limbs(ec.R - 1), dtype=np.uint64)
    assert ctx.msm_g1(P, Smax) == ec.mul((ec.R - 1) * sum(ks) % ec.R, ec.G)

@pytest.mark.gpu
def test_msm_linearity_at_scale(ctx):
    """size-independent property at 2^17 points: MSM(P, a) + MSM(P, b) == MSM(P, a + b); points = multiples of G built by
    repeated addition so the expected value is also known in closed form."""
    n = 1 << 17
    rng = np.random.default_rng(3)
    base = [ec.mul(k, ec.G) for k in range(1, 65)]
    idx = rng.integers(0, 64, size=n)
    P = np.array([_limbs(base[i][0]) + _limbs(base[i][1]) for i in idx], dtype=np.uint64)
    a = [int.from_bytes(rng.bytes(31), "little") for _ in range(n)]
    b = [int.from_bytes(rng.bytes(31), "little") for _ in range(n)]
    A = np.array([_limbs(x) for x in a], dtype=np.uint64); B = np.array([_limbs(x) for x in b], dtype=np.uint64)
    AB = np.array([_limbs((x + y) % ec.R) for x, y in zip(a, b)], dtype=np.uint64)
    ra, rb, rab = ctx.msm_g1(P, A), ctx.msm_g1(P, B), ctx.msm_g1(P, AB)
    assert ec.add(ra, rb) == rab
    assert ra == ec.mul(sum(x * (int(i) + 1) for x, i in zip(a, idx)) % ec.R, ec.G)


# ---------------------------------------------------------------- G2 ----------------------------------------------------------------
def test_oracle_g2_generator():
    assert ec.g2_on_curve(ec.G2) and ec.g2_mul(ec.R, ec.G2) is None
    assert ec.g2_on_curve(ec.g2_mul(987654321, ec.G2))
    assert ec.g2_add(ec.g2_mul(7, ec.G2), ec.g2_mul(ec.R - 7, ec.G2)) is None
    # distributivity through the discrete log, which the GPU tests rely on
    assert ec.g2_msm([ec.g2_mul(3, ec.G2), ec.g2_mul(5, ec.G2)], [11, 13]) == ec.g2_mul(3 * 11 + 5 * 13, ec.G2)

def _g2_row(p): return _limbs(p[0][0]) + _limbs(p[0][1]) + _limbs(p[1][0]) + _limbs(p[1][1])

def _rand_g2(n, seed, small=False):
    rng = np.random.default_rng(seed)
    ks = [int(x) for x in rng.integers(1, 1 << 62, size=n)]
    pts = [ec.g2_mul(k, ec.G2) for k in ks]
    if small: scal = [int(x) for x in rng.integers(0, 1 << 20, size=n)]
    else: scal = [int.from_bytes(rng.bytes(32), "little") % ec.R for _ in range(n)]
    P = np.array([_g2_row(p) for p in pts], dtype=np.uint64)
    S = np.array([_limbs(s) for s in scal], dtype=np.uint64)
    return pts, scal, ks, P, S

@pytest.mark.gpu
@pytest.mark.parametrize("n,small", [(1, False), (5, False), (200, False), (200, True)])
def test_msm_g2_matches_oracle(ctx, n, small):
    pts, scal, ks, P, S = _rand_g2(n, 40 + n, small)
    expect = ec.g2_mul(sum(s * k for s, k in zip(scal, ks)) % ec.R, ec.G2)
    if n <= 5: assert ec.g2_msm(pts, scal) == expect
    assert ctx.msm_g2(P, S) == expect

@pytest.mark.gpu
def test_msm_g2_edge_cases_and_linearity(ctx):
    pts, scal, ks, P, S = _rand_g2(32, 6)
    S0 = S.copy(); S0[:] = 0
    assert ctx.msm_g2(P, S0) is None
    P2 = P.copy(); P2[3] = 0
    assert ctx.msm_g2(P2, S) == ec.g2_mul(sum(s * k for i, (s, k) in enumerate(zip(scal, ks)) if i != 3) % ec.R, ec.G2)
    Pm = np.repeat(P[:1], 40, axis=0); Sm = np.repeat(S[:1], 40, axis=0)
    assert ctx.msm_g2(Pm, Sm) == ec.g2_mul(40 * scal[0] * ks[0] % ec.R, ec.G2)
    Sn = Sm.copy(); Sn[20:] = np.array(_limbs((ec.R - scal[0]) % ec.R), dtype=np.uint64)
    assert ctx.msm_g2(Pm, Sn) is None
    # 2^14 points drawn from 32 distinct ones: closed form through the discrete logs, and additivity in the scalars
    n = 1 << 14
    rng = np.random.default_rng(9)
    idx = rng.integers(0, 32, size=n)
    Pn = P[idx]
    a = [int.from_bytes(rng.bytes(31), "little") for _ in range(n)]
    b = [int.from_bytes(rng.bytes(31), "little") for _ in range(n)]
    A = np.array([_limbs(x) for x in a], dtype=np.uint64); B = np.array([_limbs(x) for x in b], dtype=np.uint64)
    AB = np.array([_limbs((x + y) % ec.R) for x, y in zip(a, b)], dtype=np.uint64)
    ra, rb, rab = ctx.msm_g2(Pn, A), ctx.msm_g2(Pn, B), ctx.msm_g2(Pn, AB)
    assert ec.g2_add(ra, rb) == rab
    assert ra == ec.g2_mul(sum(x * ks[int(i)] for x, i in zip(a, idx)) % ec.R, ec.G2)


@pytest.mark.gpu
@pytest.mark.parametrize("group", ["g1", "g2"])
def test_msm_groth16_like_skewed_scalars(ctx, group):
    """Groth16 witness-like scalars: most wires are 0 or 1, many are small, few are full-size.  With 2^17 points the buckets of
    digit 1 (and of the all-equal high digits of -1) hold tens of thousands of points each: they must take the sliced
    hot-bucket path (k_hot_partial / k_hot_reduce) and still give the exact group element (closed form via discrete logs)."""
    n = 1 << 17
    rng = np.random.default_rng(21)
    if group == "g1":
        base = [ec.mul(k, ec.G) for k in range(1, 65)]; rows = [_limbs(p[0]) + _limbs(p[1]) for p in base]
    else:
        base = [ec.g2_mul(k, ec.G2) for k in range(1, 65)]; rows = [_g2_row(p) for p in base]
    idx = rng.integers(0, 64, size=n)
    P = np.array(rows, dtype=np.uint64)[idx]
    kind = rng.random(n)
    scal = []
    for i in range(n):
        if kind[i] < 0.45: scal.append(0)
        elif kind[i] < 0.80: scal.append(1)
        elif kind[i] < 0.90: scal.append(int(rng.integers(0, 1 << 16)))
        elif kind[i] < 0.95: scal.append(ec.R - 1)                      # -1: every window digit equal (0xFFFF..., hot in all windows)
        else: scal.append(int.from_bytes(rng.bytes(32), "little") % ec.R)
    S = np.array([_limbs(s) for s in scal], dtype=np.uint64)
    dl = sum(s * (int(i) + 1) for s, i in zip(scal, idx)) % ec.R
    if group == "g1": assert ctx.msm_g1(P, S) == ec.mul(dl, ec.G)
    else: assert ctx.msm_g2(P, S) == ec.g2_mul(dl, ec.G2)
