"""BN254 F_r NTT (next row, SURVEY §8f-3: the FFTs inside groth16.Prove): oracle sanity on CPU, bit-exact parity and
size-independent properties on the GPU."""
import numpy as np
import pytest
from oracle import bn254_fr_ntt as fo

def _limbs(x): return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
def _to_np(vals): return np.array([_limbs(v) for v in vals], dtype=np.uint64).reshape(len(vals), 4)
def _from_np(a): return [sum(int(a[i, k]) << (64 * k) for k in range(4)) for i in range(a.shape[0])]
def _rand(n, seed):
    rng = np.random.default_rng(seed)
    return [int.from_bytes(rng.bytes(32), "little") % fo.R for _ in range(n)]

def test_oracle_domain_parameters_and_naive_dft():
    assert pow(fo.ROOT_2_28, 1 << 28, fo.R) == 1 and pow(fo.ROOT_2_28, 1 << 27, fo.R) == fo.R - 1
    assert pow(fo.GEN, (fo.R - 1) // 2, fo.R) == fo.R - 1                  # 5 is a non-residue (generates the 2-Sylow part)
    for log_n in (0, 1, 2, 4):
        a = _rand(1 << log_n, log_n)
        for cs in (False, True):
            A = fo.ntt(a, coset=cs)
            assert A == fo.naive_dft(a, cs)
            assert fo.ntt(A, inverse=True, coset=cs) == a

@pytest.fixture(scope="module")
def ctx():
    import b200_plonky2
    c = b200_plonky2.Context(0)
    yield c
    c.close()

@pytest.mark.gpu
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 6, 7, 9, 10, 12])
def test_fr_ntt_matches_oracle(ctx, log_n):
    a = _rand(1 << log_n, 100 + log_n)
    if log_n >= 2: a[0] = fo.R - 1; a[1] = 0                               # boundary values
    A = _to_np(a)
    for inv in (False, True):
        for cs in (False, True):
            got = _from_np(ctx.fr_ntt(A, inverse=inv, coset=cs))
            assert got == fo.ntt(a, inverse=inv, coset=cs), (log_n, inv, cs)

@pytest.mark.gpu
def test_fr_ntt_round_trip_and_linearity_at_scale(ctx):
    n = 1 << 18
    rng = np.random.default_rng(5)
    raw = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64); raw[:, 3] &= np.uint64((1 << 60) - 1)     # < 2^252 < r
    A = np.ascontiguousarray(raw)
    for cs in (False, True):
        F = ctx.fr_ntt(A, coset=cs)
        assert (ctx.fr_ntt(F, inverse=True, coset=cs) == A).all()
    # linearity: NTT(a) + NTT(b) == NTT(a + b) on a sample of positions (Python big-int adds)
    raw2 = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64); raw2[:, 3] &= np.uint64((1 << 60) - 1)
    a = _from_np(raw); b = _from_np(raw2)
    S = _to_np([(x + y) % fo.R for x, y in zip(a, b)])
    Fa, Fb, Fs = ctx.fr_ntt(A), ctx.fr_ntt(np.ascontiguousarray(raw2)), ctx.fr_ntt(S)
    idx = rng.integers(0, n, size=512)
    fa, fb, fs = _from_np(Fa[idx]), _from_np(Fb[idx]), _from_np(Fs[idx])
    assert all((x + y) % fo.R == z for x, y, z in zip(fa, fb, fs))
    # F[0] = sum of the inputs; a delta transforms to all ones
    assert _from_np(Fa[:1])[0] == sum(a) % fo.R
    d = np.zeros((n, 4), dtype=np.uint64); d[0, 0] = 1
    assert (ctx.fr_ntt(d) == d[0]).all()
