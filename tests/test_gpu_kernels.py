"""GPU parity (bit-exact) of the inner-loop kernels against the oracle, through the C ABI."""
import numpy as np
import pytest
from tests.util import rand_field, P, have_gpu

pytestmark = pytest.mark.gpu

@pytest.fixture(scope="module")
def ctx():
    import b200_plonky2
    c = b200_plonky2.Context(0)
    yield c
    c.close()

def test_poseidon_permutation_matches_oracle(ctx):
    from oracle import poseidon_gl as pg, cbind
    st = rand_field((300, 12), 1)
    st[0] = 0                                            # G1 zero-state KAT
    st[1] = P - 1
    out = ctx.poseidon_permute(st)
    assert [int(x) for x in out[0]] == pg.permute([0] * 12)
    for i in range(300):
        assert np.array_equal(out[i], cbind.poseidon_permute(st[i])), i

@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 5, 7, 9, 10, 12, 13, 14, 16, 17])
def test_ntt_roundtrip_and_oracle(ctx, log_n):
    from oracle import cbind
    cols = 3 if log_n < 16 else 2
    v = rand_field((cols, 1 << log_n), 100 + log_n)
    f = ctx.ntt_batch(v, log_n, inverse=False)
    assert np.array_equal(f, cbind.ntt_batch(v, log_n, False))
    b = ctx.ntt_batch(f, log_n, inverse=True)
    assert np.array_equal(b, v)

def test_ntt_edge_values(ctx):
    from oracle import cbind
    log_n = 10
    v = np.zeros((4, 1 << log_n), dtype=np.uint64)
    v[1, :] = P - 1; v[2, 0] = 1; v[3, ::2] = P - 1
    assert np.array_equal(ctx.ntt_batch(v, log_n), cbind.ntt_batch(v, log_n))

@pytest.mark.parametrize("log_n,cols", [(3, 5), (5, 1), (6, 4), (8, 20), (10, 9), (12, 135), (13, 16), (14, 3)])
def test_lde_commit_matches_oracle(ctx, log_n, cols):
    from oracle import cbind
    v = rand_field((cols, 1 << log_n), 7 * log_n + cols)
    cap, lde, co = ctx.lde_commit(v, log_n, 3, min(4, log_n + 3))
    oco, olde, odg, ocap = cbind.commit_values(v, log_n, 3, min(4, log_n + 3))
    assert np.array_equal(co, oco), "coefficients"
    assert np.array_equal(lde, olde), "LDE values"
    assert np.array_equal(cap, ocap), "Merkle cap"

@pytest.mark.parametrize("log_L,cols", [(4, 1), (4, 4), (5, 5), (6, 8), (7, 9), (10, 16), (10, 85), (11, 234)])
def test_poseidon_merkle_matches_oracle(ctx, log_L, cols):
    from oracle import cbind
    d = rand_field((cols, 1 << log_L), 31 * log_L + cols)
    cap, dg = ctx.poseidon_merkle(d, log_L, min(4, log_L), want_digests=True)
    odg = cbind.merkle_colmajor(d, log_L, min(4, log_L))
    assert np.array_equal(dg, odg)
    assert np.array_equal(cap.reshape(-1), odg[-cap.size:])

def test_lde_is_low_degree_extension_property(ctx):
    """size-independent property at a larger size: LDE restricted to the original subgroup cosets is linear, and
    the iNTT of the coset-shifted LDE block reproduces the coefficients."""
    log_n = 15
    a = rand_field((2, 1 << log_n), 5); b = rand_field((2, 1 << log_n), 6)
    s = (a.astype(object) + b.astype(object)) % P
    s = np.array(s, dtype=np.uint64)
    _, la, _ = ctx.lde_commit(a, log_n, 3, 4, want_coeffs=False)
    _, lb, _ = ctx.lde_commit(b, log_n, 3, 4, want_coeffs=False)
    _, ls, _ = ctx.lde_commit(s, log_n, 3, 4, want_coeffs=False)
    ref = np.array((la.astype(object) + lb.astype(object)) % P, dtype=np.uint64)
    assert np.array_equal(ls, ref)


@pytest.mark.parametrize("log_n", [18, 20, 21, 22, 24])
def test_ntt_large_sizes_three_pass_path(ctx, log_n):
    """BASELINE config 5 sizes (2^16..2^24): transforms above 2^20 take three kernel passes.  Exact vs the oracle on one
    column plus the size-independent round trip."""
    from oracle import cbind
    v = rand_field((1, 1 << log_n), 900 + log_n)
    f = ctx.ntt_batch(v, log_n, inverse=False)
    assert np.array_equal(f, cbind.ntt_batch(v, log_n, False))
    assert np.array_equal(ctx.ntt_batch(f, log_n, inverse=True), v)

def test_ntt_linearity_at_scale(ctx):
    log_n = 19
    a = rand_field((2, 1 << log_n), 1); b = rand_field((2, 1 << log_n), 2)
    s = np.array((a.astype(object) + b.astype(object)) % P, dtype=np.uint64)
    fa, fb, fs = ctx.ntt_batch(a, log_n), ctx.ntt_batch(b, log_n), ctx.ntt_batch(s, log_n)
    assert np.array_equal(fs, np.array((fa.astype(object) + fb.astype(object)) % P, dtype=np.uint64))

@pytest.mark.parametrize("log_len,ab", [(8, 4), (12, 4), (7, 3), (6, 2), (5, 1)])
def test_fri_fold_matches_verifier_interpolation(ctx, log_len, ab):
    """b200_fri_fold == the verifier's compute_evaluation (gnark-plonky2-verifier/fri/fri.go:244-316) on every coset."""
    import ctypes
    from b200_plonky2 import lib, c_void_p
    from oracle import gl, verifier
    n = 1 << log_len
    vals = rand_field((n, 2), 77 + log_len)
    beta = (123456789012345, 987654321098765); shift = pow(7, 1 << 4, gl.P)
    out = np.zeros((n >> ab, 2), dtype=np.uint64)
    b = (ctypes.c_uint64 * 2)(*beta)
    rc = lib.b200_fri_fold(ctx.h, vals.ctypes.data_as(c_void_p), log_len, ab, b, ctypes.c_uint64(shift), out.ctypes.data_as(c_void_p))
    assert rc == 0
    w = gl.primitive_root_of_unity(log_len)
    for i in range(0, n >> ab, max(1, (n >> ab) // 16)):
        evals = [(int(vals[(i << ab) + k][0]), int(vals[(i << ab) + k][1])) for k in range(1 << ab)]
        x_index = i << ab                                   # first point of the coset
        x = shift * pow(w, gl.reverse_bits(x_index, log_len), gl.P) % gl.P
        exp = verifier.compute_evaluation(x, 0, ab, evals, beta)
        assert (int(out[i][0]), int(out[i][1])) == exp
