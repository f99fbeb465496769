"""Poseidon over Goldilocks (width 12, rate 8, x^7, 4+22+4 rounds) + sponge helpers.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
  gnark-plonky2-verifier/poseidon/goldilocks.go:30-37   permutation = full, partial(fast), full
  goldilocks.go:41-86   hash_n_to_m_no_pad (overwrite-mode absorb, squeeze from state[0..8])
  goldilocks.go:92-115  full rounds / fast partial rounds
  goldilocks.go:251-331 partial-first-constant layer, mds_partial_layer_init, mds_partial_layer_fast
Parameter tables: oracle/data/poseidon_gl_constants.json (extracted by oracle/tools/extract_constants.py
from poseidon/goldilocks_constants.go).  hash_or_noop / two_to_one follow upstream plonky2
(hash/hashing.rs, hash/poseidon.rs @ 2244a9d; not in tree) as summarised in SURVEY.md A.1.
"""
import json, os
from .gl import P

_D = json.load(open(os.path.join(os.path.dirname(__file__), "data", "poseidon_gl_constants.json")))
RC = _D["ALL_ROUND_CONSTANTS"]
MDS_CIRC = _D["MDS_MATRIX_CIRC"]
MDS_DIAG = _D["MDS_MATRIX_DIAG"]
MDS0TO0 = _D["MDS0TO0"][0]
FP_FIRST_RC = _D["FAST_PARTIAL_FIRST_ROUND_CONSTANT"]
FP_RC = _D["FAST_PARTIAL_ROUND_CONSTANTS"]
FP_VS = [_D["FAST_PARTIAL_ROUND_VS"][11 * r:11 * r + 11] for r in range(22)]
FP_W_HATS = [_D["FAST_PARTIAL_ROUND_W_HATS"][11 * r:11 * r + 11] for r in range(22)]
FP_INIT = [_D["FAST_PARTIAL_ROUND_INITIAL_MATRIX"][11 * r:11 * r + 11] for r in range(11)]

WIDTH, RATE, HALF_FULL, N_PARTIAL = 12, 8, 4, 22

def _sbox(x):
    x2 = x * x % P
    x4 = x2 * x2 % P
    return x4 * x2 % P * x % P

def mds_layer(s):
    out = []
    for r in range(12):
        acc = 0
        for i in range(12):
            acc += s[(i + r) % 12] * MDS_CIRC[i]
        acc += s[r] * MDS_DIAG[r]
        out.append(acc % P)
    return out

def _full_rounds(s, rc_off):
    for r in range(HALF_FULL):
        s = [(s[i] + RC[rc_off + 12 * r + i]) % P for i in range(12)]
        s = [_sbox(x) for x in s]
        s = mds_layer(s)
    return s

def _partial_rounds(s):
    s = [(s[i] + FP_FIRST_RC[i]) % P for i in range(12)]
    # mds_partial_layer_init (goldilocks.go:273-300)
    res = [0] * 12
    res[0] = s[0]
    for r in range(1, 12):
        for d in range(1, 12):
            res[d] += s[r] * FP_INIT[r - 1][d - 1]
    s = [x % P for x in res]
    for r in range(N_PARTIAL):
        s[0] = (_sbox(s[0]) + FP_RC[r]) % P
        d = s[0] * MDS0TO0
        for i in range(1, 12):
            d += s[i] * FP_W_HATS[r][i - 1]
        ns = [d % P]
        for i in range(1, 12):
            ns.append((s[i] + s[0] * FP_VS[r][i - 1]) % P)
        s = ns
    return s

def permute(state):
    s = [x % P for x in state]
    assert len(s) == 12
    s = _full_rounds(s, 0)
    s = _partial_rounds(s)
    s = _full_rounds(s, 12 * (HALF_FULL + N_PARTIAL))
    return s

def permute_naive(state):
    """Textbook form (all 30 rounds with ALL_ROUND_CONSTANTS + full MDS); cross-check of the fast form."""
    s = [x % P for x in state]
    for r in range(30):
        s = [(s[i] + RC[12 * r + i]) % P for i in range(12)]
        if r < 4 or r >= 26:
            s = [_sbox(x) for x in s]
        else:
            s[0] = _sbox(s[0])
        s = mds_layer(s)
    return s

def hash_n_to_m_no_pad(inputs, m):
    s = [0] * 12
    for i in range(0, len(inputs), RATE):
        chunk = inputs[i:i + RATE]
        for j, x in enumerate(chunk):
            s[j] = x % P
        s = permute(s)
    out = []
    while True:
        for i in range(RATE):
            out.append(s[i])
            if len(out) == m:
                return out
        s = permute(s)

def hash_no_pad(inputs):
    return hash_n_to_m_no_pad(inputs, 4)

def hash_or_noop(inputs):
    if len(inputs) <= 4:
        return [x % P for x in inputs] + [0] * (4 - len(inputs))
    return hash_no_pad(inputs)

def two_to_one(l, r):
    return permute(list(l) + list(r) + [0, 0, 0, 0])[:4]
