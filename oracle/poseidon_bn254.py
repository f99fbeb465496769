"""Poseidon over BN254 Fr (t=4, 8 full + 56 partial, x^5, iden3 optimised form) and the
PoseidonBN128Hash hasher used by PoseidonBN128GoldilocksConfig.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
  crypto/plonky2_bn128/src/poseidon_bn128.rs:18-108  permutation (ark, full, partial with S table, mix M / P)
  crypto/plonky2_bn128/src/config.rs:138-166  hash_no_pad (3 GL elements -> one Fr, LE bytes)
  config.rs:178-189 hash_or_noop (<= 3 elements packed), :191-199 two_to_one, :58-69 to_vec (7-byte chunks)
Tables: oracle/data/poseidon_bn254_constants.json.
"""
import json, os

R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_D = json.load(open(os.path.join(os.path.dirname(__file__), "data", "poseidon_bn254_constants.json")))
C = [int(x) for x in _D["C"]]
S = [int(x) for x in _D["S"]]
M = [[int(x) for x in row] for row in _D["M"]]
PM = [[int(x) for x in row] for row in _D["P"]]
WIDTH, FULL, PARTIAL = 4, 8, 56

def _exp5(x):
    x2 = x * x % R
    return x2 * x2 % R * x % R

def _ark(s, it): return [(s[i] + C[it + i]) % R for i in range(4)]
def _mix(s, m): return [sum(m[j][i] * s[j] for j in range(4)) % R for i in range(4)]

def _full(s, first):
    for i in range(FULL // 2 - 1):
        s = [_exp5(x) for x in s]
        s = _ark(s, (i + 1) * 4 if first else (FULL // 2 + 1) * 4 + PARTIAL + i * 4)
        s = _mix(s, M)
    s = [_exp5(x) for x in s]
    if first:
        s = _ark(s, (FULL // 2) * 4)
        s = _mix(s, PM)
    else:
        s = _mix(s, M)
    return s

def _partial(s):
    for i in range(PARTIAL):
        s0 = (_exp5(s[0]) + C[(FULL // 2 + 1) * 4 + i]) % R
        s = [s0] + s[1:]
        new0 = sum(S[7 * i + j] * s[j] for j in range(4)) % R
        for k in range(1, 4):
            s[k] = (s[k] + s[0] * S[7 * i + 4 + k - 1]) % R
        s[0] = new0
    return s

def permute(state):
    s = _ark([x % R for x in state], 0)
    s = _full(s, True)
    s = _partial(s)
    s = _full(s, False)
    return s

def hash_no_pad(inputs):
    s = [0, 0, 0, 0]
    for i in range(0, len(inputs), 9):
        chunk = inputs[i:i + 9]
        for j in range(0, len(chunk), 3):
            v = 0
            for k, x in enumerate(chunk[j:j + 3]):
                v += x << (64 * k)
            assert v < R
            s[j // 3 + 1] = v
        s = permute(s)
    return s[0]

def hash_or_noop(inputs):
    if len(inputs) <= 3:
        v = 0
        for k, x in enumerate(inputs):
            v += x << (64 * k)
        return v
    return hash_no_pad(inputs)

def two_to_one(l, r):
    return permute([0, 0, l, r])[0]

def to_vec(h):
    """32 LE bytes in 7-byte chunks -> 5 Goldilocks elements (config.rs:58-69)."""
    b = h.to_bytes(32, "little")
    return [int.from_bytes(b[i:i + 7], "little") for i in range(0, 32, 7)]
