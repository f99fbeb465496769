"""Host-side Poseidon-BN254 (Python ints) for the circuit digest of PoseidonBN128GoldilocksConfig circuits.
Tables are read from the generated CUDA header (Montgomery form) and converted back.  Algorithm:
crypto/plonky2_bn128/src/poseidon_bn128.rs:18-108, config.rs:58-69,138-166."""
import os, re
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_RINV = pow(1 << 256, -1, R)

def _tables():
    here = os.path.dirname(os.path.abspath(__file__))
    src = open(os.path.join(here, "..", "csrc", "poseidon_bn254_tables.cuh")).read()
    t = {}
    for name, n, body in re.findall(r"X\((\w+), (\d+), ((?:\{[^}]*\},?)+)\)", src):
        vals = []
        for limbs in re.findall(r"\{([^}]*)\}", body):
            ws = [int(x.rstrip("UL"), 16) for x in limbs.split(",")]
            vals.append(sum(w << (64 * i) for i, w in enumerate(ws)) * _RINV % R)
        assert len(vals) == int(n)
        t[name] = vals
    return t
_T = _tables()
C, S = _T["BN254_C"], _T["BN254_S"]
M = [_T["BN254_M"][4 * i:4 * i + 4] for i in range(4)]; PM = [_T["BN254_P"][4 * i:4 * i + 4] for i in range(4)]

def _e5(x): x2 = x * x % R; return x2 * x2 % R * x % R
def _mix(s, m): return [sum(m[j][i] * s[j] for j in range(4)) % R for i in range(4)]
def permute(s):
    s = [(s[i] + C[i]) % R for i in range(4)]
    for r in range(3): s = _mix([(_e5(s[i]) + C[(r + 1) * 4 + i]) % R for i in range(4)], M)
    s = _mix([(_e5(s[i]) + C[16 + i]) % R for i in range(4)], PM)
    for r in range(56):
        s0 = (_e5(s[0]) + C[20 + r]) % R
        n0 = (S[7 * r] * s0 + sum(S[7 * r + j] * s[j] for j in range(1, 4))) % R
        s = [n0] + [(s[k] + s0 * S[7 * r + 3 + k]) % R for k in range(1, 4)]
    for r in range(3): s = _mix([(_e5(s[i]) + C[76 + r * 4 + i]) % R for i in range(4)], M)
    return _mix([_e5(x) for x in s], M)
def hash_no_pad(xs):
    s = [0, 0, 0, 0]
    for i in range(0, len(xs), 9):
        ch = xs[i:i + 9]
        for j in range(0, len(ch), 3): s[j // 3 + 1] = sum(int(x) << (64 * k) for k, x in enumerate(ch[j:j + 3]))
        s = permute(s)
    return s[0]
def to_vec(h):
    b = int(h).to_bytes(32, "little")
    return [int.from_bytes(b[i:i + 7], "little") for i in range(0, 32, 7)]
def circuit_digest(cap_words, degree_bits):
    """cap_words: [2^cap_height][4] u64 limbs -> digest as 4 u64 limbs."""
    x = [1]
    while (len(x) + 1) % 9 != 0: x.append(0)
    x.append(1)
    parts = []
    for h in cap_words: parts += to_vec(sum(int(h[i]) << (64 * i) for i in range(4)))
    parts += to_vec(hash_no_pad(x)) + [degree_bits]
    d = hash_no_pad(parts)
    return [(d >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
