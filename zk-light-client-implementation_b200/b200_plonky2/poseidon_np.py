"""Vectorised (numpy) Poseidon-Goldilocks permutation with the S-box-input trace that PoseidonGate exposes as wires
(host-side witness generation only; algorithm: plonky2::hash::poseidon as mirrored by
gnark-plonky2-verifier/poseidon/goldilocks.go:30-37,92-115,251-331 and plonk/gates/poseidon_gate.go:92-181)."""
import os, re
import numpy as np
from . import field_np as F

def _tables():
    # parse the generated CUDA table header (single source of the constants inside the product)
    here = os.path.dirname(os.path.abspath(__file__))
    src = open(os.path.join(here, "..", "csrc", "poseidon_tables.cuh")).read()
    t = {}
    for name, n, body in re.findall(r"X\((\w+), (\d+), ([0-9a-fxULL,]+)\)", src):
        t[name] = np.array([int(x.rstrip("UL"), 16) for x in body.split(",")], dtype=np.uint64)
        assert len(t[name]) == int(n)
    return t
_T = _tables()
RC = _T["POSEIDON_RC"]; FP_FIRST = _T["POSEIDON_FP_FIRST_RC"]; FP_RC = _T["POSEIDON_FP_RC"]
FP_VS = _T["POSEIDON_FP_VS"].reshape(22, 11); FP_W = _T["POSEIDON_FP_W_HATS"].reshape(22, 11); FP_INIT = _T["POSEIDON_FP_INIT"].reshape(11, 11)
CIRC = np.array([17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20], dtype=np.uint64)

def _sbox(x):
    x2 = F.mul(x, x); x4 = F.mul(x2, x2)
    return F.mul(x4, F.mul(x2, x))

def _mds(s):
    out = np.zeros_like(s)
    for r in range(12):
        acc = np.zeros(s.shape[0], np.uint64)
        for i in range(12):
            acc = F.add(acc, F.mul(s[:, (i + r) % 12], np.full(s.shape[0], CIRC[i], np.uint64)))
        if r == 0: acc = F.add(acc, F.mul(s[:, 0], np.full(s.shape[0], 8, np.uint64)))
        out[:, r] = acc
    return out

def permute_trace(state):
    """state [cnt][12] -> (out, full0_sbox_in[4], partial_sbox_in[22], full1_sbox_in[4])"""
    s = state.copy(); cnt = s.shape[0]
    full0, partial, full1 = [], [], []
    rc = 0
    for r in range(4):
        s = F.add(s, np.broadcast_to(RC[12 * rc:12 * rc + 12], s.shape)); full0.append(s.copy())
        s = _mds(_sbox(s)); rc += 1
    s = F.add(s, np.broadcast_to(FP_FIRST, s.shape))
    res = np.zeros_like(s); res[:, 0] = s[:, 0]
    for d in range(1, 12):
        acc = np.zeros(cnt, np.uint64)
        for r in range(1, 12): acc = F.add(acc, F.mul(s[:, r], np.full(cnt, FP_INIT[r - 1, d - 1], np.uint64)))
        res[:, d] = acc
    s = res
    for r in range(22):
        partial.append(s[:, 0].copy())
        x = F.add(_sbox(s[:, 0]), np.full(cnt, FP_RC[r], np.uint64))
        d = F.mul(x, np.full(cnt, 25, np.uint64))
        for i in range(1, 12): d = F.add(d, F.mul(s[:, i], np.full(cnt, FP_W[r, i - 1], np.uint64)))
        ns = s.copy()
        for i in range(1, 12): ns[:, i] = F.add(s[:, i], F.mul(x, np.full(cnt, FP_VS[r, i - 1], np.uint64)))
        ns[:, 0] = d; s = ns
    rc += 22
    for r in range(4):
        s = F.add(s, np.broadcast_to(RC[12 * rc:12 * rc + 12], s.shape)); full1.append(s.copy())
        s = _mds(_sbox(s)); rc += 1
    return s, full0, partial, full1

def permute(state): return permute_trace(state)[0]

def hash_no_pad(inputs):
    s = np.zeros((1, 12), np.uint64)
    for i in range(0, len(inputs), 8):
        ch = inputs[i:i + 8]
        s[0, :len(ch)] = np.array(ch, dtype=np.uint64)
        s = permute(s)
    return [int(x) for x in s[0, :4]]
