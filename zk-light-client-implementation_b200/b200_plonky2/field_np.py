"""Vectorised Goldilocks arithmetic on numpy uint64 arrays (host-side witness / circuit construction only)."""
import numpy as np

P = (1 << 64) - (1 << 32) + 1
_P = np.uint64(P)
_EPS = np.uint64(0xFFFFFFFF)
_M32 = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)

def canon(a):
    a = np.asarray(a, dtype=np.uint64)
    return np.where(a >= _P, a - _P, a)

def add(a, b):
    a = np.asarray(a, dtype=np.uint64); b = np.asarray(b, dtype=np.uint64)
    with np.errstate(over="ignore"):
        s = a + b
        s = np.where(s < a, s + _EPS, s)
    return canon(s)

def sub(a, b):
    a = np.asarray(a, dtype=np.uint64); b = np.asarray(b, dtype=np.uint64)
    with np.errstate(over="ignore"):
        d = a - b
        d = np.where(a < b, d - _EPS, d)
    return d

def neg(a):
    return sub(np.zeros_like(np.asarray(a, dtype=np.uint64)), a)

def _mul_wide(a, b):
    a = np.asarray(a, dtype=np.uint64); b = np.asarray(b, dtype=np.uint64)
    with np.errstate(over="ignore"):
        al, ah = a & _M32, a >> _S32
        bl, bh = b & _M32, b >> _S32
        ll = al * bl; lh = al * bh; hl = ah * bl; hh = ah * bh
        mid = (ll >> _S32) + (lh & _M32) + (hl & _M32)
        lo = (ll & _M32) | ((mid & _M32) << _S32)
        hi = hh + (lh >> _S32) + (hl >> _S32) + (mid >> _S32)
    return lo, hi

def mul(a, b):
    lo, hi = _mul_wide(a, b)
    with np.errstate(over="ignore"):
        hh, hl = hi >> _S32, hi & _M32
        t0 = lo - hh
        t0 = np.where(lo < hh, t0 - _EPS, t0)
        t1 = hl * _EPS
        r = t0 + t1
        r = np.where(r < t1, r + _EPS, r)
    return canon(r)

def pow_(a, e):
    a = np.asarray(a, dtype=np.uint64)
    r = np.ones_like(a)
    while e:
        if e & 1: r = mul(r, a)
        a = mul(a, a); e >>= 1
    return r

def inv(a):
    return pow_(a, P - 2)

def powers(base, n):
    """[1, base, base^2, ...] (n entries) by doubling."""
    out = np.ones(n, dtype=np.uint64)
    k = 1
    b = np.uint64(base % P)
    cur = np.array([b], dtype=np.uint64)      # base^k
    while k < n:
        m = min(k, n - k)
        out[k:k + m] = mul(out[:m], np.broadcast_to(cur, (m,)))
        cur = mul(cur, cur); k *= 2
    return out

def root_of_unity(n_log):
    r = 1753635133440165772
    for _ in range(n_log, 32): r = r * r % P
    return r
