"""Goldilocks field F_p, p = 2^64 - 2^32 + 1, and its quadratic extension F_p[X]/(X^2-7).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates
  gnark-plonky2-verifier/goldilocks/base.go:33-39,445-471    (modulus, generator 7, 2-adic root)
  gnark-plonky2-verifier/goldilocks/quadratic_extension.go:59-73 (W = 7, mul/inv formulas)
with Python ints (slow, exact).
"""
P = (1 << 64) - (1 << 32) + 1
GENERATOR = 7                         # multiplicative generator == coset shift  (base.go:35)
TWO_ADICITY = 32
POWER_OF_TWO_GENERATOR = 1753635133440165772   # order 2^32            (base.go:37)
W = 7                                 # X^2 = W                          (quadratic_extension.go:59)
DTH_ROOT = 18446744069414584320       # = p-1 : frobenius of X

def add(a, b): return (a + b) % P
def sub(a, b): return (a - b) % P
def mul(a, b): return (a * b) % P
def neg(a): return (-a) % P
def inv(a):
    assert a % P != 0, "inverse of zero"
    return pow(a, P - 2, P)
def exp(a, e): return pow(a, e, P)

def primitive_root_of_unity(n_log):
    """base.go:445-455 : root^(2^(32-n_log))."""
    assert n_log <= TWO_ADICITY
    return pow(POWER_OF_TWO_GENERATOR, 1 << (TWO_ADICITY - n_log), P)

def two_adic_subgroup(n_log):
    g = primitive_root_of_unity(n_log)
    out, x = [], 1
    for _ in range(1 << n_log):
        out.append(x); x = x * g % P
    return out

def reverse_bits(x, nbits):
    r = 0
    for _ in range(nbits):
        r = (r << 1) | (x & 1); x >>= 1
    return r

def batch_inv(xs):
    n = len(xs)
    pref = [1] * (n + 1)
    for i, x in enumerate(xs):
        pref[i + 1] = pref[i] * x % P
    acc = inv(pref[n])
    out = [0] * n
    for i in range(n - 1, -1, -1):
        out[i] = acc * pref[i] % P
        acc = acc * xs[i] % P
    return out

# ---- quadratic extension: elements are tuples (a0, a1) = a0 + a1*X ----
E_ZERO = (0, 0)
E_ONE = (1, 0)
def e(a): return (a % P, 0)
def eadd(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def esub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def eneg(a): return ((-a[0]) % P, (-a[1]) % P)
def emul(a, b):
    return ((a[0] * b[0] + W * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def escale(a, s): return (a[0] * s % P, a[1] * s % P)
def einv(a):
    # 1/(a0 + a1 X) = (a0 - a1 X)/(a0^2 - 7 a1^2)
    d = inv((a[0] * a[0] - W * a[1] * a[1]) % P)
    return (a[0] * d % P, (-a[1]) * d % P)
def ediv(a, b): return emul(a, einv(b))
def eexp(a, n):
    r = E_ONE
    while n:
        if n & 1: r = emul(r, a)
        a = emul(a, a); n >>= 1
    return r
def eexp_pow2(a, k):
    for _ in range(k): a = emul(a, a)
    return a

def reduce_with_powers_ext(terms, alpha):
    """sum_i terms[i] * alpha^i  (Horner from the back).  base.go / quadratic_extension.go ReduceWithPowers."""
    acc = E_ZERO
    for t in reversed(terms):
        acc = eadd(emul(acc, alpha), t)
    return acc


class Fp:
    """Operator-overloaded base-field element so gate evaluators can be written once (oracle/gates.py)."""
    __slots__ = ("v",)
    def __init__(self, v): self.v = v % P
    @staticmethod
    def lift(x): return x if isinstance(x, Fp) else Fp(x)
    def __add__(self, o): return Fp(self.v + Fp.lift(o).v)
    __radd__ = __add__
    def __sub__(self, o): return Fp(self.v - Fp.lift(o).v)
    def __rsub__(self, o): return Fp(Fp.lift(o).v - self.v)
    def __mul__(self, o): return Fp(self.v * Fp.lift(o).v)
    __rmul__ = __mul__
    def __neg__(self): return Fp(-self.v)
    def __eq__(self, o): return self.v == Fp.lift(o).v
    def __hash__(self): return hash(self.v)
    def is_zero(self): return self.v == 0
    def __repr__(self): return f"Fp({self.v})"
    @staticmethod
    def const(c): return Fp(c)


class Fp2:
    """Operator-overloaded F_p^2 element; ints/Fp promote to (x, 0)."""
    __slots__ = ("a", "b")
    def __init__(self, a, b=0): self.a = a % P; self.b = b % P
    @staticmethod
    def lift(x):
        if isinstance(x, Fp2): return x
        if isinstance(x, Fp): return Fp2(x.v, 0)
        if isinstance(x, tuple): return Fp2(x[0], x[1])
        return Fp2(x, 0)
    def __add__(self, o): o = Fp2.lift(o); return Fp2(self.a + o.a, self.b + o.b)
    __radd__ = __add__
    def __sub__(self, o): o = Fp2.lift(o); return Fp2(self.a - o.a, self.b - o.b)
    def __rsub__(self, o): o = Fp2.lift(o); return Fp2(o.a - self.a, o.b - self.b)
    def __mul__(self, o):
        o = Fp2.lift(o)
        return Fp2(self.a * o.a + W * self.b * o.b, self.a * o.b + self.b * o.a)
    __rmul__ = __mul__
    def __neg__(self): return Fp2(-self.a, -self.b)
    def __eq__(self, o): o = Fp2.lift(o); return self.a == o.a and self.b == o.b
    def __hash__(self): return hash((self.a, self.b))
    def is_zero(self): return self.a == 0 and self.b == 0
    def tup(self): return (self.a, self.b)
    def __repr__(self): return f"Fp2({self.a},{self.b})"
    @staticmethod
    def const(c): return Fp2(c, 0)
