"""Gate constraint evaluators, written once over an abstract field (oracle.gl.Fp or oracle.gl.Fp2).

TEST INFRASTRUCTURE.  13 standard plonky2 gates restate gnark-plonky2-verifier/plonk/gates/*.go
(EvalUnfiltered of each); the 8 custom u32 gates restate crypto/plonky2_u32/src/gates/*.rs
(eval_unfiltered of each).  Constraint ORDER is part of the spec (alpha powers follow it).
The Fp2 instantiation is pinned by verifying the golden proofs (tests/test_oracle_verifier.py);
the Fp instantiation is the same code and is what the C oracle and the CUDA quotient kernel are
checked against (tests/test_gate_parity.py).
"""
import re
from . import poseidon_gl as pg
from .gl import P, W, primitive_root_of_unity, two_adic_subgroup

# ---- "extension algebra": pairs (c0, c1) of F with X^2 = 7  (quadratic_extension_algebra.go:58-84) ----
def alg_add(a, b): return (a[0] + b[0], a[1] + b[1])
def alg_sub(a, b): return (a[0] - b[0], a[1] - b[1])
def alg_mul(a, b): return (a[0] * b[0] + (a[1] * b[1]) * W, a[0] * b[1] + a[1] * b[0])
def alg_scalar(s, a): return (s * a[0], s * a[1])

class Gate:
    num_constants = 0
    def eval(self, F, consts, wires, pi_hash): raise NotImplementedError

class NoopGate(Gate):                      # noop_gate.go:28
    degree = 0; num_wires = 0
    def id(self): return "NoopGate"
    def num_constraints(self): return 0
    def eval(self, F, c, w, pi): return []

class ConstantGate(Gate):                  # constant_gate.go:57
    degree = 1
    def __init__(self, n): self.n = n; self.num_wires = n; self.num_constants = n
    def id(self): return f"ConstantGate {{ num_consts: {self.n} }}"
    def num_constraints(self): return self.n
    def eval(self, F, c, w, pi): return [c[i] - w[i] for i in range(self.n)]

class PublicInputGate(Gate):               # public_input_gate.go:32
    degree = 1; num_wires = 4
    def id(self): return "PublicInputGate"
    def num_constraints(self): return 4
    def eval(self, F, c, w, pi): return [w[i] - F.const(pi[i]) for i in range(4)]

class ArithmeticGate(Gate):                # arithmetic_gate.go:60
    degree = 3; num_constants = 2
    def __init__(self, n): self.n = n; self.num_wires = 4 * n
    def id(self): return f"ArithmeticGate {{ num_ops: {self.n} }}"
    def num_constraints(self): return self.n
    def eval(self, F, c, w, pi):
        return [w[4 * i + 3] - (w[4 * i] * w[4 * i + 1] * c[0] + w[4 * i + 2] * c[1]) for i in range(self.n)]

class ArithmeticExtensionGate(Gate):       # arithmetic_extension_gate.go:59
    degree = 3; num_constants = 2
    def __init__(self, n): self.n = n; self.num_wires = 8 * n
    def id(self): return f"ArithmeticExtensionGate {{ num_ops: {self.n} }}"
    def num_constraints(self): return 2 * self.n
    def eval(self, F, c, w, pi):
        out = []
        for i in range(self.n):
            m0 = (w[8 * i], w[8 * i + 1]); m1 = (w[8 * i + 2], w[8 * i + 3])
            ad = (w[8 * i + 4], w[8 * i + 5]); o = (w[8 * i + 6], w[8 * i + 7])
            comp = alg_add(alg_scalar(c[1], ad), alg_scalar(c[0], alg_mul(m0, m1)))
            d = alg_sub(o, comp); out += [d[0], d[1]]
        return out

class MulExtensionGate(Gate):              # multiplication_extension_gate.go:55
    degree = 3; num_constants = 1
    def __init__(self, n): self.n = n; self.num_wires = 6 * n
    def id(self): return f"MulExtensionGate {{ num_ops: {self.n} }}"
    def num_constraints(self): return 2 * self.n
    def eval(self, F, c, w, pi):
        out = []
        for i in range(self.n):
            m0 = (w[6 * i], w[6 * i + 1]); m1 = (w[6 * i + 2], w[6 * i + 3]); o = (w[6 * i + 4], w[6 * i + 5])
            d = alg_sub(o, alg_scalar(c[0], alg_mul(m0, m1))); out += [d[0], d[1]]
        return out

class BaseSumGate(Gate):                   # base_sum_gate.go:66
    def __init__(self, num_limbs, base): self.n = num_limbs; self.base = base; self.num_wires = 1 + num_limbs; self.degree = base
    def id(self): return f"BaseSumGate {{ num_limbs: {self.n} }} + Base: {self.base}"
    def num_constraints(self): return 1 + self.n
    def eval(self, F, c, w, pi):
        limbs = [w[1 + i] for i in range(self.n)]
        acc = F.const(0)
        for l in reversed(limbs): acc = acc * self.base + l
        out = [acc - w[0]]
        for l in limbs:
            prod = F.const(1)
            for k in range(self.base): prod = prod * (l - F.const(k))
            out.append(prod)
        return out

class ReducingGate(Gate):                  # reducing_gate.go:77
    degree = 2
    def __init__(self, n): self.n = n; self.num_wires = 6 + n + 2 * (n - 1)
    def id(self): return f"ReducingGate {{ num_coeffs: {self.n} }}"
    def num_constraints(self): return 2 * self.n
    def _acc(self, w, i):
        if i == self.n - 1: return (w[0], w[1])
        s = 6 + self.n + 2 * i
        return (w[s], w[s + 1])
    def eval(self, F, c, w, pi):
        alpha = (w[2], w[3]); acc = (w[4], w[5]); out = []
        for i in range(self.n):
            coeff = (w[6 + i], F.const(0))
            t = alg_sub(alg_add(alg_mul(acc, alpha), coeff), self._acc(w, i))
            out += [t[0], t[1]]; acc = self._acc(w, i)
        return out

class ReducingExtensionGate(Gate):         # reducing_extension_gate.go:77
    degree = 2
    def __init__(self, n): self.n = n; self.num_wires = 6 + 2 * n + 2 * (n - 1)
    def id(self): return f"ReducingExtensionGate {{ num_coeffs: {self.n} }}"
    def num_constraints(self): return 2 * self.n
    def _acc(self, w, i):
        if i == self.n - 1: return (w[0], w[1])
        s = 6 + 2 * self.n + 2 * i
        return (w[s], w[s + 1])
    def eval(self, F, c, w, pi):
        alpha = (w[2], w[3]); acc = (w[4], w[5]); out = []
        for i in range(self.n):
            coeff = (w[6 + 2 * i], w[7 + 2 * i])
            t = alg_sub(alg_add(alg_mul(acc, alpha), coeff), self._acc(w, i))
            out += [t[0], t[1]]; acc = self._acc(w, i)
        return out

class RandomAccessGate(Gate):              # random_access_gate.go:131
    def __init__(self, bits, copies, extra):
        self.bits = bits; self.copies = copies; self.extra = extra; self.vec = 1 << bits
        self.routed = (2 + self.vec) * copies + extra
        self.num_wires = self.routed + copies * bits
        self.num_constants = extra; self.degree = bits + 1
    def id(self): return (f"RandomAccessGate {{ bits: {self.bits}, num_copies: {self.copies}, "
                          f"num_extra_constants: {self.extra}, _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }}<D=2>")
    def num_constraints(self): return self.copies * (self.bits + 2) + self.extra
    def eval(self, F, c, w, pi):
        out = []
        for cp in range(self.copies):
            base = (2 + self.vec) * cp
            idx = w[base]; claimed = w[base + 1]
            items = [w[base + 2 + i] for i in range(self.vec)]
            bits = [w[self.routed + cp * self.bits + i] for i in range(self.bits)]
            for b in bits: out.append(b * b - b)
            rec = F.const(0)
            for b in reversed(bits): rec = rec * 2 + b
            out.append(rec - idx)
            for b in bits:
                items = [items[i] + b * (items[i + 1] - items[i]) for i in range(0, len(items), 2)]
            out.append(items[0] - claimed)
        for i in range(self.extra):
            out.append(c[i] - w[(2 + self.vec) * self.copies + i])
        return out

class PoseidonMdsGate(Gate):               # poseidon_mds_gate.go:76
    degree = 1; num_wires = 48
    def id(self): return "PoseidonMdsGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>"
    def num_constraints(self): return 24
    def eval(self, F, c, w, pi):
        inp = [(w[2 * i], w[2 * i + 1]) for i in range(12)]
        out = []
        for r in range(12):
            acc = (F.const(0), F.const(0))
            for i in range(12):
                acc = alg_add(acc, alg_scalar(F.const(pg.MDS_CIRC[i]), inp[(i + r) % 12]))
            acc = alg_add(acc, alg_scalar(F.const(pg.MDS_DIAG[r]), inp[r]))
            o = (w[24 + 2 * r], w[25 + 2 * r])
            d = alg_sub(o, acc); out += [d[0], d[1]]
        return out

def _sbox(x):
    x2 = x * x; x4 = x2 * x2
    return x4 * (x * x2)

def _mds(F, s):
    out = []
    for r in range(12):
        acc = F.const(0)
        for i in range(12): acc = acc + s[(i + r) % 12] * pg.MDS_CIRC[i]
        acc = acc + s[r] * pg.MDS_DIAG[r]
        out.append(acc)
    return out

class PoseidonGate(Gate):                  # poseidon_gate.go:92
    degree = 7; num_wires = 135
    WIRE_SWAP = 24; START_DELTA = 25; START_FULL_0 = 29; START_PARTIAL = 29 + 36; START_FULL_1 = 29 + 36 + 22
    def id(self): return "PoseidonGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>"
    def num_constraints(self): return 123
    def eval(self, F, c, w, pi):
        out = []
        swap = w[24]
        out.append(swap * (swap - F.const(1)))
        for i in range(4):
            out.append(swap * (w[i + 4] - w[i]) - w[25 + i])
        st = [None] * 12
        for i in range(4):
            st[i] = w[i] + w[25 + i]; st[i + 4] = w[i + 4] - w[25 + i]
        for i in range(8, 12): st[i] = w[i]
        rc = 0
        for r in range(4):
            st = [st[i] + pg.RC[12 * rc + i] for i in range(12)]
            if r != 0:
                for i in range(12):
                    sin = w[self.START_FULL_0 + 12 * (r - 1) + i]
                    out.append(st[i] - sin); st[i] = sin
            st = _mds(F, [_sbox(x) for x in st]); rc += 1
        st = [st[i] + pg.FP_FIRST_RC[i] for i in range(12)]
        res = [F.const(0)] * 12; res[0] = st[0]
        for r in range(1, 12):
            for d in range(1, 12):
                res[d] = res[d] + st[r] * pg.FP_INIT[r - 1][d - 1]
        st = res
        for r in range(22):
            sin = w[self.START_PARTIAL + r]
            out.append(st[0] - sin)
            st[0] = _sbox(sin)
            if r != 21: st[0] = st[0] + pg.FP_RC[r]
            d = st[0] * pg.MDS0TO0
            for i in range(1, 12): d = d + st[i] * pg.FP_W_HATS[r][i - 1]
            st = [d] + [st[i] + st[0] * pg.FP_VS[r][i - 1] for i in range(1, 12)]
        rc += 22
        for r in range(4):
            st = [st[i] + pg.RC[12 * rc + i] for i in range(12)]
            for i in range(12):
                sin = w[self.START_FULL_1 + 12 * r + i]
                out.append(st[i] - sin); st[i] = sin
            st = _mds(F, [_sbox(x) for x in st]); rc += 1
        for i in range(12): out.append(st[i] - w[12 + i])
        return out

class CosetInterpolationGate(Gate):        # coset_interpolation_gate.go:151
    def __init__(self, subgroup_bits, degree, weights):
        self.sb = subgroup_bits; self.degree = degree; self.weights = weights; self.np = 1 << subgroup_bits
        self.n_inter = (self.np - 2) // (degree - 1)
        self.start_inter = 1 + 2 * self.np + 4
        self.num_wires = self.start_inter + 4 * self.n_inter + 2
    def id(self):
        ws = ", ".join(str(x) for x in self.weights)
        return (f"CosetInterpolationGate {{ subgroup_bits: {self.sb}, degree: {self.degree}, barycentric_weights: [{ws}], "
                f"_phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }}<D=2>")
    def num_constraints(self): return 2 + 4 * self.n_inter + 2
    def _partial(self, F, dom, vals, ws, point, ev, prod):
        for x, v, wt in zip(dom, vals, ws):
            term = alg_sub(point, (F.const(x), F.const(0)))
            wv = alg_scalar(F.const(wt), v)
            ev = alg_add(alg_mul(ev, term), alg_mul(wv, prod))
            prod = alg_mul(prod, term)
        return ev, prod
    def eval(self, F, c, w, pi):
        out = []
        shift = w[0]
        pair = lambda s: (w[s], w[s + 1])
        ep = pair(1 + 2 * self.np); sep = pair(self.start_inter + 4 * self.n_inter)
        t = alg_add(alg_scalar(F.const(0) - shift, sep), ep); out += [t[0], t[1]]
        dom = two_adic_subgroup(self.sb)
        vals = [pair(1 + 2 * i) for i in range(self.np)]
        ev, prod = self._partial(F, dom[:self.degree], vals[:self.degree], self.weights[:self.degree], sep,
                                 (F.const(0), F.const(0)), (F.const(1), F.const(0)))
        for i in range(self.n_inter):
            ie = pair(self.start_inter + 2 * i); ip = pair(self.start_inter + 2 * (self.n_inter + i))
            d = alg_sub(ie, ev); out += [d[0], d[1]]
            d = alg_sub(ip, prod); out += [d[0], d[1]]
            s = 1 + (self.degree - 1) * (i + 1); e = min(s + self.degree - 1, self.np)
            ev, prod = self._partial(F, dom[s:e], vals[s:e], self.weights[s:e], sep, ie, ip)
        evv = pair(1 + 2 * self.np + 2)
        d = alg_sub(evv, ev); out += [d[0], d[1]]
        return out

# ---------------- custom u32 gates (crypto/plonky2_u32/src/gates) ----------------
def _range_prod(F, x, n):
    prod = F.const(1)
    for k in range(n): prod = prod * (x - F.const(k))
    return prod
_PH = "_phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField>"

class U32ArithmeticGate(Gate):             # arithmetic_u32.rs:108-167
    degree = 4
    def __init__(self, num_ops): self.n = num_ops; self.num_wires = 38 * num_ops
    @staticmethod
    def num_ops_for(num_wires, num_routed): return min(num_wires // 38, num_routed // 6)
    def id(self): return f"U32ArithmeticGate {{ num_ops: {self.n}, {_PH} }}"
    def num_constraints(self): return self.n * 36
    def eval(self, F, c, w, pi):
        out = []
        for i in range(self.n):
            m0, m1, ad, lo, hi, inv = (w[6 * i + k] for k in range(6))
            comp = m0 * m1 + ad
            out.append((inv * (F.const(0xFFFFFFFF) - hi) - F.const(1)) * lo)
            out.append(hi * F.const(1 << 32) + lo - comp)
            cl = F.const(0); ch = F.const(0)
            for j in reversed(range(32)):
                limb = w[6 * self.n + 32 * i + j]
                out.append(_range_prod(F, limb, 4))
                if j < 16: cl = cl * 4 + limb
                else: ch = ch * 4 + limb
            out.append(cl - lo); out.append(ch - hi)
        return out

class U32AddManyGate(Gate):                # add_many_u32.rs:109-147
    degree = 4
    def __init__(self, num_addends, num_ops):
        self.na = num_addends; self.n = num_ops; self.num_wires = (num_addends + 3 + 18) * num_ops
    @staticmethod
    def num_ops_for(num_addends, num_wires, num_routed):
        return min(num_wires // (num_addends + 3 + 18), num_routed // (num_addends + 3))
    def id(self): return f"U32AddManyGate {{ num_addends: {self.na}, num_ops: {self.n}, {_PH} }}"
    def num_constraints(self): return self.n * 21
    def eval(self, F, c, w, pi):
        out = []; r = self.na + 3
        for i in range(self.n):
            comp = F.const(0)
            for j in range(self.na): comp = comp + w[r * i + j]
            comp = comp + w[r * i + self.na]
            res = w[r * i + self.na + 1]; carry = w[r * i + self.na + 2]
            out.append(carry * F.const(1 << 32) + res - comp)
            cr = F.const(0); cc = F.const(0)
            for j in reversed(range(18)):
                limb = w[r * self.n + 18 * i + j]
                out.append(_range_prod(F, limb, 4))
                if j < 16: cr = cr * 4 + limb
                else: cc = cc * 4 + limb
            out.append(cr - res); out.append(cc - carry)
        return out

class U32SubtractionGate(Gate):            # subtraction_u32.rs:102-137
    degree = 4
    def __init__(self, num_ops): self.n = num_ops; self.num_wires = 21 * num_ops
    @staticmethod
    def num_ops_for(num_wires, num_routed): return min(num_wires // 21, num_routed // 5)
    def id(self): return f"U32SubtractionGate {{ num_ops: {self.n}, {_PH} }}"
    def num_constraints(self): return self.n * 19
    def eval(self, F, c, w, pi):
        out = []
        for i in range(self.n):
            x, y, bi, res, bo = (w[5 * i + k] for k in range(5))
            out.append(res - (x - y - bi + F.const(1 << 32) * bo))
            cl = F.const(0)
            for j in reversed(range(16)):
                limb = w[5 * self.n + 16 * i + j]
                out.append(_range_prod(F, limb, 4))
                cl = cl * 4 + limb
            out.append(cl - res)
            out.append(bo * (F.const(1) - bo))
        return out

class U32RangeCheckGate(Gate):             # range_check_u32.rs:71-93
    degree = 4
    def __init__(self, n): self.n = n; self.num_wires = 17 * n
    def id(self): return f"U32RangeCheckGate {{ num_input_limbs: {self.n}, {_PH} }}"
    def num_constraints(self): return self.n * 17
    def eval(self, F, c, w, pi):
        out = []
        for i in range(self.n):
            aux = [w[self.n + 16 * i + j] for j in range(16)]
            acc = F.const(0)
            for a in reversed(aux): acc = acc * 4 + a
            out.append(acc - w[i])
            for a in aux: out.append(_range_prod(F, a, 4))
        return out

class U32InterleaveGate(Gate):             # interleave_u32.rs:101-138
    degree = 2
    def __init__(self, num_ops): self.n = num_ops; self.num_wires = 34 * num_ops
    @staticmethod
    def num_ops_for(num_wires, num_routed): return min(num_wires // 34, num_routed // 2)
    def id(self): return f"U32InterleaveGate {{ num_ops: {self.n} }}"
    def num_constraints(self): return self.n * 34
    def eval(self, F, c, w, pi):
        out = []
        for i in range(self.n):
            bits = [w[2 * self.n + 32 * i + j] for j in range(32)]
            cx = F.const(0); ci = F.const(0)
            for b in bits: cx = cx * 2 + b; ci = ci * 4 + b      # reduce_with_powers(bits.rev(), B) == big-endian Horner
            out.append(cx - w[2 * i]); out.append(ci - w[2 * i + 1])
            for b in bits: out.append(_range_prod(F, b, 2))
        return out

class _Uninterleave(Gate):
    degree = 2
    def __init__(self, num_ops): self.n = num_ops; self.num_wires = 67 * num_ops
    @staticmethod
    def num_ops_for(num_wires, num_routed): return min(num_wires // 67, num_routed // 3)
    def num_constraints(self): return self.n * 67
    def eval(self, F, c, w, pi):
        out = []
        for i in range(self.n):
            bits = [w[3 * self.n + 64 * i + j] for j in range(64)]
            ci = F.const(0)
            for b in bits: ci = ci * 2 + b
            out.append(ci - w[3 * i])
            ce = F.const(0); co = F.const(0)
            for j in range(32):
                coeff = F.const(self._coeff(j))
                ce = ce + coeff * bits[2 * j]; co = co + coeff * bits[2 * j + 1]
            out.append(ce - w[3 * i + 1]); out.append(co - w[3 * i + 2])
            for b in bits: out.append(_range_prod(F, b, 2))
        return out

class UninterleaveToU32Gate(_Uninterleave):  # uninterleave_to_u32.rs:110-156
    def id(self): return f"UninterleaveToU32Gate {{ num_ops: {self.n} }}"
    def _coeff(self, j): return 1 << (31 - j)

class UninterleaveToB32Gate(_Uninterleave):  # uninterleave_to_b32.rs:114-162
    def id(self): return f"UninterleaveToB32Gate {{ num_ops: {self.n} }}"
    def _coeff(self, j): return 1 << (2 * (31 - j))

class ComparisonGate(Gate):                # comparison.rs:118-195
    def __init__(self, num_bits, num_chunks):
        self.nb = num_bits; self.nc = num_chunks; self.cb = -(-num_bits // num_chunks)
        self.num_wires = 4 + 5 * num_chunks + self.cb + 1; self.degree = 1 << self.cb
    def id(self): return f"ComparisonGate {{ num_bits: {self.nb}, num_chunks: {self.nc}, {_PH} }}<D=2>"
    def num_constraints(self): return 6 + 5 * self.nc + self.cb
    def eval(self, F, c, w, pi):
        nc, cb = self.nc, self.cb
        out = []
        f = [w[4 + i] for i in range(nc)]; s = [w[4 + nc + i] for i in range(nc)]
        fc = F.const(0); sc = F.const(0)
        for x in reversed(f): fc = fc * (1 << cb) + x
        for x in reversed(s): sc = sc * (1 << cb) + x
        out.append(fc - w[0]); out.append(sc - w[1])
        msd = F.const(0)
        for i in range(nc):
            out.append(_range_prod(F, f[i], 1 << cb)); out.append(_range_prod(F, s[i], 1 << cb))
            diff = s[i] - f[i]
            dummy = w[4 + 2 * nc + i]; eq = w[4 + 3 * nc + i]
            out.append(diff * dummy - (F.const(1) - eq)); out.append(eq * diff)
            inter = w[4 + 4 * nc + i]
            out.append(inter - eq * msd)
            msd = inter + (F.const(1) - eq) * diff
        out.append(w[3] - msd)
        bits = [w[4 + 5 * nc + i] for i in range(cb + 1)]
        for b in bits: out.append(b * (F.const(1) - b))
        comb = F.const(0)
        for b in reversed(bits): comb = comb * 2 + b
        out.append(F.const(1 << cb) + w[3] - comb)
        out.append(w[2] - bits[cb])
        return out

_PARSERS = [
    (r"^NoopGate$", lambda m: NoopGate()),
    (r"^ConstantGate \{ num_consts: (\d+) \}$", lambda m: ConstantGate(int(m[1]))),
    (r"^PublicInputGate$", lambda m: PublicInputGate()),
    (r"^ArithmeticGate \{ num_ops: (\d+) \}$", lambda m: ArithmeticGate(int(m[1]))),
    (r"^ArithmeticExtensionGate \{ num_ops: (\d+) \}$", lambda m: ArithmeticExtensionGate(int(m[1]))),
    (r"^MulExtensionGate \{ num_ops: (\d+) \}$", lambda m: MulExtensionGate(int(m[1]))),
    (r"^BaseSumGate \{ num_limbs: (\d+) \} \+ Base: (\d+)$", lambda m: BaseSumGate(int(m[1]), int(m[2]))),
    (r"^ReducingGate \{ num_coeffs: (\d+) \}$", lambda m: ReducingGate(int(m[1]))),
    (r"^ReducingExtensionGate \{ num_coeffs: (\d+) \}$", lambda m: ReducingExtensionGate(int(m[1]))),
    (r"^RandomAccessGate \{ bits: (\d+), num_copies: (\d+), num_extra_constants: (\d+)", lambda m: RandomAccessGate(int(m[1]), int(m[2]), int(m[3]))),
    (r"^PoseidonMdsGate", lambda m: PoseidonMdsGate()),
    (r"^PoseidonGate", lambda m: PoseidonGate()),
    (r"^CosetInterpolationGate \{ subgroup_bits: (\d+), degree: (\d+), barycentric_weights: \[([0-9, ]+)\]",
     lambda m: CosetInterpolationGate(int(m[1]), int(m[2]), [int(x) for x in m[3].split(",")])),
    (r"^U32ArithmeticGate \{ num_ops: (\d+)", lambda m: U32ArithmeticGate(int(m[1]))),
    (r"^U32AddManyGate \{ num_addends: (\d+), num_ops: (\d+)", lambda m: U32AddManyGate(int(m[1]), int(m[2]))),
    (r"^U32SubtractionGate \{ num_ops: (\d+)", lambda m: U32SubtractionGate(int(m[1]))),
    (r"^U32RangeCheckGate \{ num_input_limbs: (\d+)", lambda m: U32RangeCheckGate(int(m[1]))),
    (r"^U32InterleaveGate \{ num_ops: (\d+)", lambda m: U32InterleaveGate(int(m[1]))),
    (r"^UninterleaveToU32Gate \{ num_ops: (\d+)", lambda m: UninterleaveToU32Gate(int(m[1]))),
    (r"^UninterleaveToB32Gate \{ num_ops: (\d+)", lambda m: UninterleaveToB32Gate(int(m[1]))),
    (r"^ComparisonGate \{ num_bits: (\d+), num_chunks: (\d+)", lambda m: ComparisonGate(int(m[1]), int(m[2]))),
]

def gate_from_id(s):
    for pat, mk in _PARSERS:
        m = re.match(pat, s)
        if m: return mk(m)
    raise ValueError(f"unknown gate id {s!r}")

UNUSED_SELECTOR = (1 << 32) - 1

def eval_filtered_constraints(F, cd, gate_objs, consts, wires, pi_hash):
    """evaluate_gates.go:33-105 : sum over gates of filter_g * constraints_g, padded to num_gate_constraints."""
    total = [F.const(0)] * cd.num_gate_constraints
    nsel = cd.num_selectors
    for row, g in enumerate(gate_objs):
        si = cd.selector_indices[row]
        s, e = cd.groups[si]
        sel = consts[si]
        filt = F.const(1)
        for i in range(s, e):
            if i != row: filt = filt * (F.const(i) - sel)
        if nsel > 1:
            filt = filt * (F.const(UNUSED_SELECTOR) - sel)
        cons = g.eval(F, consts[nsel:], wires, pi_hash)
        assert len(cons) == g.num_constraints(), (g.id(), len(cons), g.num_constraints())
        assert len(cons) <= cd.num_gate_constraints
        for i, cv in enumerate(cons):
            total[i] = total[i] + cv * filt
    return total
