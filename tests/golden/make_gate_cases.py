#!/usr/bin/env python3
"""Fixtures from the reference's own custom-gate unit tests (crypto/plonky2_u32/src/gates/*.rs `test_gate_constraint*`,
`test_canonicity`): each test builds the gate's wire row from integer inputs with a `get_wires` recipe (the witness the gate's
SimpleGenerator would produce) and asserts that eval_unfiltered is all-zero (or, for the negative tests, is not).  The
reference draws its inputs from OsRng; here the same recipes run on seeded inputs and the rows are written out, so that the
oracle evaluators, the C oracle and the product's device evaluators can be held to the reference's own expectations without a
Rust toolchain.  Recipes (file:line in /root/reference/crypto/plonky2_u32/src/gates/):
  arithmetic_u32.rs:487-532 get_wires, :536-571 test_gate_constraint, :573-607 test_canonicity (non-canonical output must fail)
  add_many_u32.rs:417-495   get_wires + test_gate_constraint (NUM_ADDENDS = 10, 3 ops)
  subtraction_u32.rs:401-475 get_wires + test_gate_constraint (3 ops)
  range_check_u32.rs:262-329 get_wires, test_gate_constraint_good (u32 limbs), test_gate_constraint_bad (u64 limbs: must fail)
  comparison.rs:612-725     get_wires + test_gate_constraint (num_bits 40, num_chunks 5)
Usage: python tests/golden/make_gate_cases.py  (writes tests/golden/gate_cases.json; deterministic)."""
import json, os, random
P = (1 << 64) - (1 << 32) + 1
PH = "_phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField>"
rng = random.Random(20260924)
cases = []

def limbs(val, bits, num):
    out = []
    for _ in range(num): out.append(val % (1 << bits)); val >>= bits
    return out

# ---- U32ArithmeticGate, num_ops 3 (arithmetic_u32.rs:487-532) ----
def arith_wires(m0s, m1s, adds):
    v0, v1 = [], []
    for m0, m1, a in zip(m0s, m1s, adds):
        out = m0 * m1 + a                       # u64 arithmetic in the reference; inputs are u32 so there is no wrap
        lo, hi = out & 0xFFFFFFFF, out >> 32
        diff = (0xFFFFFFFF - hi) % P
        inv = pow(diff, P - 2, P) if diff else 0        # F::inverse / zero when hi == u32::MAX (test_canonicity's case)
        v0 += [m0 % P, m1 % P, a % P, lo % P, hi % P, inv]
        v1 += limbs(out, 2, 32)
    return v0 + v1
gid = f"U32ArithmeticGate {{ num_ops: 3, {PH} }}"
for t in range(4):
    cases.append(dict(gate=gid, src="arithmetic_u32.rs:536-571 test_gate_constraint", satisfied=True,
                      wires=arith_wires([rng.getrandbits(32) for _ in range(3)], [rng.getrandbits(32) for _ in range(3)], [rng.getrandbits(32) for _ in range(3)])))
cases.append(dict(gate=gid, src="arithmetic_u32.rs:573-607 test_canonicity", satisfied=False,
                  wires=arith_wires([0] * 3, [0] * 3, [0xFFFFFFFF00000001] * 3)))
cases.append(dict(gate=gid, src="arithmetic_u32.rs extremes (u32::MAX operands)", satisfied=True,
                  wires=arith_wires([0xFFFFFFFF, 0, 1], [0xFFFFFFFF, 0xFFFFFFFF, 1], [0xFFFFFFFF, 0, 0])))

# ---- U32AddManyGate, 10 addends, 3 ops (add_many_u32.rs:423-465) ----
def addmany_wires(addends, carries):
    v0, v1 = [], []
    for adds, ca in zip(addends, carries):
        out = sum(adds) + ca
        res, car = out & 0xFFFFFFFF, out >> 32
        v0 += list(adds) + [ca, res, car]
        v1 += limbs(res, 2, 16) + limbs(car, 2, 2)
    return v0 + v1
gid = f"U32AddManyGate {{ num_addends: 10, num_ops: 3, {PH} }}"
for t in range(4):
    cases.append(dict(gate=gid, src="add_many_u32.rs:417-495 test_gate_constraint", satisfied=True,
                      wires=addmany_wires([[rng.getrandbits(32) for _ in range(10)] for _ in range(3)], [rng.getrandbits(32) % 8 for _ in range(3)])))
w = addmany_wires([[0xFFFFFFFF] * 10 for _ in range(3)], [7, 0, 1]); cases.append(dict(gate=gid, src="add_many_u32.rs extremes", satisfied=True, wires=w))
w = list(cases[-2]["wires"]); w[11] = (w[11] + 1) % P; cases.append(dict(gate=gid, src="add_many_u32.rs: result off by one", satisfied=False, wires=w))

# ---- U32SubtractionGate, 3 ops (subtraction_u32.rs:407-451) ----
def sub_wires(xs, ys, bs):
    v0, v1 = [], []
    for x, y, b in zip(xs, ys, bs):
        r0 = (x - y - b) % P
        bo = 1 if r0 > (1 << 32) else 0
        res = (r0 + (1 << 32) * bo) % P
        v0 += [x, y, b, res, bo]
        v1 += limbs(res, 2, 16)
    return v0 + v1
gid = f"U32SubtractionGate {{ num_ops: 3, {PH} }}"
for t in range(4):
    cases.append(dict(gate=gid, src="subtraction_u32.rs:401-475 test_gate_constraint", satisfied=True,
                      wires=sub_wires([rng.getrandbits(32) for _ in range(3)], [rng.getrandbits(32) for _ in range(3)], [rng.getrandbits(1) for _ in range(3)])))
cases.append(dict(gate=gid, src="subtraction_u32.rs extremes", satisfied=True, wires=sub_wires([0, 0xFFFFFFFF, 5], [0xFFFFFFFF, 0, 5], [1, 1, 0])))
w = sub_wires([9, 8, 7], [1, 2, 3], [0, 0, 0]); w[4] = 2; cases.append(dict(gate=gid, src="subtraction_u32.rs: borrow not boolean", satisfied=False, wires=w))

# ---- U32RangeCheckGate, 8 input limbs (range_check_u32.rs:262-329) ----
def rc_wires(ins):
    v = []
    for x in ins: v += limbs(x, 2, 16)
    return [x % P for x in ins] + v
gid = f"U32RangeCheckGate {{ num_input_limbs: 8, {PH} }}"
for t in range(4):
    cases.append(dict(gate=gid, src="range_check_u32.rs:316-321 test_gate_constraint_good", satisfied=True, wires=rc_wires([rng.getrandbits(32) for _ in range(8)])))
for t in range(2):
    ins = [rng.getrandbits(64) | (1 << 40) for _ in range(8)]
    cases.append(dict(gate=gid, src="range_check_u32.rs:323-329 test_gate_constraint_bad (should_panic)", satisfied=False, wires=rc_wires(ins)))

# ---- ComparisonGate, 40 bits, 5 chunks (comparison.rs:620-690) ----
def cmp_wires(f, s, num_bits=40, num_chunks=5):
    cb = num_bits // num_chunks
    res = 1 if f <= s else 0
    fc, sc = limbs(f, cb, num_chunks), limbs(s, cb, num_chunks)
    eq = [1 if a == b else 0 for a, b in zip(fc, sc)]
    dum = [1 if a == b else pow((b - a) % P, P - 2, P) for a, b in zip(fc, sc)]
    msd, inter = 0, []
    for a, b in zip(fc, sc):
        if a != b: msd = (b - a) % P; inter.append(0)
        else: inter.append(msd)
    bits = limbs((1 << cb) + msd, 1, cb + 1)
    return [f, s, res, msd] + fc + sc + dum + eq + inter + bits
gid = f"ComparisonGate {{ num_bits: 40, num_chunks: 5, {PH} }}<D=2>"
for t in range(4):
    a = rng.randrange(1 << 39); b = rng.randrange(a, 1 << 39)
    cases.append(dict(gate=gid, src="comparison.rs:612-725 test_gate_constraint (first <= second)", satisfied=True, wires=cmp_wires(a, b)))
a = rng.randrange(1 << 39); cases.append(dict(gate=gid, src="comparison.rs: equal inputs", satisfied=True, wires=cmp_wires(a, a)))
a = rng.randrange(1, 1 << 39); b = rng.randrange(0, a)
w = cmp_wires(a, b); w[2] = 1; cases.append(dict(gate=gid, src="comparison.rs: first > second but result_bool claims <=", satisfied=False, wires=w))

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gate_cases.json")
json.dump(cases, open(out, "w"), indent=0)
print(len(cases), "cases ->", out)
