"""C-oracle base-field gate evaluators == the generic Python evaluators (which are pinned by the golden proofs),
on random rows, for all 21 gate kinds; plus: every synthetic workload row satisfies its gate constraints."""
import ctypes, random
import numpy as np
import pytest
from oracle import gates as pg, gl, cgates, cbind
from oracle.gl import Fp

COSET = ("CosetInterpolationGate { subgroup_bits: 4, degree: 6, barycentric_weights: [17293822565076172801, "
         "18374686475376656385, 18446744069413535745, 281474976645120, 17592186044416, 256, 18446744000695107601, "
         "18446744065119617025, 1152921504338411520, 72057594037927936, 1048576, 18446462594437939201, "
         "18446726477228539905, 18446744069414584065, 68719476720, 4294967296], _phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField> }<D=2>")
PH = "_phantom: PhantomData<plonky2_field::goldilocks_field::GoldilocksField>"
GATE_IDS = [
    "NoopGate", "ConstantGate { num_consts: 2 }", "PublicInputGate", "ArithmeticGate { num_ops: 20 }",
    "ArithmeticExtensionGate { num_ops: 10 }", "MulExtensionGate { num_ops: 13 }", "BaseSumGate { num_limbs: 63 } + Base: 2",
    "ReducingGate { num_coeffs: 43 }", "ReducingExtensionGate { num_coeffs: 32 }",
    f"RandomAccessGate {{ bits: 4, num_copies: 4, num_extra_constants: 2, {PH} }}<D=2>",
    "PoseidonMdsGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>",
    "PoseidonGate(PhantomData<plonky2_field::goldilocks_field::GoldilocksField>)<WIDTH=12>", COSET,
    f"U32ArithmeticGate {{ num_ops: 6, {PH} }}", f"U32AddManyGate {{ num_addends: 5, num_ops: 9, {PH} }}",
    f"U32SubtractionGate {{ num_ops: 11, {PH} }}", f"U32RangeCheckGate {{ num_input_limbs: 8, {PH} }}",
    "U32InterleaveGate { num_ops: 6 }", "UninterleaveToU32Gate { num_ops: 3 }", "UninterleaveToB32Gate { num_ops: 3 }",
    f"ComparisonGate {{ num_bits: 32, num_chunks: 16, {PH} }}<D=2>",
]

@pytest.mark.parametrize("gid", GATE_IDS)
def test_c_gate_matches_python(gid):
    lib = cbind.lib()
    g = pg.gate_from_id(gid)
    assert g.id() == gid                       # id() round-trips the reference's Debug string
    arr, keep = cgates.make_gate_array([gid], [0], [(0, 1)])
    rng = random.Random(hash(gid) & 0xFFFF)
    for trial in range(3):
        wires = [rng.randrange(gl.P) for _ in range(240)]
        if trial == 1: wires = [rng.randrange(4) for _ in range(240)]     # small values hit the range-check factors
        consts = [rng.randrange(gl.P) for _ in range(4)]
        pi = [rng.randrange(gl.P) for _ in range(4)]
        ref = [x.v for x in g.eval(Fp, [Fp(c) for c in consts], [Fp(w) for w in wires], pi)]
        w = np.array(wires, dtype=np.uint64); c = np.array(consts, dtype=np.uint64); p = np.array(pi, dtype=np.uint64)
        out = np.zeros(512, dtype=np.uint64)
        k = lib.orc_eval_gate(ctypes.byref(arr[0]), c.ctypes.data_as(ctypes.c_void_p), w.ctypes.data_as(ctypes.c_void_p),
                              p.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
        assert k == len(ref) == g.num_constraints()
        assert [int(x) for x in out[:k]] == ref

@pytest.mark.parametrize("kind,db", [("sig", 8), ("rec", 6), ("sha", 7)])
def test_workload_rows_satisfy_gates(kind, db):
    from b200_plonky2 import workloads
    from oracle import poseidon_gl as pgl, proof_io as io
    cd = workloads.build(kind, db, seed=3)
    c = cd.common
    gate_objs = [pg.gate_from_id(s) for s in c["gates"]]
    lib = cbind.lib()
    arr, keep = cgates.make_gate_array(c["gates"], c["selector_indices"], c["groups"])
    pih = np.array(pgl.hash_no_pad(cd.public_inputs), dtype=np.uint64)
    tot = np.zeros(c["num_gate_constraints"], dtype=np.uint64); scr = np.zeros(512, dtype=np.uint64)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for row in range(cd.n):
        cs = np.ascontiguousarray(cd.constants_sigmas[:c["num_constants"], row]); ws = np.ascontiguousarray(cd.wires[:, row])
        lib.orc_eval_filtered(arr, len(c["gates"]), len(c["groups"]), c["num_gate_constraints"], P(cs), P(ws), P(pih), P(tot), P(scr))
        assert not tot.any(), f"row {row} violates constraints {np.nonzero(tot)[0][:4]}"
    # the copy constraints hold: sigma is a permutation of the identity labels k_i * w^row
    R = c["num_routed_wires"]; n = cd.n
    from b200_plonky2 import field_np as F
    ids = F.mul(np.array(c["k_is"], dtype=np.uint64)[:, None], F.powers(F.root_of_unity(db), n)[None, :])
    sig = cd.constants_sigmas[c["num_constants"]:]
    assert np.array_equal(np.sort(ids.reshape(-1)), np.sort(sig.reshape(-1)))
    assert (sig != ids).any()                  # non-trivial permutation (there are real copy constraints)
