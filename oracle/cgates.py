"""Gate-id strings -> orc_gate descriptors for the C oracle (oracle/c/gates.h).  TEST INFRASTRUCTURE."""
import ctypes, re
import numpy as np
from . import gates as pg

KINDS = ["NOOP", "CONSTANT", "PUBLIC_INPUT", "ARITHMETIC", "ARITHMETIC_EXT", "MUL_EXT", "BASE_SUM", "REDUCING",
         "REDUCING_EXT", "RANDOM_ACCESS", "POSEIDON_MDS", "POSEIDON", "COSET_INTERPOLATION", "U32_ARITHMETIC",
         "U32_ADD_MANY", "U32_SUBTRACTION", "U32_RANGE_CHECK", "U32_INTERLEAVE", "UNINTERLEAVE_U32",
         "UNINTERLEAVE_B32", "COMPARISON"]
K = {k: i for i, k in enumerate(KINDS)}

class OrcGate(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_uint32), ("p", ctypes.c_uint32 * 4), ("selector_index", ctypes.c_uint32),
                ("group_start", ctypes.c_uint32), ("group_end", ctypes.c_uint32), ("row", ctypes.c_uint32),
                ("num_constraints", ctypes.c_uint32), ("extra", ctypes.c_void_p)]

def describe(gid):
    """-> (kind, params[4], extra list)"""
    g = pg.gate_from_id(gid)
    t = type(g).__name__
    if t == "NoopGate": return K["NOOP"], [0, 0, 0, 0], None
    if t == "ConstantGate": return K["CONSTANT"], [g.n, 0, 0, 0], None
    if t == "PublicInputGate": return K["PUBLIC_INPUT"], [0, 0, 0, 0], None
    if t == "ArithmeticGate": return K["ARITHMETIC"], [g.n, 0, 0, 0], None
    if t == "ArithmeticExtensionGate": return K["ARITHMETIC_EXT"], [g.n, 0, 0, 0], None
    if t == "MulExtensionGate": return K["MUL_EXT"], [g.n, 0, 0, 0], None
    if t == "BaseSumGate": return K["BASE_SUM"], [g.n, g.base, 0, 0], None
    if t == "ReducingGate": return K["REDUCING"], [g.n, 0, 0, 0], None
    if t == "ReducingExtensionGate": return K["REDUCING_EXT"], [g.n, 0, 0, 0], None
    if t == "RandomAccessGate": return K["RANDOM_ACCESS"], [g.bits, g.copies, g.extra, 0], None
    if t == "PoseidonMdsGate": return K["POSEIDON_MDS"], [0, 0, 0, 0], None
    if t == "PoseidonGate": return K["POSEIDON"], [0, 0, 0, 0], None
    if t == "CosetInterpolationGate": return K["COSET_INTERPOLATION"], [g.sb, g.degree, 0, 0], list(g.weights)
    if t == "U32ArithmeticGate": return K["U32_ARITHMETIC"], [g.n, 0, 0, 0], None
    if t == "U32AddManyGate": return K["U32_ADD_MANY"], [g.na, g.n, 0, 0], None
    if t == "U32SubtractionGate": return K["U32_SUBTRACTION"], [g.n, 0, 0, 0], None
    if t == "U32RangeCheckGate": return K["U32_RANGE_CHECK"], [g.n, 0, 0, 0], None
    if t == "U32InterleaveGate": return K["U32_INTERLEAVE"], [g.n, 0, 0, 0], None
    if t == "UninterleaveToU32Gate": return K["UNINTERLEAVE_U32"], [g.n, 0, 0, 0], None
    if t == "UninterleaveToB32Gate": return K["UNINTERLEAVE_B32"], [g.n, 0, 0, 0], None
    if t == "ComparisonGate": return K["COMPARISON"], [g.nb, g.nc, 0, 0], None
    raise ValueError(gid)

def make_gate_array(gate_ids, selector_indices, groups):
    arr = (OrcGate * len(gate_ids))()
    keep = []
    for i, gid in enumerate(gate_ids):
        kind, p, extra = describe(gid)
        arr[i].kind = kind
        for j in range(4): arr[i].p[j] = p[j]
        si = selector_indices[i]
        arr[i].selector_index = si; arr[i].group_start, arr[i].group_end = groups[si]; arr[i].row = i
        arr[i].num_constraints = pg.gate_from_id(gid).num_constraints()
        if extra is not None:
            e = np.array(extra, dtype=np.uint64); keep.append(e); arr[i].extra = e.ctypes.data
    return arr, keep
