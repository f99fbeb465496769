"""Lane programs of the cooperative kernels, emulated on the host: the arithmetic and indexing of a warp-cooperative kernel
can be checked without a GPU by replaying each lane's instructions over an array of lanes (shuffles become array reads)."""
import os, shutil, subprocess
import pytest
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_x12_lane_program_matches_poseidon_permute(tmp_path):
    exe = str(tmp_path / "x12_emu")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "zk-light-client-implementation_b200", "csrc"),
                           os.path.join(ROOT, "tests", "x12_lane_program_emulation.cu"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "x12 emulation == poseidon_permute" in r.stdout, r.stdout + r.stderr

@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_limb_resident_permutation_matches_poseidon_permute(tmp_path):
    """poseidon_permute_limb (the permutation the Merkle kernels run: limb-resident partial rounds, pushed round constants,
    lazy carries) compiled for the host == the textbook permutation on 20 000 states incl. edge values; limb ranges asserted."""
    exe = str(tmp_path / "limb_emu")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "zk-light-client-implementation_b200", "csrc"),
                           os.path.join(ROOT, "tests", "poseidon_limb_emulation.cu"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "limb-resident permutation == poseidon_permute" in r.stdout, r.stdout + r.stderr

@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_device_gate_evaluators_match_generic_on_host(tmp_path):
    """eval_gate<KIND> of gates.cuh (the quotient kernels' evaluators: lazy 128-bit recompositions, non-canonical products,
    batched column loads) compiled for the host == the generic evaluators b200_verify uses, for all 20 constraint-bearing gate
    kinds (several parameterisations) on random, small-valued and near-p rows."""
    exe = str(tmp_path / "gate_emu")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "zk-light-client-implementation_b200", "csrc"),
                           os.path.join(ROOT, "tests", "gate_eval_emulation.cu"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "device gate evaluators == generic evaluators" in r.stdout, r.stdout + r.stderr
    # the precondition of the quotient's sub-coset evaluation: gate_degree() bounds (tightly) every constraint's degree
    assert "gate_degree() bounds every constraint and is tight for 24 gate parameterisations" in r.stdout, r.stdout

def test_limb_table_generator_is_reproducible(tmp_path):
    """the committed poseidon_limb_tables.cuh is what the generator derives from the published round constants (its own
    value-level replay of the limb algorithm against the textbook rounds runs inside the generator)"""
    import shutil as sh
    src = os.path.join(ROOT, "zk-light-client-implementation_b200", "csrc")
    work = tmp_path / "csrc"; (work / "tools").mkdir(parents=True)
    sh.copy(os.path.join(src, "poseidon_tables.cuh"), work)
    sh.copy(os.path.join(src, "tools", "gen_poseidon_limb_tables.py"), work / "tools")
    subprocess.check_call(["python", str(work / "tools" / "gen_poseidon_limb_tables.py")])
    assert open(work / "poseidon_limb_tables.cuh").read() == open(os.path.join(src, "poseidon_limb_tables.cuh")).read()


# ---- mont256.cuh: the even/odd 32-bit-limb Montgomery schedule, limb for limb, in Python ----
M32 = (1 << 32) - 1

def _limbs32(x, n): return [(x >> (32 * i)) & M32 for i in range(n)]

def _mac_row(c, base, xs, y, carry_slot):
    """c[base .. base+7] += sum_k xs[k] * y << 64k as ONE carry chain (MONT_MAC_ROW); the carry lands in c[base+8] or must be 0"""
    carry = 0
    for k in range(4):
        p = xs[k] * y
        s = c[base + 2 * k] + (p & M32) + carry; c[base + 2 * k] = s & M32; carry = s >> 32
        s = c[base + 2 * k + 1] + (p >> 32) + carry; c[base + 2 * k + 1] = s & M32; carry = s >> 32
    if carry_slot:
        s = c[base + 8] + carry; assert s <= M32; c[base + 8] = s
    else:
        assert carry == 0

def _mont256_mul_model(a, b, mod):
    inv64 = (-pow(mod, -1, 1 << 64)) % (1 << 64)
    A, B, m = _limbs32(a, 8), _limbs32(b, 8), _limbs32(mod, 8)
    e, o = [0] * 11, [0] * 9
    for i in range(4):
        assert e[8] == e[9] == e[10] == 0 and o[7] == o[8] == 0
        for phase in range(2):
            if phase == 0: xs, y0, y1 = A, B[2 * i], B[2 * i + 1]
            else:
                q = ((((e[1] + o[0]) & M32) << 32 | e[0]) * inv64) & ((1 << 64) - 1)
                xs, y0, y1 = m, q & M32, q >> 32
            _mac_row(e, 0, xs[0::2], y0, True); _mac_row(o, 0, xs[1::2], y0, True)
            _mac_row(o, 0, xs[0::2], y1, True); _mac_row(e, 2, xs[1::2], y1, True)
        assert e[0] == 0 and (e[1] + o[0]) & M32 == 0 and e[10] == 0           # low 64 bits cleared, E <= T < 2^320
        carry = (e[1] + o[0]) >> 32
        s = e[2] + o[1] + carry; e[2] = s & M32; carry = s >> 32
        for k in range(3, 10):
            s = e[k] + carry; e[k] = s & M32; carry = s >> 32
        assert carry == 0
        e = e[2:10] + [0, 0, 0]; o = o[2:9] + [0, 0]
    r, carry = [e[0]], 0
    for k in range(1, 8):
        s = e[k] + o[k - 1] + carry; r.append(s & M32); carry = s >> 32
    assert carry == 0
    v = sum(x << (32 * i) for i, x in enumerate(r))
    assert v < 2 * mod
    return v - mod if v >= mod else v

@pytest.mark.parametrize("mod", [21888242871839275222246405745257275088696311157297823662689037894645226208583,      # F_q (msm.cu)
                                 21888242871839275222246405745257275088548364400416034343698204186575808495617])     # F_r (bn254.cuh, fr_ntt.cu)
def test_mont256_even_odd_schedule_model(mod):
    import random
    rng = random.Random(mod & 0xFFFF)
    rinv = pow(1 << 256, -1, mod)
    cases = [(mod - 1, mod - 1), (0, mod - 1), (1, 1), (mod - 1, 1)] + [(rng.randrange(mod), rng.randrange(mod)) for _ in range(3000)]
    for a, b in cases:
        assert _mont256_mul_model(a, b, mod) == a * b * rinv % mod
