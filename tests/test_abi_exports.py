"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol that
include/b200_plonky2.h declares; without a GPU the library fails loudly (no CPU fallback)."""
import ctypes, os, re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "zk-light-client-implementation_b200", "libb200plonky2.so")
HDR = os.path.join(ROOT, "include", "b200_plonky2.h")

def _declared():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))

@pytest.fixture(scope="module")
def so():
    if not os.path.exists(SO):
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(SO)

def test_every_declared_symbol_is_exported(so):
    names = _declared()
    assert len(names) >= 18
    missing = [n for n in names if not hasattr(so, n)]
    assert not missing, f"declared in the header but not exported: {missing}"

def test_python_binding_lists_the_same_exports(so):
    import b200_plonky2
    assert sorted(b200_plonky2.EXPORTS) == _declared()

def test_no_cpu_fallback_without_device(so):
    from tests.util import have_gpu
    if have_gpu():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = so.b200_ctx_create(0, ctypes.byref(h))
    assert rc == -2 and not h.value          # B200_ERR_CUDA: the product refuses to run without the device
    import b200_plonky2
    with pytest.raises(b200_plonky2.B200Error):
        b200_plonky2.Context(0)

def test_cpp_host_mirror_compiles_links_and_fails_loudly(so, tmp_path):
    """include/b200_plonky2.hpp (C++ mirror of CircuitData::prove) builds against the .so; without a device the
    Context constructor throws B200_ERR_CUDA (no CPU fallback), with a device a bad circuit is rejected cleanly."""
    import subprocess
    exe = str(tmp_path / "cpp_mirror_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp_mirror_check.cpp"),
                           "-o", exe, "-L", os.path.dirname(SO), "-lb200plonky2", "-Wl,-rpath," + os.path.dirname(SO)])
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert ("no device" in r.stdout) or ("rejected as expected" in r.stdout)
