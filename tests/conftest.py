import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "zk-light-client-implementation_b200")
if PKG not in sys.path:
    sys.path.insert(0, PKG)

GOLDEN = os.path.join(ROOT, "tests", "golden")

def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    # a fresh checkout has no built artefacts (they are git-ignored): build the product library and the oracle once
    if not os.path.exists(os.path.join(PKG, "libb200plonky2.so")):
        import __graft_entry__
        __graft_entry__.build()

@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
