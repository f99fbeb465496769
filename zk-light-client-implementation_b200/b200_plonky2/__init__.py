"""Host-side mirror of the reference's plonky2 surface for the hot path (CircuitBuilder.build() -> CircuitData,
CircuitData.prove / verify: circuit_builder.py) over the C ABI of libb200plonky2.so (native.py).

The CUDA library is the only implementation: anything that proves, verifies or touches a Context goes through
`b200_plonky2.native`, whose import raises if the .so is not built -- there is no CPU fallback.  The binding is
loaded on first use of one of its names (PEP 562) so that the pure-Python circuit/workload generators
(circuit_builder, workloads: numpy only) can be imported by the CPU reference arm of bench.py without mapping the
product library into that process.
"""

import importlib, os

_SUBMODULES = {f[:-3] for f in os.listdir(os.path.dirname(os.path.abspath(__file__))) if f.endswith(".py") and f != "__init__.py"}

def __getattr__(name):
    if name.startswith("__") or name in _SUBMODULES:      # submodules are imported by the import system, not looked up in native
        raise AttributeError(name)
    native = importlib.import_module(".native", __name__)
    try:
        return getattr(native, name)
    except AttributeError:
        raise AttributeError(f"module 'b200_plonky2' has no attribute {name!r}") from None
