"""CPU checks of the C-ABI library: it loads, and exports every symbol include/hortimapping_b200.h declares."""
import ctypes
import os
import re

from hortimapping_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "hortimapping_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hm_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == decl, "hortimapping_b200/_lib.py EXPORTS is out of sync with the header"
    lib.hm_version.restype = ctypes.c_int
    assert lib.hm_version() >= 200


def test_product_library_exports_no_debug_hooks_and_the_testing_library_has_them():
    """Probe kernels, instrumentation read-outs and single-function hooks live in the test-only superset build
    (libhortimapping_b200_testing.so: -DHM_TESTING + csrc/testing/*.cu); the product library exports the header's ABI only."""
    import subprocess
    from hortimapping_b200 import _testing
    import __graft_entry__ as g
    g.build()
    syms = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (hm_[a-z0-9_]+)", syms)))
    assert exported == declared_symbols(), set(exported) ^ set(declared_symbols())
    tlib = ctypes.CDLL(_testing.LIB_PATH)
    for name in _testing.DEBUG_EXPORTS + declared_symbols():
        assert hasattr(tlib, name), name


def test_struct_layouts_match_header_sizes():
    # hm_opt_params: 10 int32 + 16 doubles; hm_fruit_batch: int32 (+pad) + 14 pointers
    assert ctypes.sizeof(_lib.OptParams) == 10 * 4 + 16 * 8
    assert ctypes.sizeof(_lib.FruitBatch) == 8 + 14 * 8
    assert ctypes.sizeof(_lib.DecoderDesc) == 3 * 4 + 9 * 4 * 2 + 4 + 9 * 8 * 2
    assert ctypes.sizeof(_lib.Counters) == 19 * 8


def test_product_path_has_no_cpu_fallback():
    """No module of the product package may import the oracle (checked textually)."""
    pkg = os.path.join(ROOT, "hortimapping_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in txt and "from oracle" not in txt, fn
