"""CPU-side boundary checks made with the reference's own artefacts (no GPU, no compute calls):
the apps of the hot path compile and link unchanged against include/ + libcwipc_util_cuda, and the symbol list the
ABI tests use is the one parsed from the reference's python/cwipc/util.py, not a hand-typed copy."""
import os
import re

import pytest

import ref_artifacts as ra
from test_abi_exports import REFERENCE_BOUND_SYMBOLS, declared_functions


def reference_util_py():
    for path in (os.path.join(ra.REFERENCE, "python", "cwipc", "util.py"), os.path.join(ra.REF_DIR, "cwipc", "util.py")):
        if os.path.exists(path):
            return path
    return None


def test_reference_apps_compile_and_link_unchanged(lib):
    """ref: apps/cwipc_downsample/CMakeLists.txt:13, cwipc_remove_outliers/CMakeLists.txt:13, cwipc_tilefilter/CMakeLists.txt:13
    (`target_link_libraries(<app> cwipc_util)`): here the same sources, our header tree, our library."""
    if not ra.reference_present():
        pytest.skip("reference tree not present on this box")
    ra.build_apps(force=True)
    assert ra.apps_built()


def test_bound_symbols_are_the_ones_the_reference_binds(lib):
    """Every `_cwipc_util_dll_reference.<name>.argtypes = ...` of python/cwipc/util.py:387-550 resolves in our library,
    and the hand-typed list of test_abi_exports.py is exactly that set (+ cwipc_write, which only the C++ apps use)."""
    path = reference_util_py()
    if path is None:
        pytest.skip("neither /root/reference nor baseline/_ref is present")
    text = open(path).read()
    start = text.index("def cwipc_util_dll_load")
    bound = sorted(set(re.findall(r"_cwipc_util_dll_reference\.(\w+)\.argtypes", text[start:])))
    assert len(bound) > 50
    missing = [name for name in bound if not hasattr(lib, name)]
    assert not missing, f"library lacks symbols the reference binds at load time: {missing}"
    declared = set(declared_functions())
    assert not [name for name in bound if name not in declared]
    assert set(bound) == set(REFERENCE_BOUND_SYMBOLS) - {"cwipc_write"}
