"""CPU-side checks of the drop-in boundary: the library builds, loads, and exports every symbol the
header declares and every symbol the reference's python/cwipc/util.py binds at load time."""
import ctypes
import os
import re
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "cwipc_util_cuda.h")

# symbols python/cwipc/util.py:387-550 touches in cwipc_util_dll_load()
REFERENCE_BOUND_SYMBOLS = """
cwipc_get_version cwipc_log_configure cwipc_dangling_allocations _cwipc_log_emit cwipc_read cwipc_write_ext
cwipc_from_points cwipc_from_packet cwipc_read_debugdump cwipc_write_debugdump
cwipc_pointcloud_free cwipc_pointcloud__shallowcopy cwipc_pointcloud_timestamp cwipc_pointcloud_cellsize
cwipc_pointcloud__set_cellsize cwipc_pointcloud__set_timestamp cwipc_pointcloud_count
cwipc_pointcloud_get_uncompressed_size cwipc_pointcloud_copy_uncompressed cwipc_pointcloud_copy_packet
cwipc_pointcloud_access_metadata
cwipc_activesource_start cwipc_activesource_stop cwipc_activesource_request_metadata
cwipc_activesource_is_metadata_requested cwipc_activesource_reload_config cwipc_activesource_get_config
cwipc_activesource_seek cwipc_activesource_maxtile cwipc_activesource_get_tileinfo
cwipc_activesource_auxiliary_operation
cwipc_source_get cwipc_source_available cwipc_source_eof cwipc_source_free
cwipc_sink_free cwipc_sink_feed cwipc_sink_caption cwipc_sink_interact
cwipc_synthetic cwipc_capturer cwipc_window cwipc_downsample cwipc_remove_outliers cwipc_tilefilter
cwipc_tilemap cwipc_colormap cwipc_crop cwipc_join cwipc_proxy
cwipc_metadata_count cwipc_metadata_name cwipc_metadata_description cwipc_metadata_pointer cwipc_metadata_size
cwipc_write
""".split()


def declared_functions():
    text = open(HEADER).read()
    names = re.findall(r"^_CWIPC_UTIL_EXPORT\s+[^;(]*?\b(\w+)\s*\(", text, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_reference_surface():
    declared = set(declared_functions())
    missing = [s for s in REFERENCE_BOUND_SYMBOLS if s not in declared]
    assert not missing, f"header lacks {missing}"
    assert len(declared) >= len(REFERENCE_BOUND_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    missing = []
    for name in declared_functions():
        try:
            getattr(lib, name)
        except AttributeError:
            missing.append(name)
    assert not missing, f"libcwipc_util_cuda.so does not export {missing}"


def test_dropin_copy_exists(lib):
    from cwipc_util_b200 import build
    assert os.path.exists(os.path.join(build.LIB_DIR, build.DROPIN_NAME))


@pytest.mark.parametrize("compiler,lang", [("gcc", "c"), ("g++", "c++")])
def test_header_compiles_as_c_and_cpp(compiler, lang, tmp_path):
    src = tmp_path / ("t.c" if lang == "c" else "t.cpp")
    src.write_text('#include "cwipc_util/api.h"\nint main(void) { return sizeof(struct cwipc_point) == 16 ? 0 : 1; }\n')
    r = subprocess.run([compiler, "-fsyntax-only", "-Wall", "-I", os.path.join(REPO, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_struct_layouts():
    from cwipc_util_b200 import util
    assert ctypes.sizeof(util.cwipc_point) == 16
    assert util.cwipc_point_numpy_dtype.itemsize == 16
    assert [util.cwipc_point_numpy_dtype.fields[f][1] for f in ("x", "y", "z", "r", "g", "b", "tile")] == [0, 4, 8, 12, 13, 14, 15]


def test_version_and_logging_without_gpu(lib):
    """Entry points that need no device work on a CPU-only box."""
    from cwipc_util_b200 import util
    assert "cwipc_util_cuda" in util.cwipc_get_version()
    seen = []
    util.cwipc_log_configure(util.CWIPC_LOG_LEVEL_WARNING, lambda level, msg: seen.append((level, msg)))
    lib._cwipc_log_emit(util.CWIPC_LOG_LEVEL_WARNING, b"test", b"hello")
    lib._cwipc_log_emit(util.CWIPC_LOG_LEVEL_DEBUG, b"test", b"filtered out")
    util.cwipc_log_configure(util.CWIPC_LOG_LEVEL_WARNING, None)
    assert seen == [(2, b"test: Warning: hello")]
    assert isinstance(util.cwipc_dangling_allocations(False), int)


def test_api_version_is_checked(lib):
    """ref: src/cwipc_util.cpp:663-670 -- a wrong apiVersion yields NULL plus a message, before any device work."""
    err = ctypes.c_char_p()
    rv = lib.cwipc_from_points(None, 0, 0, 0, ctypes.byref(err), 0x20200101)
    assert not rv
    assert b"incorrect apiVersion" in err.value
    err = ctypes.c_char_p()
    assert not lib.cwipc_synthetic(0, 0, ctypes.byref(err), 1)
    assert b"incorrect apiVersion" in err.value


def test_out_of_scope_factories_fail_like_the_reference(lib):
    """capturer/window/proxy: NULL + errorMessage (ref: python/test_cwipc_util.py:423-426)."""
    from cwipc_util_b200 import util
    util.cwipc_log_configure(util.CWIPC_LOG_LEVEL_NONE, lambda level, msg: None)
    try:
        lib.cwipc_capturer.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_ulonglong]
        lib.cwipc_capturer.restype = ctypes.c_void_p
        err = ctypes.c_char_p()
        assert not lib.cwipc_capturer(b'{"type":"nonexistent"}', ctypes.byref(err), util.CWIPC_API_VERSION)
        assert err.value
    finally:
        util.cwipc_log_configure(util.CWIPC_LOG_LEVEL_WARNING, None)


def test_no_device_fails_loudly(lib):
    """Without a CUDA device the product path must raise, not fall back to the CPU."""
    import cwipc_util_b200 as cw
    if cw.cuda_device_count() > 0:
        pytest.skip("a CUDA device is present")
    cw.cwipc_log_configure(cw.CWIPC_LOG_LEVEL_NONE, lambda level, msg: None)
    try:
        with pytest.raises(cw.CwipcError):
            cw.cwipc_from_points([(0, 0, 0, 0, 0, 0, 1)], 0)
    finally:
        cw.cwipc_log_configure(cw.CWIPC_LOG_LEVEL_WARNING, None)


def test_product_does_not_use_oracle():
    """The oracle is test infrastructure: nothing under cwipc_util_b200/ may name it."""
    pkg = os.path.join(REPO, "cwipc_util_b200")
    offenders = []
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".hpp", ".cuh", ".h")):
                text = open(os.path.join(root, f), errors="ignore").read()
                if re.search(r"libcwipc_oracle|import oracle|from oracle|orc_[a-z_]+\(", text):
                    offenders.append(f)
    assert not offenders, offenders


def test_library_has_no_torch_or_triton_dependency(lib):
    from cwipc_util_b200 import build
    out = subprocess.run(["ldd", build.lib_path()], capture_output=True, text=True).stdout
    assert "torch" not in out and "triton" not in out and "libc10" not in out
