"""`-m "not gpu"` checks of the C-ABI library: it builds, loads, exports exactly what
include/modelardb_cuda.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import pytest

from modelardb_rs_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build_library()
    return _native.lib()


def test_header_symbols_match_binding_list():
    header = open(os.path.join(ROOT, "include", "modelardb_cuda.h")).read()
    declared = set(re.findall(r"\b(mdbcu_[a-z_]+)\s*\(", header))
    assert declared == set(_native.SYMBOLS)


def _c_declarations(text):
    """name -> number of parameters, for every `mdbcu_*(...)` declaration in a C header."""
    out = {}
    for name, args in re.findall(r"\b(mdbcu_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", re.sub(r"/\*.*?\*/", "", text, flags=re.S)):
        args = args.strip()
        out[name] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def _rust_declarations(text):
    """name -> number of parameters, for every `pub fn mdbcu_*(...)` inside an extern block of Rust source."""
    out = {}
    for name, args in re.findall(r"pub fn (mdbcu_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*(?:->[^;]*)?;", re.sub(r"//[^\n]*", "", text)):
        args = args.strip()
        out[name] = 0 if args == "" else len([a for a in args.split(",") if a.strip()])
    return out


@pytest.mark.parametrize("path", ["rust/modelardb_cuda/src/lib.rs", "INTEGRATION.md"])
def test_rust_extern_block_lists_every_entry_point_of_the_header(path):
    """The reference-side binding (rust/: the FFI crate as files; INTEGRATION.md: the same block in the text) cannot be
    compiled here (no cargo / rustc), so at least its declarations are held against the header: every entry point, with the
    header's number of arguments."""
    header = _c_declarations(open(os.path.join(ROOT, "include", "modelardb_cuda.h")).read())
    rust = _rust_declarations(open(os.path.join(ROOT, path)).read())
    assert set(header) == set(_native.SYMBOLS)
    missing = sorted(set(header) - set(rust))
    assert not missing, f"{path} does not declare {missing}"
    wrong = {n: (rust[n], header[n]) for n in header if rust[n] != header[n]}
    assert not wrong, f"{path}: argument counts differ from the header (rust, header): {wrong}"
    assert not sorted(set(rust) - set(header)), f"{path} declares entry points the header does not have"


def test_library_exports_every_declared_symbol(lib):
    for name in _native.SYMBOLS:
        assert hasattr(lib, name), f"{name} is not exported by libmodelardb_cuda.so"


def test_version_and_device_count(lib):
    assert re.fullmatch(r"\d+\.\d+\.\d+", lib.mdbcu_version().decode())
    assert lib.mdbcu_device_count() >= 0


def test_no_gpu_means_failure_not_fallback(lib):
    if lib.mdbcu_device_count() > 0:
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    assert lib.mdbcu_context_create(0, C.byref(ctx)) == 1
    assert b"no CUDA device" in lib.mdbcu_last_error()
    # compute entry points refuse a null context instead of computing anything on the host
    out = C.c_void_p()
    assert lib.mdbcu_compress(None, 0, None, None, None, 0, None, None, C.byref(out)) == 1
    total = C.c_uint64()
    assert lib.mdbcu_grid_count(None, 0, None, None, C.byref(total)) == 1


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "modelardb_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                text = open(os.path.join(dirpath, f)).read()
                assert "mdb_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_error_bound_constructors():
    from modelardb_rs_b200.compression import ErrorBound
    # modelardb_types/src/types.rs:312-334
    for bad in (0.0, -1.0, float("inf"), float("nan")):
        with pytest.raises(ValueError):
            ErrorBound.try_new_absolute(bad)
    for bad in (0.0, -1.0, 100.5, float("nan")):
        with pytest.raises(ValueError):
            ErrorBound.try_new_relative(bad)
    assert ErrorBound.try_new_relative(100.0).kind == 2
    assert ErrorBound.try_new_absolute(1.0).kind == 1


def test_split_into_buffers():
    from modelardb_rs_b200.compression import split_into_buffers
    off = split_into_buffers([1_000_000, 10], capacity=65536)
    assert off[0] == 0 and off[-1] == 1_000_010
    assert len(off) - 1 == 16 + 1
    assert int((off[1:] - off[:-1]).max()) == 65536
